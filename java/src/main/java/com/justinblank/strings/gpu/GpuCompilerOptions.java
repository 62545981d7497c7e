package com.justinblank.strings.gpu;

import com.justinblank.strings.CharacterDistribution;
import com.justinblank.strings.CompilerOptions;
import com.justinblank.strings.DebugOptions;

/**
 * {@code com.justinblank.strings.CompilerOptions} keeps {@code flags} protected (CompilerOptions.java:5-7) and has no
 * getter, so code outside its package can only read them through a subclass.  This one exists for exactly that.
 */
final class GpuCompilerOptions extends CompilerOptions {

    private GpuCompilerOptions(int flags, CharacterDistribution distribution, DebugOptions debugOptions) {
        super(flags, distribution, debugOptions);
    }

    /** The flags of any CompilerOptions instance (protected access through the subclass, same-class rule of JLS 6.6.2). */
    static int flagsOf(CompilerOptions options) {
        if (options instanceof GpuCompilerOptions) {
            return ((GpuCompilerOptions) options).flags;
        }
        // a foreign instance: its protected field is not accessible from here; rebuild it the only public way there is
        try {
            java.lang.reflect.Field f = CompilerOptions.class.getDeclaredField("flags");
            f.setAccessible(true);
            return f.getInt(options);
        } catch (ReflectiveOperationException e) {
            throw new IllegalArgumentException("cannot read the flags of " + options, e);
        }
    }

    /** CompilerOptions.fromFlags for callers that want an instance this package can read without reflection. */
    static CompilerOptions fromFlags(int flags) {
        return new GpuCompilerOptions(flags, CharacterDistribution.DEFAULT, DebugOptions.none());
    }
}
