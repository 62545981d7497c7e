package com.justinblank.strings.gpu;

import com.justinblank.strings.CompilerOptions;
import com.justinblank.strings.Pattern;
import com.justinblank.strings.PatternClassCompilationException;

import java.nio.charset.StandardCharsets;
import java.util.Objects;

/**
 * Drop-in for {@code com.justinblank.strings.DFACompiler}: same static entry points, but the regex is
 * compiled to a table blob by libneedle_b200 (ndl_compile) and matched on the GPU (ndl_match_batch)
 * instead of being turned into a JVM class.
 *
 * NOT COMPILED IN THIS REPOSITORY'S CI: the build image has no JDK.  The class is a mechanical mirror of
 * include/needle_b200.h; see INTEGRATION.md.
 */
public final class GpuDFACompiler {

    private GpuDFACompiler() {}

    /** DFACompiler.compile(String, String) */
    public static Pattern compile(String regex, String className) {
        return compile(regex, className, 0);
    }

    /** DFACompiler.compile(String, String, int) */
    public static Pattern compile(String regex, String className, int flags) {
        Objects.requireNonNull(className, "name cannot be null");
        byte[] blob = compileToBytes(regex, className, flags);
        return new GpuPattern(blob, /*device=*/0);
    }

    /**
     * DFACompiler.compile(String, String, CompilerOptions) (DFACompiler.java:25).  Of the options only the flags reach the
     * tables: the character distribution steers the reference's choice of search accelerators (recorded in the blob for
     * the CPU baseline, not used on the GPU) and the debug options print JVM-side artefacts that do not exist here.
     * CompilerOptions keeps its fields protected, so the flags are read through {@link GpuCompilerOptions}.
     */
    public static Pattern compile(String regex, String className, CompilerOptions options) {
        Objects.requireNonNull(options, "options cannot be null");
        return compile(regex, className, GpuCompilerOptions.flagsOf(options));
    }

    /** DFACompiler.compileToBytes(String, String, CompilerOptions) */
    public static byte[] compileToBytes(String regex, String className, CompilerOptions options) {
        Objects.requireNonNull(options, "options cannot be null");
        return compileToBytes(regex, className, GpuCompilerOptions.flagsOf(options));
    }

    /** Every visible GPU: one replica per device, tables sent with one NCCL broadcast; batches are sharded by the library. */
    public static Pattern compileOnAllDevices(String regex, String className, int flags) {
        Objects.requireNonNull(className, "name cannot be null");
        return new GpuPattern(compileToBytes(regex, className, flags), /*device=*/-1);
    }

    /** DFACompiler.compileToBytes: returns the table blob (what Precompile writes to disk). */
    public static byte[] compileToBytes(String regex, String className, int flags) {
        Objects.requireNonNull(regex, "regex string cannot be null");
        if ((flags & ~Pattern.ALL_FLAGS) != 0) {
            throw new IllegalArgumentException("Unrecognized flags=" + flags);   // CompilerOptions.java:10-12
        }
        try {
            return NeedleNative.compile(regex, flags);                           // throws on NDL_E*
        } catch (RuntimeException e) {
            // DFACompiler.compileToBytes wraps every failure, including the parser's PatternSyntaxException
            throw new PatternClassCompilationException("Failed to create pattern class for regex '" + regex + "'", e);
        }
    }
}
