package com.justinblank.strings.gpu;

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
