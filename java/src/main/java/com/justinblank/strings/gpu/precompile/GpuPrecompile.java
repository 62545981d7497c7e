package com.justinblank.strings.gpu.precompile;

import com.justinblank.strings.CompilerOptions;
import com.justinblank.strings.gpu.GpuDFACompiler;
import com.justinblank.strings.gpu.GpuPattern;

import java.io.File;
import java.io.FileOutputStream;
import java.io.IOException;
import java.nio.file.Files;

/**
 * Drop-in for {@code com.justinblank.strings.precompile.Precompile} (precompile/Precompile.java:30-53): same three
 * entry points, but what lands in {@code <directory>/<className>.ndlb} is the table blob of libneedle_b200 (the unit
 * that is also broadcast between GPUs) instead of a JVM class file.  {@link #load} is the counterpart of putting the
 * precompiled class on the class path.
 */
public final class GpuPrecompile {

    private GpuPrecompile() {}

    /** Precompile.precompile(String, String, File) */
    public static void precompile(String regex, String className, File directory) throws IOException {
        precompile(regex, className, directory, 0);
    }

    /** Precompile.precompile(String, String, File, int): returns the path written. */
    public static String precompile(String regex, String className, File directory, int flags) throws IOException {
        return write(GpuDFACompiler.compileToBytes(regex, className, flags), className, directory);
    }

    /** Precompile.precompile(String, String, File, CompilerOptions) */
    public static String precompile(String regex, String className, File directory, CompilerOptions options) throws IOException {
        return write(GpuDFACompiler.compileToBytes(regex, className, options), className, directory);
    }

    /** A pattern from a precompiled blob, resident on {@code device} (-1: every visible GPU). */
    public static GpuPattern load(File blobFile, int device) throws IOException {
        return new GpuPattern(Files.readAllBytes(blobFile.toPath()), device);
    }

    private static String write(byte[] blob, String className, File directory) throws IOException {
        String target = directory.getAbsolutePath() + "/" + className + ".ndlb";
        try (FileOutputStream fos = new FileOutputStream(target)) {
            fos.write(blob);
        }
        return target;
    }
}
