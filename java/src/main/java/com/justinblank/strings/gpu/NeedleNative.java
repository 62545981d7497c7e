package com.justinblank.strings.gpu;

import java.nio.ByteBuffer;

/** JNI bindings of include/needle_b200.h.  Implemented in src/main/native/needle_jni.c. */
final class NeedleNative {

    static {
        System.loadLibrary("needle_jni");   // which links libneedle_b200.so
    }

    private NeedleNative() {}

    /** ndl_compile; throws PatternSyntaxException / IllegalStateException / RuntimeException per NDL_E* code. */
    static native byte[] compile(String regex, int flags);

    /** ndl_pattern_create; throws RuntimeException("NDL_ECUDA ...") when no GPU is usable (no CPU fallback). */
    static native long patternCreate(byte[] blob, int device);

    static native void patternDestroy(long handle);

    /** ndl_match_batch with NDL_MEM_HOST on direct buffers. */
    static native void matchBatch(long handle, int mode, ByteBuffer data, ByteBuffer offsets, int n, int charWidth,
                                  byte[] matched, int[] start, int[] end);

    /** One string through ndl_match_batch (GetStringCritical -> UTF-16 code units, char_width 2). Returns {matched, start, end}. */
    static native int[] matchOne(long handle, int mode, String s, int from);

    /** Packs the strings with GetStringRegion into one pinned staging buffer and runs find() on all of them. */
    static native GpuPattern.BatchResult findAllStrings(long handle, String[] haystacks);
}
