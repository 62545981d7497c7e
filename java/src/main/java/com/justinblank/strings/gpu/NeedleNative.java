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

    /** ndl_pattern_create (device -1: one replica per visible GPU, tables NCCL-broadcast); throws RuntimeException("NDL_ECUDA ...")
     *  when no GPU is usable (no CPU fallback). */
    static native long patternCreate(byte[] blob, int device);

    static native void patternDestroy(long handle);

    /** ndl_match_batch with NDL_MEM_HOST on direct buffers. */
    static native void matchBatch(long handle, int mode, ByteBuffer data, ByteBuffer offsets, int n, int charWidth,
                                  byte[] matched, int[] start, int[] end);

    /** One string through ndl_match_batch (GetStringCritical -> UTF-16 code units, char_width 2). Returns {matched, start, end}. */
    static native int[] matchOne(long handle, int mode, String s, int from);

    /** ndl_match_lines with NDL_MEM_HOST: n fixed-length records of lineChars chars in a direct buffer, no offsets. */
    static native void matchLines(long handle, int mode, ByteBuffer data, int n, int lineChars, int charWidth,
                                  byte[] matched, int[] start, int[] end);

    /**
     * ndl_find_all_batch with NDL_MEM_HOST on direct buffers: the loop {@code while (m.find())} for every haystack.
     * Pass 1: {@code matchOffsets == null} fills {@code counts}; pass 2: {@code matchOffsets} = exclusive prefix sum
     * (n + 1 longs, little endian), {@code starts}/{@code ends} sized to its last element.
     */
    static native void findAllBatch(long handle, ByteBuffer data, ByteBuffer offsets, int n, int charWidth, int[] counts,
                                    ByteBuffer matchOffsets, int[] starts, int[] ends);

    /** Packs the strings with GetStringRegion into one page-locked staging buffer (ndl_host_alloc) and runs find() on all of them. */
    static native GpuPattern.BatchResult findAllStrings(long handle, String[] haystacks);

    /** ndl_find_long with NDL_MEM_HOST on a direct buffer: one haystack, 64-bit indices.  Returns {matched, start, end}. */
    static native long[] findLong(long handle, ByteBuffer data, long nChars, int charWidth, long from);

    /** ndl_host_alloc: a direct ByteBuffer over page-locked memory - batches in it are DMA'd without a bounce copy. */
    static native ByteBuffer pinnedAlloc(long bytes);

    /** ndl_host_free for a buffer of {@link #pinnedAlloc}; the buffer must not be used afterwards. */
    static native void pinnedFree(ByteBuffer buffer);
}
