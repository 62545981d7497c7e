package com.justinblank.strings.gpu;

import com.justinblank.strings.Matcher;
import com.justinblank.strings.Pattern;

import java.nio.ByteBuffer;
import java.nio.ByteOrder;

/**
 * {@link Pattern} backed by an {@code ndl_pattern*}.  Stateless and shareable between threads, like the
 * class the reference generates (DFACompiler.java:96-111).  Adds the one call a GPU needs: a batch.
 */
public final class GpuPattern implements Pattern, AutoCloseable {

    final long handle;          // ndl_pattern*
    private final byte[] blob;

    public GpuPattern(byte[] blob, int device) {
        this.blob = blob;
        this.handle = NeedleNative.patternCreate(blob, device);
    }

    @Override
    public Matcher matcher(String s) {
        return new GpuMatcher(this, s);
    }

    /** Result of one batch: matched[i], start[i], end[i] exactly as Matcher.find()/start()/end() would give. */
    public static final class BatchResult {
        public final byte[] matched;
        public final int[] start;
        public final int[] end;

        BatchResult(int n) {
            matched = new byte[n];
            start = new int[n];
            end = new int[n];
        }
    }

    /**
     * Run {@code mode} (0 matches, 1 containedIn, 2 find) over n haystacks packed in a direct buffer.
     *
     * @param data      direct ByteBuffer with the packed chars (charWidth 1 = Latin-1 bytes, 2 = UTF-16LE)
     * @param offsets   n+1 offsets in chars (direct LongBuffer backing store)
     */
    public BatchResult matchBatch(int mode, ByteBuffer data, ByteBuffer offsets, int n, int charWidth) {
        BatchResult r = new BatchResult(n);
        NeedleNative.matchBatch(handle, mode, data, offsets.order(ByteOrder.LITTLE_ENDIAN), n, charWidth, r.matched, r.start, r.end);
        return r;
    }

    /** {@code mode} over n fixed-length records of {@code lineChars} chars (no offsets array). */
    public BatchResult matchLines(int mode, ByteBuffer data, int n, int lineChars, int charWidth) {
        BatchResult r = new BatchResult(n);
        NeedleNative.matchLines(handle, mode, data, n, lineChars, charWidth, r.matched, r.start, r.end);
        return r;
    }

    /**
     * All non-overlapping matches of every haystack (CSR): {@code while (m.find())} per haystack, in two passes.
     * A match that does not move the search position forward (an empty match) is reported once and ends that haystack's list;
     * the reference's {@code while (m.find())} loop would report it forever.
     */
    public static final class AllMatches {
        public int[] counts;
        public long[] matchOffsets;
        public int[] starts;
        public int[] ends;
    }

    public AllMatches findAllBatch(ByteBuffer data, ByteBuffer offsets, int n, int charWidth) {
        AllMatches r = new AllMatches();
        r.counts = new int[n];
        ByteBuffer off = offsets.order(ByteOrder.LITTLE_ENDIAN);
        NeedleNative.findAllBatch(handle, data, off, n, charWidth, r.counts, null, null, null);
        r.matchOffsets = new long[n + 1];
        ByteBuffer mo = ByteBuffer.allocateDirect(8 * (n + 1)).order(ByteOrder.LITTLE_ENDIAN);
        long total = 0;
        for (int i = 0; i < n; i++) {
            mo.putLong(8 * i, total);
            r.matchOffsets[i] = total;
            total += r.counts[i] & 0xffffffffL;
        }
        mo.putLong(8 * n, total);
        r.matchOffsets[n] = total;
        r.starts = new int[(int) total];
        r.ends = new int[(int) total];
        NeedleNative.findAllBatch(handle, data, off, n, charWidth, r.counts, mo, r.starts, r.ends);
        return r;
    }

    /** Convenience: find() on every string. */
    /**
     * find() over ONE haystack in a direct buffer (a mapped file, a document) - no per-line structure, 64-bit indices.  On a
     * pattern compiled for all devices the library cuts the haystack into one chunk per GPU.  Returns {matched (0/1), start, end};
     * start = end = -1 without a match.
     */
    public long[] findLong(ByteBuffer data, long nChars, int charWidth, long from) {
        return NeedleNative.findLong(handle, data, nChars, charWidth, from);
    }

    public BatchResult findAll(String[] haystacks) {
        return NeedleNative.findAllStrings(handle, haystacks);
    }

    /**
     * A direct buffer over page-locked memory (ndl_host_alloc) for batch data / offsets: the GPU's copy engine reads it
     * directly.  Heap or ordinary direct buffers work too - the library then stages them through its own pinned ring.
     * Release it with {@link #freePinned}.
     */
    public static ByteBuffer allocatePinned(long bytes) {
        return NeedleNative.pinnedAlloc(bytes).order(ByteOrder.LITTLE_ENDIAN);
    }

    public static void freePinned(ByteBuffer buffer) {
        NeedleNative.pinnedFree(buffer);
    }

    public byte[] blob() {
        return blob.clone();
    }

    @Override
    public void close() {
        NeedleNative.patternDestroy(handle);
    }
}
