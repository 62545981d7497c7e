package com.justinblank.strings.gpu;

import com.justinblank.strings.Matcher;

/**
 * {@link Matcher} over one string.  Same mutable state as the generated class (nextStart, start, end -
 * DFAClassBuilder.java:688-694), so it is not thread safe either.  Every call is one ndl_match_batch with
 * n = 1; use {@link GpuPattern#matchBatch} for throughput.
 *
 * <p>Empty matches: like the reference (DFAClassBuilder.java:634-635), {@code find()} does not move nextStart past an empty
 * match, so {@code while (m.find())} over a pattern that matches the empty string does not terminate - this class keeps
 * that behaviour on purpose.  {@link GpuPattern#findAllBatch} reports such a match once and ends the haystack's list.
 */
final class GpuMatcher implements Matcher {

    private final GpuPattern pattern;
    private final String string;
    private int nextStart = 0;
    private int start = -1;
    private int end = -1;

    GpuMatcher(GpuPattern pattern, String string) {
        this.pattern = pattern;
        this.string = string;
    }

    @Override
    public boolean matches() {
        return NeedleNative.matchOne(pattern.handle, 0, string, 0)[0] != 0;
    }

    @Override
    public boolean containedIn() {
        return NeedleNative.matchOne(pattern.handle, 1, string, 0)[0] != 0;
    }

    @Override
    public boolean find() {
        return find(nextStart, string.length());
    }

    /** Like the reference (DFAClassBuilder.java:349), `to` is ignored by the forward scan. */
    @Override
    public boolean find(int from, int to) {
        if (nextStart == -1) {
            return false;
        }
        int[] r = NeedleNative.matchOne(pattern.handle, 2, string, from);   // {matched, start, end}
        end = nextStart = r[2];
        if (r[0] != 0) {
            start = r[1];
            return true;
        }
        return false;
    }

    @Override
    public int start() {
        return start;
    }

    @Override
    public int end() {
        return end;
    }
}
