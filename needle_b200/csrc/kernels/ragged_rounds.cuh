// Ragged batches of LONGER lines (mean length >= kRrMinMeanBytes: log lines, JSON lines): the sorted streaming walk.
// Included by lines8.cuh; shared by lines8_kernel and linesq_kernel.
//
// The ragged TILE walk (l8_run_ragged) stages 2 KB tiles: with lines of 150 bytes a tile holds 13 of them, a pair of tiles 26 -
// lanes idle before any difference in length - and from about 250 bytes on it streams 32 consecutive lines at a time, where
// a warp walks as long as its longest line.  Here every warp
//   1. takes a WINDOW of 128 consecutive lines (64 or 32 when the batch has few lines per warp) and sorts them by walk length in
//      registers (a bitonic network over 4 keys per lane: no shared memory, no CTA barrier);
//   2. walks the window as four batches of 32 lines of nearly equal length with the streaming walk of the long-line path
//      (l8_stream_lines: every lane copies its own line 64 bytes at a time with cp.async into its slots of the tile buffers,
//      double buffered, and walks them from shared memory): all lanes busy, no planning, no pairing, any line length;
//   3. collects the lines whose find() needs the table-driven reverse pass (indexBackwards, DFAClassBuilder.java:529-586) in a
//      queue of one entry per lane and, whenever 32 are waiting, walks them backwards in lockstep straight from L2
//      (16-byte loads, one chunk ahead): only lines that matched, so every lane has work.
// Results are those of the generated loops of the reference (indexForwards :335-471, indexBackwards :529-614, glue :616-667),
// bit for bit - per line, the arithmetic is the tile walk's.  For SHORT lines (the e-mail config: 8..120 bytes) this engine was
// measured at par with the tile walk (exp/kernels_tried/bucket.cuh, its ancestor), so those keep the tile walk.
#pragma once

namespace ndl {

constexpr uint32_t kRrWindow = 128;        // lines per window (4 per lane)
constexpr uint32_t kRrMinLines = 4096;     // smaller ragged batches keep the tile walk
constexpr uint32_t kRrMinMeanBytes = 96;   // mean line length from which the sorted streaming walk is used

__device__ __forceinline__ uint4 rr_ldg16(const uint8_t* ptr) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(ptr));
  return v;
}

// The reverse queue: lines that matched and wait for indexBackwards(end - 1, from) (DFAClassBuilder.java:529-586), one per lane.
// Whenever 32 are waiting they are walked backwards in lockstep straight from global memory / L2 (16-byte loads, one chunk ahead):
// only lines that matched, so every lane has work - where each lane running the generic loop over its own matches leaves most
// lanes idle.  Used by the sorted streaming walk below and by the rounds walks of fixed-length records (lines8.cuh): in both the
// line is no longer in shared memory when its forward walk ends.
template <int CM>
struct L8RevQueue {
  uint32_t qn = 0, q_line = 0;
  int32_t q_end = 0, q_from = 0;

  // lockstep over the first `count` entries
  __device__ __forceinline__ void run(const Lines8Params& p, const L8Ctx& cx, const uint32_t lane, const uint32_t count) {
    constexpr uint32_t kCharBytes = L8Chars<CM>::kBytes;
    constexpr uint32_t kPer = L8Chars<CM>::kPerChunk;
    constexpr uint32_t kFull = 0xffffffffu;
    constexpr uint32_t kStateMask = L8Enc<CM>::kStateMask;
    const BatchParams& g = p.g;
    const uint8_t* const data = static_cast<const uint8_t*>(g.data);
    const bool has = lane < count;
    uint32_t total = 0;  // chars to walk: [from, end)
    uint64_t o0 = 0;
    if (has) {
      o0 = batch_off(g, q_line);
      total = static_cast<uint32_t>(q_end - q_from);
    }
    const uint8_t* const lo_ptr = data + (o0 + static_cast<uint32_t>(q_from)) * kCharBytes;  // first byte that may be walked
    const uint8_t* const h_ptr = lo_ptr + static_cast<uint64_t>(total) * kCharBytes;          // one past the last
    // window k = bytes [h - 16 (k + 1), h - 16 k): chunk x holds its first byte, chunk y the rest
    const uint8_t* const w0p = h_ptr - 16;
    const uint32_t a = static_cast<uint32_t>(reinterpret_cast<uintptr_t>(w0p)) & 15u;
    const uint8_t* xp = w0p - a;
    const uint4 zero = make_uint4(0, 0, 0, 0);
    // (a chunk is only read when it holds a byte of [lo_ptr, h_ptr): nothing outside the line is touched)
    uint4 y = (total != 0 && a != 0) ? rr_ldg16(xp + 16) : zero;
    uint4 x = (total != 0 && xp + 16 > lo_ptr) ? rr_ldg16(xp) : zero;
    const L8Align al(a);
    uint32_t e = cx.bwd_root, pos = 0;
    int32_t w = g.bwd.root_accepting ? static_cast<int32_t>(total) : -1;  // lastMatch = lowerBound when the root accepts (:543-547)
    while (__ballot_sync(kFull, pos < total) != 0) {
      xp -= 16;
      const uint4 x_next = (pos + kPer < total && xp + 16 > lo_ptr) ? rr_ldg16(xp) : zero;  // one chunk ahead of the walk
      if (pos < total) {
        const uint4 wv = al.apply(x, y);
        uint32_t mask = 0;
        l8_chunk_rev<CM>(wv, p.q, cx, e, mask);
        const uint32_t valid = min(kPer, total - pos);
        mask >>= (kPer - valid);  // drop the steps taken before `from`
        const int32_t cand = static_cast<int32_t>(pos + valid + 1) - __ffs(mask);  // chars walked up to the last accepting step
        w = mask ? cand : w;
        pos += kPer;
        if ((e & kStateMask) == cx.bwd_dead) pos = total;
      }
      y = x;
      x = x_next;
    }
    if (has) g.start[q_line] = w == -1 ? 0x7fffffff : static_cast<int32_t>(total) - w + q_from;
  }

  // append the lanes of `rv` (their line / end / from); runs a batch whenever 32 entries are waiting
  __device__ __forceinline__ void push(const Lines8Params& p, const L8Ctx& cx, const uint32_t lane, const uint32_t rv, const uint32_t line,
                                       const int32_t end, const int32_t from) {
    constexpr uint32_t kFull = 0xffffffffu;
    const uint32_t k = __popc(rv);
    {  // slot s of the queue takes the (s - qn)-th lane of rv
      const bool mine = lane >= qn && lane < qn + k;
      const uint32_t src = mine ? __fns(rv, 0, static_cast<int>(lane - qn) + 1) & 31u : lane;
      const uint32_t nl = __shfl_sync(kFull, line, src);
      const int32_t ne = __shfl_sync(kFull, end, src), nf = __shfl_sync(kFull, from, src);
      if (mine) { q_line = nl; q_end = ne; q_from = nf; }
    }
    if (qn + k >= 32) {
      run(p, cx, lane, 32);
      const uint32_t done = 32 - qn, rest = k - done;  // `done` lanes of rv went into the batch; the rest start a new queue
      const bool mine = lane < rest;
      const uint32_t src = mine ? __fns(rv, 0, static_cast<int>(done + lane) + 1) & 31u : lane;
      const uint32_t nl = __shfl_sync(kFull, line, src);
      const int32_t ne = __shfl_sync(kFull, end, src), nf = __shfl_sync(kFull, from, src);
      if (mine) { q_line = nl; q_end = ne; q_from = nf; }
      qn = rest;
    } else {
      qn += k;
    }
  }

  __device__ __forceinline__ void flush(const Lines8Params& p, const L8Ctx& cx, const uint32_t lane) {
    if (qn) run(p, cx, lane, qn);
    qn = 0;
  }
};

// Ascending bitonic sort of 128 keys, element e = 32 j + lane in key[j].
__device__ __forceinline__ void rr_sort128(uint32_t (&key)[4], const uint32_t lane) {
#pragma unroll
  for (uint32_t k = 2; k <= 128; k <<= 1) {
#pragma unroll
    for (uint32_t d = k >> 1; d > 0; d >>= 1) {
      if (d >= 32) {
        const uint32_t jd = d >> 5;
#pragma unroll
        for (uint32_t j = 0; j < 4; j++) {
          if ((j & jd) == 0) {
            const bool up = ((32 * j) & k) == 0;  // (k >= 64 here, so bit k of e is a bit of j)
            const uint32_t a = key[j], b = key[j | jd];
            const uint32_t lo = min(a, b), hi = max(a, b);
            key[j] = up ? lo : hi;
            key[j | jd] = up ? hi : lo;
          }
        }
      } else {
#pragma unroll
        for (uint32_t j = 0; j < 4; j++) {
          const uint32_t other = __shfl_xor_sync(0xffffffffu, key[j], d);
          const bool up = ((32 * j + lane) & k) == 0;
          const bool lower = (lane & d) == 0;
          key[j] = (lower == up) ? min(key[j], other) : max(key[j], other);
        }
      }
    }
  }
}

template <int CM>
__device__ __forceinline__ void l8_run_ragged_rounds(const Lines8Params& p, const L8Ctx& cx, const uint32_t buf0, const uint32_t buf1,
                                                     const uint32_t lane, const uint32_t warp_global, const uint32_t n_warps) {
  using CharT = typename std::conditional<L8Chars<CM>::kBytes == 1, uint8_t, uint16_t>::type;
  constexpr uint32_t kCharBytes = L8Chars<CM>::kBytes;
  constexpr uint32_t kFull = 0xffffffffu;
  const BatchParams& g = p.g;
  const uint32_t n = static_cast<uint32_t>(g.n);
  // windows are dealt out round-robin; smaller windows when there are few lines per warp (long lines), so that every warp has
  // several and the last ones finish together
  const uint32_t win = n >= 3u * kRrWindow * n_warps ? kRrWindow : n >= 3u * 64u * n_warps ? 64u : 32u;
  const uint32_t n_windows = (n + win - 1) / win;
  const uint32_t mode = static_cast<uint32_t>(g.mode);
  const bool use_from = g.from != nullptr && mode == 2;
  const bool defer_rev = mode == 2 && g.reverse_mode == 0 && p.has_bwd != 0;

  L8RevQueue<CM> queue;  // lines that matched and wait for their table-driven reverse pass

  for (uint32_t wi = warp_global; wi < n_windows; wi += n_warps) {
    const uint32_t w0 = wi * win;
    const uint32_t m = min(win, n - w0);
    // ---- 1. the window's lines sorted by walk length (16-byte steps): key = steps << 7 | index in the window
    uint32_t key[4];
#pragma unroll
    for (uint32_t k = 0; k < 4; k++) {
      const uint32_t j = lane + 32 * k;
      key[k] = kFull;
      if (j < m) {
        const uint64_t len = batch_off(g, w0 + j + 1) - batch_off(g, w0 + j);
        const uint64_t steps = (len * kCharBytes + 15) >> 4;
        key[k] = (steps < (1u << 24) ? static_cast<uint32_t>(steps) : (1u << 24)) << 7 | j;
      }
    }
    rr_sort128(key, lane);
    // ---- 2. four batches of 32 lines of nearly equal length (batch b = key[b] across the lanes)
#pragma unroll 1
    for (uint32_t b = 0; b * 32 < m; b++) {
      const uint32_t kb = b == 0 ? key[0] : b == 1 ? key[1] : b == 2 ? key[2] : key[3];
      const bool own = kb != kFull;
      const uint32_t line = w0 + (own ? (kb & 127u) : 0u);
      int32_t end = 0, from = 0;
      const bool want_rev = l8_stream_lines<CM, CharT>(p, cx, buf0, buf1, lane, own, line, use_from, defer_rev, &end, &from);
      if (defer_rev) {
        const uint32_t rv = __ballot_sync(kFull, want_rev);
        if (rv) queue.push(p, cx, lane, rv, line, end, from);
      }
    }
  }
  queue.flush(p, cx, lane);
}

}  // namespace ndl
