// Large ragged batches: the bucketed streaming walk.  Included by lines8.cuh; shared by lines8_kernel and linesq_kernel.
//
// The ragged TILE walk (l8_run_ragged) gives every lane one line of a 32-line tile, so a warp walks as long as its longest
// line; sorting and pairing the lines of two tiles recovers most of that, at 50 instructions per step for the planning, and
// the reverse passes of find() are pooled per tile pair - few jobs of very different length, 26 % lane utilisation, 28 % of
// all instructions on the e-mail regex.  Here, instead, every warp
//   1. takes a WINDOW of 128 consecutive lines of its range and counting-sorts them by walk length (16-byte steps) in a
//      few hundred bytes of shared memory of its own - no CTA barrier anywhere;
//   2. walks the window as four batches of 32 lines of (nearly) equal length with the streaming walk of the long-line
//      path (l8_stream_lines: every lane copies its own line 64 bytes at a time with cp.async into its slots of the tile
//      buffers, double buffered, and walks them from shared memory) - equal lengths, so no lane waits for a neighbour, no
//      planning, no pairing, and lines of any length take the same path; the window's bytes were asked for one window
//      ahead with a bulk L2 prefetch (SASS UBLKPF), so the DRAM side stays a sequential stream;
//   3. collects the lines whose find() needs the table-driven reverse pass (indexBackwards, DFAClassBuilder.java:529-586)
//      in a queue of one entry per lane and, whenever 32 are waiting, walks them backwards in lockstep straight from L2
//      (16-byte loads, one chunk ahead): only lines that matched, so every lane has work.
// Results are those of the generated loops of the reference (indexForwards :335-471, indexBackwards :529-614, glue
// :616-667), bit for bit - per line, the arithmetic is the tile walk's.
#pragma once

namespace ndl {

constexpr uint32_t kBkWindow = 128;         // lines per window (4 per lane)
constexpr uint32_t kBkMinLines = 65536;     // smaller ragged batches keep the tile walk
constexpr uint32_t kBkScratchBytes = 512;   // shared memory per warp: 64 counters (u32) + the window's permutation (u8)
constexpr uint32_t kBkDonorWarps = 4;       // the last warps of the CTA give their tile buffers (16 KB) as scratch and sit out
constexpr uint32_t kBkBuckets = 64;

__device__ __forceinline__ uint4 bk_ldg16(const uint8_t* ptr) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(ptr));
  return v;
}
__device__ __forceinline__ void bk_sts32(uint32_t addr, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t bk_lds32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void bk_sts8(uint32_t addr, uint32_t v) { asm volatile("st.shared.u8 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t bk_lds8(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t bk_atom_add(uint32_t addr, uint32_t v) {
  uint32_t old;
  asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(addr), "r"(v) : "memory");
  return old;
}

// warp_walker / n_walkers: this warp's index among, and the number of, the warps of the grid that walk (donors excluded);
// scratch: this warp's kBkScratchBytes of shared memory.
template <int CM>
__device__ __forceinline__ void l8_run_bucketed(const Lines8Params& p, const L8Ctx& cx, const uint32_t buf0, const uint32_t buf1,
                                                const uint32_t lane, const uint32_t warp_walker, const uint32_t n_walkers,
                                                const uint32_t scratch) {
  using CharT = typename std::conditional<L8Chars<CM>::kBytes == 1, uint8_t, uint16_t>::type;
  constexpr uint32_t kCharBytes = L8Chars<CM>::kBytes;
  constexpr uint32_t kPer = L8Chars<CM>::kPerChunk;
  constexpr uint32_t kFull = 0xffffffffu;
  constexpr uint32_t kStateMask = L8Enc<CM>::kStateMask;
  const BatchParams& g = p.g;
  const uint8_t* const data = static_cast<const uint8_t*>(g.data);
  const uint32_t n = static_cast<uint32_t>(g.n);
  // contiguous ranges of whole windows per warp
  const uint32_t n_windows = (n + kBkWindow - 1) / kBkWindow;
  const uint32_t win_per_warp = (n_windows + n_walkers - 1) / n_walkers;
  const uint32_t lo = min(n, warp_walker * win_per_warp * kBkWindow), hi = min(n, lo + win_per_warp * kBkWindow);
  const uint32_t mode = static_cast<uint32_t>(g.mode);
  const bool use_from = g.from != nullptr && mode == 2;
  const bool defer_rev = mode == 2 && g.reverse_mode == 0 && p.has_bwd != 0;
  const uint32_t perm = scratch + 4 * kBkBuckets;

  auto prefetch = [&](uint32_t w0) {  // the bytes of the window that starts at line w0 -> L2 (at most 256 KB of them)
    if (lane == 0 && w0 < hi) {
      const uint64_t b0 = batch_off(g, w0) * kCharBytes & ~static_cast<uint64_t>(15);
      const uint64_t b1 = batch_off(g, min(w0 + kBkWindow, hi)) * kCharBytes;
      uint64_t left = min(b1 - b0, static_cast<uint64_t>(256u << 10)) & ~static_cast<uint64_t>(15);
      const uint8_t* ptr = data + b0;
      while (left) {
        const uint32_t piece = static_cast<uint32_t>(min(left, static_cast<uint64_t>(32u << 10)));
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(ptr), "r"(piece) : "memory");
        ptr += piece;
        left -= piece;
      }
    }
  };

  // ---- the reverse queue: lane l < qn holds a line that matched and waits for indexBackwards(end - 1, from)
  uint32_t qn = 0, q_line = 0;
  int32_t q_end = 0, q_from = 0;
  auto reverse_batch = [&](uint32_t count) {  // lockstep over the first `count` entries, straight from global memory / L2
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
    uint4 y = (total != 0 && a != 0) ? bk_ldg16(xp + 16) : zero;
    uint4 x = (total != 0 && xp + 16 > lo_ptr) ? bk_ldg16(xp) : zero;
    const L8Align al(a);
    uint32_t e = cx.bwd_root, pos = 0;
    int32_t w = g.bwd.root_accepting ? static_cast<int32_t>(total) : -1;  // lastMatch = lowerBound when the root accepts (:543-547)
    while (__ballot_sync(kFull, pos < total) != 0) {
      xp -= 16;
      const uint4 x_next = (pos + kPer < total && xp + 16 > lo_ptr) ? bk_ldg16(xp) : zero;  // one chunk ahead of the walk
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
  };
  // append the lanes of `rv` (their line / end / from) to the queue; runs a reverse batch whenever 32 entries are waiting
  auto enqueue = [&](uint32_t rv, uint32_t line, int32_t end, int32_t from) {
    const uint32_t k = __popc(rv);
    // slot s of the queue takes the (s - qn)-th lane of rv
    {
      const bool mine = lane >= qn && lane < qn + k;
      const uint32_t src = mine ? __fns(rv, 0, static_cast<int>(lane - qn) + 1) & 31u : lane;
      const uint32_t nl = __shfl_sync(kFull, line, src);
      const int32_t ne = __shfl_sync(kFull, end, src), nf = __shfl_sync(kFull, from, src);
      if (mine) { q_line = nl; q_end = ne; q_from = nf; }
    }
    if (qn + k >= 32) {
      reverse_batch(32);
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
  };

  prefetch(lo);
  for (uint32_t w0 = lo; w0 < hi; w0 += kBkWindow) {
    const uint32_t m = min(kBkWindow, hi - w0);
    prefetch(w0 + kBkWindow);
    // ---- 1. counting sort of the window's lines by walk length, longest first
    uint32_t bucket[4];
#pragma unroll
    for (uint32_t k = 0; k < 4; k++) {
      const uint32_t j = lane + 32 * k;
      bucket[k] = kFull;
      if (j < m) {
        const uint64_t len = batch_off(g, w0 + j + 1) - batch_off(g, w0 + j);
        const uint64_t steps = (len * kCharBytes + 15) >> 4;
        bucket[k] = steps < kBkBuckets - 1 ? static_cast<uint32_t>(steps) : kBkBuckets - 1;
      }
    }
    bk_sts32(scratch + 4 * lane, 0);
    bk_sts32(scratch + 4 * (lane + 32), 0);
    __syncwarp();
#pragma unroll
    for (uint32_t k = 0; k < 4; k++)
      if (bucket[k] != kFull) bk_atom_add(scratch + 4 * bucket[k], 1);
    __syncwarp();
    {  // counters -> first rank of every bucket (descending): lane l owns buckets 2l and 2l + 1
      const uint32_t c0 = bk_lds32(scratch + 8 * lane), c1 = bk_lds32(scratch + 8 * lane + 4);
      uint32_t v = c0 + c1;  // -> inclusive suffix sum over the lanes
#pragma unroll
      for (uint32_t d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_down_sync(kFull, v, d);
        if (lane + d < 32) v += t;
      }
      const uint32_t above = v - (c0 + c1);  // lines in buckets > 2l + 1
      __syncwarp();
      bk_sts32(scratch + 8 * lane + 4, above);
      bk_sts32(scratch + 8 * lane, above + c1);
    }
    __syncwarp();
#pragma unroll
    for (uint32_t k = 0; k < 4; k++)
      if (bucket[k] != kFull) bk_sts8(perm + bk_atom_add(scratch + 4 * bucket[k], 1), lane + 32 * k);
    __syncwarp();

    // ---- 2. four batches of 32 lines of (nearly) equal length
    for (uint32_t q = 0; q * 32 < m; q++) {
      const uint32_t r = q * 32 + lane;
      const bool own = r < m;
      const uint32_t line = w0 + (own ? bk_lds8(perm + r) : 0u);
      int32_t end = 0, from = 0;
      const bool want_rev = l8_stream_lines<CM, CharT>(p, cx, buf0, buf1, lane, own, line, use_from, defer_rev, &end, &from);
      if (defer_rev) {
        const uint32_t rv = __ballot_sync(kFull, want_rev);
        if (rv) enqueue(rv, line, end, from);
      }
    }
    __syncwarp();  // the next window rewrites the scratch
  }
  if (qn) reverse_batch(qn);
}

}  // namespace ndl
