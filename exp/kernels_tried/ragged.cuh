// Ragged batches, third generation: length-bucketed lockstep walkers that read the haystacks from global memory / L2.
// Included by lines8.cuh; runs inside lines8_kernel / linesq_kernel (same table images, same l8_chunk steps) for ragged
// batches of at least kRgMinLines lines.  No tiles: the shared memory the tile buffers would take holds the per-window
// bookkeeping instead.  A CTA takes WINDOWS of kRgWindow consecutive lines; per window:
//
//   0. prefetch   one thread asks for the NEXT window's bytes with bulk L2 prefetches (UBLKPF): the DRAM side of the scan is
//                 a sequential stream, the lanes' own 16-byte loads then hit L2
//   1. histogram  every line's walk length in 16-byte steps (shared-memory counters)
//   2. scatter    line numbers into `perm`, longest bucket first (a counting sort on the walk length, in shared memory)
//   3. walk       a warp takes 32 consecutive entries of `perm` - lines of (nearly) the same length - one per lane, and
//                 walks them in lockstep: each lane reads its own line as aligned 16-byte chunks (one chunk in flight ahead
//                 of the walk), realigns them in registers and steps the automaton exactly as the tile kernels do.  The
//                 lanes of a warp have the same number of steps, so nobody waits for a long neighbour (the tile walk ran
//                 at 51 % lane utilisation on lines of 8..120 bytes, 82 % after sorting and pairing the lines of every
//                 tile pair - at 150 + 50 instructions per step for the bookkeeping), and lines of ANY length take the
//                 same path.  Lines whose find() needs the table-driven reverse pass (indexBackwards,
//                 DFAClassBuilder.java:529-586) are appended to `rev`.
//   4. reverse    the same lockstep walk backwards from end - 1 for the lines in `rev`: only lines that matched, so the
//                 lanes have work (the pooled reverse pass on the staged tiles ran at 26 %).
//
// Results are those of the generated loops of the reference (indexForwards :335-471, indexBackwards :529-614, glue
// :616-667), bit for bit - the per-line arithmetic is l8_run_ragged's.
#pragma once

namespace ndl {

constexpr uint32_t kRgBuckets = 64;
constexpr uint32_t kRgWindow = 8192;  // lines per window (perm / rev entries are 16 bits)
// shared-memory scratch of a CTA (bytes from its base): hist[64] | base[64] | cursor[64] | rev_count | perm u16[W] | rev u16[W]
constexpr uint32_t kRgSmHist = 0, kRgSmBase = 256, kRgSmCursor = 512, kRgSmRevCount = 768, kRgSmPerm = 1024,
                   kRgSmRev = kRgSmPerm + 2 * kRgWindow, kRgSmBytes = kRgSmRev + 2 * kRgWindow;

__device__ __forceinline__ uint4 rg_ldg16(const uint8_t* ptr) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(ptr));
  return v;
}
__device__ __forceinline__ void rg_sts32(uint32_t addr, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t rg_lds32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void rg_sts16(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"(static_cast<uint16_t>(v)) : "memory");
}
__device__ __forceinline__ uint32_t rg_lds16(uint32_t addr) {
  uint16_t v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t rg_atom_add(uint32_t addr, uint32_t v) {
  uint32_t old;
  asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(addr), "r"(v) : "memory");
  return old;
}
__device__ __forceinline__ void rg_cta_sync(uint32_t cta_threads) { asm volatile("bar.sync 1, %0;" ::"r"(cta_threads) : "memory"); }

// walk length in 16-byte steps -> bucket: exact up to 32 steps (512 bytes), then in groups of 16 steps
template <int CM>
__device__ __forceinline__ uint32_t rg_bucket(const BatchParams& g, uint32_t i) {
  const uint64_t len = batch_off(g, i + 1) - batch_off(g, i);
  const uint64_t steps = (len * L8Chars<CM>::kBytes + 15) >> 4;
  if (steps <= 32) return static_cast<uint32_t>(steps);
  const uint64_t b = 32 + ((steps - 32) >> 4);
  return b < kRgBuckets - 1 ? static_cast<uint32_t>(b) : kRgBuckets - 1;
}

template <int CM>
__device__ __forceinline__ void rg_pipeline(const Lines8Params& p, const L8Ctx& cx, const uint32_t sm, const uint32_t lane,
                                            const uint32_t warp_in_cta, const uint32_t usable_warps) {
  using CharT = typename std::conditional<L8Chars<CM>::kBytes == 1, uint8_t, uint16_t>::type;
  constexpr uint32_t kCharBytes = L8Chars<CM>::kBytes;
  constexpr uint32_t kPer = L8Chars<CM>::kPerChunk;
  constexpr uint32_t kFull = 0xffffffffu;
  constexpr uint32_t kStateMask = L8Enc<CM>::kStateMask;
  const BatchParams& g = p.g;
  const uint8_t* const data = static_cast<const uint8_t*>(g.data);
  const uint32_t n = static_cast<uint32_t>(g.n);
  const uint32_t cta_threads = usable_warps * 32;
  const uint32_t tid = warp_in_cta * 32 + lane;  // among the participating threads of the CTA
  const uint32_t mode = static_cast<uint32_t>(g.mode);
  const bool use_from = g.from != nullptr && mode == 2;
  const bool table_rev = mode == 2 && g.reverse_mode == 0 && p.has_bwd != 0;
  // lines per window: at most kRgWindow, and few enough that every CTA gets several windows (long lines: fewer per window)
  uint32_t win_lines = (n / (8u * gridDim.x) + 31u) & ~31u;
  win_lines = win_lines < 256u ? 256u : win_lines > kRgWindow ? kRgWindow : win_lines;
  const uint32_t n_windows = (n + win_lines - 1) / win_lines;

  auto prefetch_window = [&](uint32_t win) {  // the window's bytes -> L2, at most 512 KB of them
    if (win >= n_windows) return;
    const uint32_t l0 = win * win_lines, m = min(win_lines, n - l0);
    const uint64_t b0 = batch_off(g, l0) * kCharBytes, b1 = batch_off(g, l0 + m) * kCharBytes;
    const uint8_t* ptr = data + (b0 & ~static_cast<uint64_t>(15));
    uint64_t left = min(b1 - (b0 & ~static_cast<uint64_t>(15)), static_cast<uint64_t>(512u << 10)) & ~static_cast<uint64_t>(15);
    while (left) {
      const uint32_t piece = static_cast<uint32_t>(min(left, static_cast<uint64_t>(32u << 10)));
      asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(ptr), "r"(piece) : "memory");
      ptr += piece;
      left -= piece;
    }
  };
  if (tid == 0) prefetch_window(blockIdx.x);

  for (uint32_t win = blockIdx.x; win < n_windows; win += gridDim.x) {
    const uint32_t l0 = win * win_lines, m = min(win_lines, n - l0);
    if (tid == 0) prefetch_window(win + gridDim.x);
    // ---- 1. histogram of the walk lengths
    if (tid < kRgBuckets) {
      rg_sts32(sm + kRgSmHist + 4 * tid, 0);
      rg_sts32(sm + kRgSmCursor + 4 * tid, 0);
      if (tid == 0) rg_sts32(sm + kRgSmRevCount, 0);
    }
    rg_cta_sync(cta_threads);
    // (lanes with the same bucket elect a leader: one shared-memory atomic per bucket and warp step, not one per line)
    for (uint32_t j0 = tid - lane; j0 < m; j0 += cta_threads) {
      const uint32_t j = j0 + lane;
      const uint32_t b = j < m ? rg_bucket<CM>(g, l0 + j) : kFull;
      const uint32_t peers = __match_any_sync(kFull, b);
      if (j < m && lane == static_cast<uint32_t>(__ffs(peers) - 1)) rg_atom_add(sm + kRgSmHist + 4 * b, static_cast<uint32_t>(__popc(peers)));
    }
    rg_cta_sync(cta_threads);
    // ---- 2. scatter, longest bucket first
    if (tid < kRgBuckets) {
      uint32_t base = 0;
      for (uint32_t b = kRgBuckets - 1; b > tid; b--) base += rg_lds32(sm + kRgSmHist + 4 * b);
      rg_sts32(sm + kRgSmBase + 4 * tid, base);
    }
    rg_cta_sync(cta_threads);
    for (uint32_t j0 = tid - lane; j0 < m; j0 += cta_threads) {
      const uint32_t j = j0 + lane;
      const uint32_t b = j < m ? rg_bucket<CM>(g, l0 + j) : kFull;
      const uint32_t peers = __match_any_sync(kFull, b);
      const uint32_t leader = __ffs(peers) - 1;
      uint32_t pos = 0;
      if (j < m && lane == leader) pos = rg_atom_add(sm + kRgSmCursor + 4 * b, static_cast<uint32_t>(__popc(peers)));
      pos = __shfl_sync(kFull, pos, leader) + __popc(peers & ((1u << lane) - 1u));
      if (j < m) rg_sts16(sm + kRgSmPerm + 2 * (rg_lds32(sm + kRgSmBase + 4 * b) + pos), j);
    }
    rg_cta_sync(cta_threads);

    // ---- 3. forward walks, 32 lines of one bucket per warp
    const uint32_t n_batches = (m + 31) / 32;
    for (uint32_t bt = warp_in_cta; bt < n_batches; bt += usable_warps) {
      const uint32_t idx = bt * 32 + lane;
      const bool has = idx < m;
      const uint32_t j = has ? rg_lds16(sm + kRgSmPerm + 2 * idx) : 0u;
      const uint32_t i = l0 + j;
      uint64_t o0 = 0;
      uint32_t len = 0;
      int32_t from = 0;
      bool slow = false;
      if (has) {
        o0 = batch_off(g, i);
        const uint64_t l64 = batch_off(g, i + 1) - o0;
        slow = l64 >= (1ull << 31);
        len = slow ? 0u : static_cast<uint32_t>(l64);
        if (use_from && !slow) {
          const int32_t f = g.from[i];
          if (f < 0 || (f != 0 && static_cast<uint32_t>(f) >= len)) {
            slow = true;  // keeps the reference's corner cases: generic walk
            len = 0;
          } else {
            from = f;
            o0 += static_cast<uint32_t>(f);
            len -= static_cast<uint32_t>(f);
          }
        }
      }
      const uint8_t* const ptr = data + o0 * kCharBytes;  // first byte walked
      const uint8_t* const end_ptr = ptr + static_cast<uint64_t>(len) * kCharBytes;
      const uint32_t a = static_cast<uint32_t>(reinterpret_cast<uintptr_t>(ptr)) & 15u;
      const uint8_t* cur = ptr - a;  // aligned chunk of the first byte
      const uint4 zero = make_uint4(0, 0, 0, 0);
      // (a chunk is only read when it holds a byte of the line: nothing outside the haystacks is touched)
      uint4 x = cur < end_ptr ? rg_ldg16(cur) : zero;
      uint4 y = cur + 16 < end_ptr ? rg_ldg16(cur + 16) : zero;
      const L8Align al(a);
      uint32_t e = cx.root, pos = 0;
      int32_t last = g.fwd.root_accepting ? 0 : -1;
      uint32_t tail_bit = g.fwd.root_accepting ? 1u : 0u;
      while (__ballot_sync(kFull, pos < len) != 0) {
        cur += 16;
        const uint4 y_next = (pos < len && cur + 16 < end_ptr) ? rg_ldg16(cur + 16) : zero;  // one chunk ahead of the walk
        if (pos < len) {
          const uint4 wv = al.apply(x, y);
          uint32_t mask = 0;
          l8_chunk<CM>(wv, p.q, cx, e, mask);
          const uint32_t valid = min(kPer, len - pos);
          mask >>= (kPer - valid);  // drop the accept bits of chars past the end of the line
          const int32_t cand = static_cast<int32_t>(pos + valid + 1) - __ffs(mask);
          last = mask ? cand : last;
          tail_bit = mask & 1u;
          pos += kPer;
          // a dead automaton stays dead (and never accepts); containedIn() has its answer at the first accept
          if ((e & kStateMask) == cx.fwd_dead || (mode == 1 && last != -1)) pos = len;
        }
        x = y;
        y = y_next;
      }
      // results (l8_finish's cases, without a staged tile)
      bool want_rev = false;
      if (has) {
        if (slow) {
          l8_slow_line<CharT>(g, i);
        } else if (mode == 0) {
          bool mt = tail_bit != 0;
          if (g.min_length > 4 && static_cast<uint32_t>(g.min_length) > len) mt = false;  // DFAMethodComponents.java:75-93
          if (g.max_length != -1 && len > static_cast<uint32_t>(g.max_length)) mt = false;
          g.matched[i] = mt;
        } else if (mode == 1) {
          g.matched[i] = last != -1;
        } else {
          int32_t st = -1, end = last;
          if (last != -1) {
            end = last + from;
            if (g.reverse_mode == 2) {  // start = end - minLength (:640-646)
              st = end - g.min_length;
            } else if (table_rev) {     // phase 4
              want_rev = true;
            } else {                    // single-char reverse scan, or no resident BACKWARDS rows: the generic loop
              st = static_cast<int32_t>(dev_index_backwards<CharT>(g, static_cast<const CharT*>(g.data) + batch_off(g, i), end - 1, from, 0x7fffffff));
            }
          }
          g.matched[i] = last != -1;
          g.start[i] = st;
          g.end[i] = end;
        }
      }
      const uint32_t rv = __ballot_sync(kFull, want_rev);
      if (rv) {
        uint32_t base = 0;
        if (lane == 0) base = rg_atom_add(sm + kRgSmRevCount, static_cast<uint32_t>(__popc(rv)));
        base = __shfl_sync(kFull, base, 0);
        // the end offset travels with the entry's line number?  No: 16-bit entries; the reverse walk re-reads g.end (own CTA's write)
        if (want_rev) rg_sts16(sm + kRgSmRev + 2 * (base + __popc(rv & ((1u << lane) - 1u))), j);
      }
    }
    if (table_rev) {
      __threadfence_block();
      rg_cta_sync(cta_threads);
      // ---- 4. reverse walks: indexBackwards(end - 1, from) for the lines of this window that matched
      const uint32_t n_rev = rg_lds32(sm + kRgSmRevCount);
      const uint32_t n_rev_batches = (n_rev + 31) / 32;
      for (uint32_t bt = warp_in_cta; bt < n_rev_batches; bt += usable_warps) {
        const uint32_t idx = bt * 32 + lane;
        const bool has = idx < n_rev;
        const uint32_t i = l0 + (has ? rg_lds16(sm + kRgSmRev + 2 * idx) : 0u);
        int32_t from = 0;
        uint32_t total = 0;  // chars to walk: [from, end)
        uint64_t o0 = 0;
        if (has) {
          o0 = batch_off(g, i);
          from = use_from ? g.from[i] : 0;
          total = static_cast<uint32_t>(__ldcg(g.end + i) - from);
        }
        const uint8_t* const lo_ptr = data + (o0 + static_cast<uint32_t>(from)) * kCharBytes;  // first byte that may be walked
        const uint8_t* const h_ptr = lo_ptr + static_cast<uint64_t>(total) * kCharBytes;        // one past the last
        // window k = bytes [h - 16 (k + 1), h - 16 k): chunk x holds its first byte, chunk y the rest
        const uint8_t* const w0 = h_ptr - 16;
        const uint32_t a = static_cast<uint32_t>(reinterpret_cast<uintptr_t>(w0)) & 15u;
        const uint8_t* xp = w0 - a;
        const uint4 zero = make_uint4(0, 0, 0, 0);
        // (a chunk is only read when it holds a byte of [lo_ptr, h_ptr))
        uint4 y = (total != 0 && a != 0) ? rg_ldg16(xp + 16) : zero;
        uint4 x = (total != 0 && xp + 16 > lo_ptr) ? rg_ldg16(xp) : zero;
        const L8Align al(a);
        uint32_t e = cx.bwd_root, pos = 0;
        int32_t w = g.bwd.root_accepting ? static_cast<int32_t>(total) : -1;  // lastMatch = lowerBound when the root accepts (:543-547)
        while (__ballot_sync(kFull, pos < total) != 0) {
          xp -= 16;
          const uint4 x_next = (pos + kPer < total && xp + 16 > lo_ptr) ? rg_ldg16(xp) : zero;  // one chunk ahead of the walk
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
        if (has) g.start[i] = w == -1 ? 0x7fffffff : static_cast<int32_t>(total) - w + from;
      }
    }
    rg_cta_sync(cta_threads);  // the next window reuses the shared-memory lists
  }
}

}  // namespace ndl
