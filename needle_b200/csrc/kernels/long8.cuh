// long8: find() over ONE long haystack (BASELINE config 4: a single 8 GiB string; bytes, or UTF-16 code units), chunk-parallel.
//
// A DFA walk is sequential, but an unanchored search DFA forgets: while no match has been seen its state
// is (for most patterns) a function of the last few chars only, because every non-accepting state carries
// the restart thread (NFAToDFACompiler.java:70-72, 108-112).  So the haystack is cut into 256-byte
// segments, one per lane, 32 per warp tile (staged as 64-byte pieces through the same swizzled shared-memory
// tiles, walked with the same table images as lines8 / linesq), and every lane
//   1. guesses its entry state by walking the 16 bytes before its segment from the root,
//   2. walks its segment from that guess, recording its exit state and whether it saw an accepting state,
//   3. checks, by warp shuffle, that the exit state of the lane before it equals its own guess
//      (lane 0 / lane 31 leave their guess / exit in global memory; a second tiny kernel checks the seams).
// If every check up to the first segment that saw an accept holds, then by induction from the exact head
// every guess was the true state, and the first accepting segment is the true one; a single thread then
// re-walks from that segment's start to find the exact end (and start) of the match.  If a check fails
// (patterns that remember far back, e.g. `a.*c`), the call falls back to a plain sequential walk - correct,
// slow.  The warp ballot picks the first accepting lane; an atomicMin publishes the first segment.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "generic.cuh"
#include "lines8.cuh"

namespace ndl {

struct SeqResult {
  int64_t pos;         // where the walk stopped (first unread index)
  int64_t last;        // index after the last accepting step, or -1
  int32_t state;       // state at `pos` (n_states == DEAD)
  int32_t pad;
};

// One thread walks [p0, p1) from `state0` with the generic tables; stops at DEAD.  `count_from`: accepting
// steps at indices < count_from are ignored (warm-up).  last_init seeds `last`.
template <typename CharT>
__device__ __forceinline__ SeqResult dev_seq_walk(const DevTable& t, const CharT* s, int64_t p0, int64_t p1, int32_t state0, int64_t count_from,
                                                  int64_t last_init) {
  const int dead = t.n_states;
  int state = state0;
  int64_t last = last_init;
  int64_t i = p0;
  for (; i < p1; i++) {
    state = dev_step(t, state, s[i]);
    if (state == dead) break;
    if (i >= count_from && __ldg(t.accept + state)) last = i + 1;
  }
  SeqResult r;
  r.pos = i;
  r.last = last;
  r.state = state;
  r.pad = 0;
  return r;
}

template <typename CharT>
__global__ void seq_walk_kernel(DevTable t, const CharT* s, int64_t p0, int64_t p1, int32_t state0, int64_t count_from,
                                int64_t last_init, SeqResult* out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  *out = dev_seq_walk<CharT>(t, s, p0, p1, state0, count_from, last_init);
}

// Backwards: indexBackwards(index, lower) (DFAClassBuilder.java:529-614), one thread.
template <typename CharT>
__global__ void seq_back_kernel(BatchParams g, const CharT* s, int64_t index, int64_t lower, int64_t int_max, int64_t* out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  *out = dev_index_backwards<CharT>(g, s, index, lower, int_max);
}

// The same loop from an arbitrary BACKWARDS state (one rank's part of a reverse pass that crosses chunk
// boundaries): scans s[lower, index] downwards; out[0] = smallest accepting index or last_init, out[1] = exit state
// (n_states = DEAD; single-char form: 0 = not found yet, n_states = found).
template <typename CharT>
__global__ void seq_back_from_kernel(BatchParams g, const CharT* s, int64_t index, int64_t lower, int32_t state0, int64_t last_init,
                                     int64_t* out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const int dead = g.bwd.n_states;
  int64_t last = last_init;
  int state = state0;
  if (g.reverse_mode == 1) {
    state = 0;
    for (; index >= lower; index--)
      if (static_cast<int>(s[index]) == g.reverse_char) {
        last = index;
        state = dead;
        break;
      }
  } else {
    for (; state != dead && index >= lower; index--) {
      state = dev_step(g.bwd, state, s[index]);
      if (state == dead) break;
      if (__ldg(g.bwd.accept + state)) last = index;
    }
  }
  out[0] = last;
  out[1] = state;
}

constexpr uint32_t kLongSeg = 256;  // bytes per segment = per lane and tile

struct Long8Params {
  const uint8_t* data;     // 16-byte aligned start of segment 0
  uint64_t n_segs;         // 256-byte segments; 32 per tile, the last tile may be partial
  const uint8_t* image;    // lines8 / linesq image of the FORWARDS table
  uint32_t trans_bytes;
  uint32_t root_entry;
  uint32_t row_bytes;      // kCmBytes1 layout
  uint32_t entry0;         // exact entry of segment 0 in the canonical encoding (see canon())
  uint32_t ua, ub, xa, xb; // UTF-16 class-map modes (kCmHi / kCmMixed), as in Lines8Params
  int mixed_page, replicated;
  SwarDev q;               // SWAR modes
  uint32_t* seam_guess;    // [n_tiles] lane 0's guessed entry (canonical)
  uint32_t* seam_exit;     // [n_tiles] exit of the tile's last segment (canonical)
  uint32_t* seam_acc;      // [n_tiles] of the tile's first accepting segment: (canonical state at the start of its first
                           // accepting 64-byte piece) << 2 | piece - where the exact re-walk starts
  unsigned long long* first_seg;   // atomicMin: first segment (global index) that saw an accepting state
  unsigned long long* first_bad;   // atomicMin: first segment whose in-warp check failed
  // Refinement passes (a pattern that remembers further back than the 16-byte warm-up, e.g. `q[a-z ]*7`): instead of
  // guessing from the warm-up, segment s enters in entry_in[s - 1] - the exit of segment s - 1 in the PREVIOUS pass - and
  // every segment leaves its exit in exit_out.  The same checks decide whether the pass is exact; each pass extends the
  // distance the automaton may remember by one segment (256 bytes).  Both NULL: the plain guessing pass.
  const uint32_t* entry_in;
  uint32_t* exit_out;
};

// A lane owns 256 contiguous bytes of a tile and walks them as four 64-byte pieces through the usual swizzled
// 2 KB warp buffers (double buffered across pieces and tiles), so only 16 of every 272 bytes walked are warm-up.
// States are compared across lanes in a canonical encoding: the table entry without flags and - in the SWAR
// layouts, where every lane addresses its own table copy - without the lane's copy offset.
template <int CM>
__global__ void __launch_bounds__(cm_is_swar(CM) ? kQThreads : kL8Threads, 1) long8_kernel(const Long8Params p) {
  constexpr bool kSwar = cm_is_swar(CM);
  constexpr uint32_t kBlockWarps = kSwar ? kQWarps : kL8Warps;
  constexpr uint32_t kStateMask = L8Enc<CM>::kStateMask;
  const uint32_t tid = threadIdx.x;
  const uint32_t lane = tid & 31, warp = tid >> 5;
  bool layout_ok;
  uint32_t buf0, buf1, usable_warps;
  if constexpr (kSwar) {
    extern __shared__ __align__(128) uint8_t l8_dyn_smem[];
    const uint32_t base = static_cast<uint32_t>(__cvta_generic_to_shared(l8_dyn_smem));
    layout_ok = base <= kQAbsTrans;
    const uint32_t tiles_lo = (kQAbsTrans + p.trans_bytes + 127u) & ~127u;
    usable_warps = min(kBlockWarps, (kL8AbsBar - tiles_lo) / kL8WarpBuf / 2);
    buf0 = tiles_lo + 2 * warp * kL8WarpBuf;
    buf1 = buf0 + kL8WarpBuf;
  } else {
    L8Setup su(CM == kCmBytes1, warp);
    layout_ok = su.layout_ok;
    usable_warps = su.usable_warps;
    buf0 = su.buf0;
    buf1 = su.buf1;
  }
  if (tid == 0) {
    mbar_init(kL8AbsBar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  if (!layout_ok) {  // cannot happen with the launch configuration used; report every tile as unverified
    if (tid == 0 && blockIdx.x == 0) atomicMin(p.first_bad, 0ull);
    return;
  }
  if (tid == 0) {
    if constexpr (kSwar) {
      mbar_expect_tx(kL8AbsBar, p.trans_bytes);
      for (uint32_t off = 0; off < p.trans_bytes; off += 0x8000u)
        tma_bulk_g2s(kQAbsTrans + off, p.image + off, min(0x8000u, p.trans_bytes - off), kL8AbsBar);
    } else {
      const uint32_t cmap_bytes = CM == kCmBytes1 ? kS1CmapBytes : kL8CmapBytes;
      mbar_expect_tx(kL8AbsBar, cmap_bytes + p.trans_bytes);
      tma_bulk_g2s(kL8AbsCmap, p.image, cmap_bytes, kL8AbsBar);
      tma_bulk_g2s(CM == kCmBytes1 ? kS1AbsTrans : kL8AbsTrans, p.image + cmap_bytes, p.trans_bytes, kL8AbsBar);
    }
  }
  constexpr int LOG2CPL = 2;  // 64-byte pieces
  uint32_t dst_off[4];
#pragma unroll
  for (uint32_t k = 0; k < 4; k++) {
    const uint32_t c = lane + 32 * k;
    dst_off[k] = l8_slot(c >> LOG2CPL, c & 3, LOG2CPL) << 4;
  }
  const uint32_t lane_off = kSwar ? (lane & p.q.copy_mask) * p.q.copy_bytes : 0u;
  L8Ctx cx;
  cx.sel_a = 0x00010000u | (lane * 4);
  cx.sel_b = cx.sel_a | 0x80u;
  cx.page1 = static_cast<uint32_t>(p.mixed_page) << 8;
  cx.page3 = static_cast<uint32_t>(p.mixed_page) << 24;
  cx.ua = p.ua;
  cx.ub = p.ub + (p.replicated == 32 ? lane * 4 : 0);
  cx.xa = p.xa;
  cx.xb = p.xb + (p.replicated == 32 ? lane * 4 : 0);
  cx.row_bytes = p.row_bytes;
  cx.root = p.root_entry + lane_off;
  cx.bwd_root = cx.bwd_dead = 0;
  const uint64_t n_tiles = (p.n_segs + 31) / 32;
  // piece j of tile t: lane's copies are 16-byte chunks c = lane + 32 k of the piece, chunk c = part (c & 3) of segment (c >> 2)
  auto stage = [&](uint64_t t, uint32_t j, uint32_t buf) {
    const uint64_t seg0 = t * 32;
    const uint32_t segs_here = static_cast<uint32_t>(min(static_cast<uint64_t>(32), p.n_segs - seg0));
    const uint8_t* src = p.data + seg0 * kLongSeg + j * 64 + (lane >> 2) * kLongSeg + (lane & 3) * 16;
#pragma unroll
    for (uint32_t k = 0; k < 4; k++)
      if ((lane >> 2) + 8 * k < segs_here) cp_async16(buf + dst_off[k], src + 8 * k * kLongSeg);
    cp_async_commit();
  };
  const uint64_t n_warps = static_cast<uint64_t>(gridDim.x) * usable_warps;
  uint64_t t = static_cast<uint64_t>(blockIdx.x) * usable_warps + warp;
  uint32_t cur = buf0, nxt = buf1;
  mbar_wait(kL8AbsBar, 0);
  if (warp >= usable_warps) return;  // no pair of tile buffers for this warp in this layout
  if (t < n_tiles) stage(t, 0, cur);

  for (; t < n_tiles; t += n_warps) {
    const uint64_t seg0 = t * 32;
    const uint32_t segs_here = static_cast<uint32_t>(min(static_cast<uint64_t>(32), p.n_segs - seg0));
    const bool act = lane < segs_here;
    // a match in an earlier segment makes this tile - and every later one of this warp - irrelevant
    // (not while refining: an accept seen from a wrong entry state must not hide the segments behind it)
    if (p.exit_out == nullptr && *reinterpret_cast<volatile unsigned long long*>(p.first_seg) < seg0) {
      if (lane == 0)
        for (uint64_t u = t; u < n_tiles; u += n_warps) {
          p.seam_guess[u] = 0xffffffffu;  // skipped: not part of the verified prefix
          p.seam_exit[u] = 0xfffffffeu;
        }
      break;
    }
    // warm-up bytes: the 16 bytes before the lane's segment, straight from global memory
    uint4 pre = make_uint4(0, 0, 0, 0);
    uint32_t entry_prev = 0;
    if (act && (seg0 + lane) != 0) {
      if (p.entry_in) entry_prev = p.entry_in[seg0 + lane - 1];
      else pre = *reinterpret_cast<const uint4*>(p.data + (seg0 + lane) * kLongSeg - 16);
    }
    uint32_t e = cx.root, mask = 0, any = 0, guess = 0, acc_at = 0;
#pragma unroll 1
    for (uint32_t j = 0; j < 4; j++) {
      if (j < 3) stage(t, j + 1, nxt);
      else if (t + n_warps < n_tiles) stage(t + n_warps, 0, nxt);
      else cp_async_commit();
      cp_async_wait<1>();
      __syncwarp();
      if (act) {
        if (j == 0) {  // 1. guess the entry state
          if (p.entry_in) e = entry_prev + lane_off;  // (canonical encoding + this lane's table copy)
          else l8_chunk<CM>(pre, p.q, cx, e, mask);
          if (seg0 + lane == 0) e = p.entry0 + lane_off;  // the head was walked exactly
          guess = (e & kStateMask) - lane_off;
        }
        const uint32_t e_piece = (e & kStateMask) - lane_off;
        uint32_t any_piece = 0;
#pragma unroll
        for (uint32_t c = 0; c < 4; c++) {  // 2. the piece itself
          const uint4 w = lds_data16(cur + (l8_slot(lane, c, LOG2CPL) << 4));
          mask = 0;
          l8_chunk<CM>(w, p.q, cx, e, mask);
          any_piece |= mask;
        }
        if (any == 0 && any_piece != 0) acc_at = e_piece << 2 | j;
        any |= any_piece;
      }
      __syncwarp();
      const uint32_t tmp = cur;
      cur = nxt;
      nxt = tmp;
    }
    // 3. in-warp check + seams
    const uint32_t exit_state = (e & kStateMask) - lane_off;
    const uint32_t prev_exit = __shfl_up_sync(0xffffffffu, exit_state, 1);
    const bool bad = act && lane != 0 && prev_exit != guess;
    const uint32_t bad_lanes = __ballot_sync(0xffffffffu, bad);
    const uint32_t acc_lanes = __ballot_sync(0xffffffffu, act && any != 0);
    if (lane == 0) {
      p.seam_guess[t] = guess;
      if (bad_lanes) atomicMin(p.first_bad, static_cast<unsigned long long>(seg0 + (__ffs(bad_lanes) - 1)));
      if (acc_lanes) atomicMin(p.first_seg, static_cast<unsigned long long>(seg0 + (__ffs(acc_lanes) - 1)));
    }
    if (lane == segs_here - 1) p.seam_exit[t] = exit_state;
    if (p.exit_out && act) p.exit_out[seg0 + lane] = exit_state;
    if (acc_lanes && lane == static_cast<uint32_t>(__ffs(acc_lanes) - 1)) p.seam_acc[t] = acc_at;
  }
  cp_async_wait<0>();
}

// What ndl_find_long does once the segments are walked and the seams checked, in one launch (one host round
// trip instead of three): decide from first_seg / first_bad, re-walk exactly from the first accepting segment (its
// entry state re-derived from the 16 bytes before it - the verified guess), or continue from the last segment's
// exit into the tail.  status 1: a guess was wrong before any match - the host falls back to the sequential walk.
struct Long8Epilogue {
  SeqResult r;
  int32_t status;
  int32_t pad;
};
struct Long8Decode {  // canonical table entry -> state id (see long8_kernel)
  int32_t kind;       // 0 pair table (row_bytes), 1 stride-1 table (row index), 2 SWAR image
  uint32_t row_bytes, w_rows, entry_bytes;
};
#ifdef NDL_MAIN_TU
template <typename CharT>
__global__ void long8_epilogue_kernel(DevTable t, const CharT* s, int64_t head_end, int64_t n, uint64_t n_segs, int32_t head_state,
                                      Long8Decode dec, const uint32_t* seam_exit, const uint32_t* seam_acc, const unsigned long long* first_seg,
                                      const unsigned long long* first_bad, Long8Epilogue* out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  constexpr int64_t kCharBytes = sizeof(CharT);  // positions are in chars, segments in bytes
  const unsigned long long kNone = ~0ull;
  const unsigned long long fs = *first_seg, fb = *first_bad;
  Long8Epilogue o;
  o.status = 0;
  o.pad = 0;
  o.r.pos = head_end;
  o.r.last = -1;
  o.r.state = head_state;
  o.r.pad = 0;
  auto decode = [&](uint32_t canon) -> int32_t {
    if (dec.kind == 2) {
      const uint32_t off = canon - kQAbsTrans;
      return static_cast<int32_t>((off / 128u) * dec.w_rows + (off % 128u) / dec.entry_bytes);
    }
    return static_cast<int32_t>(dec.kind == 1 ? canon : canon / dec.row_bytes);
  };
  if (fb != kNone && (fs == kNone || fb <= fs)) {
    o.status = 1;
  } else if (fs != kNone) {
    // the first accepting segment published the (verified) state at the start of its first accepting 64-byte piece
    const uint32_t acc = seam_acc[fs / 32];
    const int64_t pos = head_end + (static_cast<int64_t>(fs) * kLongSeg + static_cast<int64_t>(acc & 3u) * 64) / kCharBytes;
    o.r = dev_seq_walk<CharT>(t, s, pos, n, decode(acc >> 2), pos, -1);
  } else {
    const uint64_t n_tiles = (n_segs + 31) / 32;
    const int32_t state = decode(seam_exit[n_tiles - 1]);
    const int64_t pos = head_end + static_cast<int64_t>(n_segs) * kLongSeg / kCharBytes;
    if (pos < n) {
      o.r = dev_seq_walk<CharT>(t, s, pos, n, state, pos, -1);
    } else {
      o.r.pos = pos;
      o.r.state = state;
    }
  }
  *out = o;
}

#endif  // NDL_MAIN_TU

// Seams between tiles: lane 0 of tile t must have guessed the exit of lane 31 of tile t-1.
#ifdef NDL_MAIN_TU
__global__ void long8_seam_kernel(const uint32_t* guess, const uint32_t* exit_state, uint64_t n_tiles, unsigned long long* first_bad) {
  const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
  for (uint64_t t = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x + 1; t < n_tiles; t += stride)
    if (guess[t] != exit_state[t - 1]) atomicMin(first_bad, t * 32ull);
}

#endif  // NDL_MAIN_TU

}  // namespace ndl
