// long8: find() over ONE long byte haystack (BASELINE config 4: a single 8 GiB string), chunk-parallel.
//
// A DFA walk is sequential, but an unanchored search DFA forgets: while no match has been seen its state
// is (for most patterns) a function of the last few chars only, because every non-accepting state carries
// the restart thread (NFAToDFACompiler.java:70-72, 108-112).  So the haystack is cut into 64-byte
// segments, one per lane, 32 per warp tile (the same swizzled shared-memory tiles and pair tables as
// lines8), and every lane
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
__global__ void seq_walk_kernel(DevTable t, const CharT* s, int64_t p0, int64_t p1, int32_t state0, int64_t count_from,
                                int64_t last_init, SeqResult* out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const int dead = t.n_states;
  int state = state0;
  int64_t last = last_init;
  int64_t i = p0;
  for (; i < p1; i++) {
    state = dev_step(t, state, s[i]);
    if (state == dead) break;
    if (i >= count_from && __ldg(t.accept + state)) last = i + 1;
  }
  out->pos = i;
  out->last = last;
  out->state = state;
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

struct Long8Params {
  const uint8_t* data;     // 2048-byte aligned start of tile 0
  uint64_t n_tiles;        // full 2 KB tiles
  const uint8_t* image;    // lines8 image of the FORWARDS table
  uint32_t trans_bytes;
  uint32_t root_entry;
  uint32_t row_bytes;      // kCmBytes1 layout
  uint32_t entry0;         // exact entry of tile 0, lane 0 (same encoding as the table entries, flags clear)
  uint32_t* seam_guess;    // [n_tiles] lane 0's guessed entry
  uint32_t* seam_exit;     // [n_tiles] lane 31's exit
  unsigned long long* first_seg;   // atomicMin: first segment (global index) that saw an accepting state
  unsigned long long* first_bad;   // atomicMin: first segment whose in-warp check failed
};

template <int CM>
__global__ void __launch_bounds__(kL8Threads, 1) long8_kernel(const Long8Params p) {
  const uint32_t tid = threadIdx.x;
  const uint32_t lane = tid & 31, warp = tid >> 5;
  L8Setup su(CM == kCmBytes1, warp);
  const uint32_t buf0 = su.buf0, buf1 = su.buf1;
  if (tid == 0) {
    mbar_init(kL8AbsBar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  if (!su.layout_ok) {  // cannot happen with the launch configuration used; report every tile as unverified
    if (tid == 0 && blockIdx.x == 0) atomicMin(p.first_bad, 0ull);
    return;
  }
  if (tid == 0) {
    mbar_expect_tx(kL8AbsBar, su.cmap_bytes + p.trans_bytes);
    tma_bulk_g2s(kL8AbsCmap, p.image, su.cmap_bytes, kL8AbsBar);
    tma_bulk_g2s(su.abs_trans, p.image + su.cmap_bytes, p.trans_bytes, kL8AbsBar);
  }
  constexpr int LOG2CPL = 2;  // 64-byte segments
  uint32_t dst_off[4];
#pragma unroll
  for (uint32_t k = 0; k < 4; k++) {
    const uint32_t c = lane + 32 * k;
    dst_off[k] = l8_slot(c >> LOG2CPL, c & 3, LOG2CPL) << 4;
  }
  L8Ctx cx;
  cx.sel_a = 0x00010000u | (lane * 4);
  cx.sel_b = cx.sel_a | 0x80u;
  cx.page1 = cx.page3 = cx.ua = cx.ub = cx.xa = cx.xb = 0;
  cx.row_bytes = p.row_bytes;
  auto stage = [&](uint64_t t, uint32_t buf) {
    const uint8_t* src = p.data + t * 2048 + lane * 16;
#pragma unroll
    for (uint32_t k = 0; k < 4; k++) cp_async16(buf + dst_off[k], src + 512 * k);
    cp_async_commit();
  };
  const uint64_t n_warps = static_cast<uint64_t>(gridDim.x) * su.usable_warps;
  uint64_t t = static_cast<uint64_t>(blockIdx.x) * su.usable_warps + warp;
  uint32_t cur = buf0, nxt = buf1;
  mbar_wait(kL8AbsBar, 0);
  if (!su.warp_ok) return;  // no pair of tile buffers for this warp in this layout
  if (t < p.n_tiles) stage(t, cur);

  for (; t < p.n_tiles; t += n_warps) {
    if (t + n_warps < p.n_tiles) stage(t + n_warps, nxt);
    else cp_async_commit();
    // lane 0's warm-up bytes are the last 16 bytes of the previous tile: fetch them while the copy lands
    uint4 pre = make_uint4(0, 0, 0, 0);
    if (lane == 0 && t > 0) pre = *reinterpret_cast<const uint4*>(p.data + t * 2048 - 16);
    cp_async_wait<1>();
    __syncwarp();
    // a match in an earlier segment makes this tile irrelevant
    const unsigned long long seg0 = t * 32ull;
    if (*reinterpret_cast<volatile unsigned long long*>(p.first_seg) >= seg0) {
      // 1. guess: 16 bytes before the segment, from the root
      if (lane != 0) pre = lds_data16(cur + (l8_slot(lane - 1, 3, LOG2CPL) << 4));
      uint32_t e = p.root_entry, mask = 0;
      l8_word<CM>(pre.x, cx, e, mask);
      l8_word<CM>(pre.y, cx, e, mask);
      l8_word<CM>(pre.z, cx, e, mask);
      l8_word<CM>(pre.w, cx, e, mask);
      if (lane == 0 && t == 0) e = p.entry0;  // the head was walked exactly
      constexpr uint32_t kStateMask = CM == kCmBytes1 ? 0x7fffu : kL8FlagMask;
      const uint32_t guess = e & kStateMask;
      // 2. the segment itself
      uint32_t any = 0;
#pragma unroll
      for (uint32_t c = 0; c < 4; c++) {
        const uint4 w = lds_data16(cur + (l8_slot(lane, c, LOG2CPL) << 4));
        mask = 0;
        l8_word<CM>(w.x, cx, e, mask);
        l8_word<CM>(w.y, cx, e, mask);
        l8_word<CM>(w.z, cx, e, mask);
        l8_word<CM>(w.w, cx, e, mask);
        any |= mask;
      }
      const uint32_t exit_state = e & kStateMask;
      // 3. in-warp check + seams
      const uint32_t prev_exit = __shfl_up_sync(0xffffffffu, exit_state, 1);
      const bool bad = lane != 0 && prev_exit != guess;
      const uint32_t bad_lanes = __ballot_sync(0xffffffffu, bad);
      const uint32_t acc_lanes = __ballot_sync(0xffffffffu, any != 0);
      if (lane == 0) {
        p.seam_guess[t] = guess;
        if (bad_lanes) atomicMin(p.first_bad, seg0 + (__ffs(bad_lanes) - 1));
        if (acc_lanes) atomicMin(p.first_seg, seg0 + (__ffs(acc_lanes) - 1));
      }
      if (lane == 31) p.seam_exit[t] = exit_state;
    } else if (lane == 0) {
      p.seam_guess[t] = 0xffffffffu;  // skipped: not part of the verified prefix
      p.seam_exit[t] = 0xfffffffeu;
    }
    __syncwarp();
    const uint32_t tmp = cur;
    cur = nxt;
    nxt = tmp;
  }
  cp_async_wait<0>();
}

// Seams between tiles: lane 0 of tile t must have guessed the exit of lane 31 of tile t-1.
__global__ void long8_seam_kernel(const uint32_t* guess, const uint32_t* exit_state, uint64_t n_tiles, unsigned long long* first_bad) {
  const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
  for (uint64_t t = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x + 1; t < n_tiles; t += stride)
    if (guess[t] != exit_state[t - 1]) atomicMin(first_bad, t * 32ull);
}

}  // namespace ndl
