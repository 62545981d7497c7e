// lines8: the tuned batch kernel for byte haystacks (char_width == 1), one haystack ("line") per lane.
//
// Per char the generated Java loop does two dependent array loads (BYTE_CLASSES[c], then
// STATES[class + state*stride]) plus bookkeeping (DFAClassBuilder.java:438-465).  Here the same two
// lookups are shared-memory loads laid out so that a warp's 32 lanes never collide on a bank, and the
// bookkeeping is folded into the table encoding, so one char costs PRMT + LDS + IADD + LDS + SHF:
//
//   Shared memory is addressed ABSOLUTELY (shared-window addresses), with a fixed map:
//     [base, 0x10000)      tile buffer A (haystack bytes)
//     [0x10000, 0x20000)   cmap: 256 slots x 256 B.  Slot b holds, per lane, CM[b][lane] (int32) for the
//                          forward table in its first 128 B and for the BACKWARDS table in its second
//                          128 B.  Lane l only reads word l of a slot -> bank l -> conflict free, and
//                          because the region starts at 0x10000 the address of that word is formed by
//                          ONE byte-permute: bytes {lane*4, haystack byte, 0x01, 0x00}.
//     [0x20000, 0x2FC00)   tile buffer B
//     [0x2FC00, 0x387C0)   trans: rows x cols entries E (int32), replicated per lane (entry (row, col)
//                          is 128 B, lane l reads word l) when that fits, else a single copy.
//                          Rows of accepting states lie BELOW a midpoint, all others above it; E is
//                          the byte offset of the target row from the midpoint (+ lane*4 when
//                          replicated) and CM = absolute midpoint address + column offset, so
//                              next entry address = E + CM[byte][lane]      (one IADD)
//                              accepting(target)  = E < 0                   (sign bit; one SHF shifts it
//                                                                            into a per-line bit mask)
//                          DEAD is an ordinary absorbing row (device_image.h): no per-char branch.
//     [0x387C0, ...)       mbarrier for the TMA bulk copies that bring cmap + trans in (UBLKCP)
//
//   Haystack bytes: coalesced 16-byte cp.async (LDGSTS) global->shared copies of a tile of lines, double
//   buffered against the walk, stored with an XOR swizzle of the 16-byte chunk index so that the
//   per-lane 16-byte reads of 32 consecutive lines are conflict free as well.
//
// The fast path needs every line of a tile to have the same length L in {16, 32, 64, 128} and 16-byte
// alignment; any other tile is walked with the generic per-thread code (generic.cuh) - same results.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstring>
#include <vector>

#include "../device_image.h"
#include "generic.cuh"

namespace ndl {

constexpr int kL8Threads = 1024;
constexpr uint32_t kL8AbsCmap = 0x10000;
constexpr uint32_t kL8CmapBytes = 0x10000;
constexpr uint32_t kL8AbsBufB = 0x20000;
constexpr uint32_t kL8BufBytes = 0xFC00;  // 63 KB
constexpr uint32_t kL8AbsTrans = 0x2FC00;
constexpr uint32_t kL8AbsBar = 0x387C0;
constexpr uint32_t kL8MaxTransBytes = kL8AbsBar - kL8AbsTrans;  // 35776
constexpr uint32_t kL8AbsEnd = kL8AbsBar + 16;
constexpr uint32_t kL8DynSmem = kL8AbsEnd;  // covers the map for any dynamic base in [0, 0x400]

struct Lines8Image {
  bool available = false;   // false: table too large for this kernel -> generic path
  int replicated = 0;       // 32 or 1
  int32_t mid_off = 0;      // byte offset of the midpoint inside the trans image
  int32_t trans_bytes = 0;
  int cols = 0;
  std::vector<int32_t> row_off;   // per state: byte offset relative to mid
  std::vector<uint16_t> cmap;     // 256: byte -> column
  std::vector<uint16_t> trans;    // (n_states + 1) * cols -> next state (copy of the device automaton)
};

// Device image for one mode: [cmap 64 KB][trans], plus what the kernel needs to start a walk.
struct Lines8Blob {
  uint8_t* dev = nullptr;
  uint32_t trans_bytes = 0;  // multiple of 16
  int32_t fwd_root = 0, bwd_root = 0;
  int fwd_repl = 0, bwd_repl = 0;
  bool has_bwd = false;
  bool ok = false;
};

inline void lines8_build(const HostDeviceTable& t, Lines8Image& out) {
  out = Lines8Image();
  const int rows = t.n_states + 1, cols = t.n_classes;
  out.cols = cols;
  const long bytes32 = static_cast<long>(rows) * cols * 128;
  const long bytes1 = static_cast<long>(rows) * cols * 4;
  if (bytes32 <= static_cast<long>(kL8MaxTransBytes) / 2)
    out.replicated = 32;
  else if (bytes1 <= static_cast<long>(kL8MaxTransBytes) / 2)
    out.replicated = 1;
  else
    return;
  const int row_bytes = cols * 4 * out.replicated;
  int n_acc = 0;
  for (int s = 0; s < rows; s++) n_acc += t.accept[s] ? 1 : 0;
  out.mid_off = n_acc * row_bytes;
  out.trans_bytes = rows * row_bytes;
  out.row_off.resize(rows);
  int below = 0, above = 0;
  for (int s = 0; s < rows; s++) out.row_off[s] = t.accept[s] ? -(++below) * row_bytes : (above++) * row_bytes;
  out.cmap.assign(t.cmap.begin(), t.cmap.begin() + 256);
  out.trans = t.trans;
  out.available = true;
}

// Lay the forward (and optionally backward) image out exactly as it will sit in shared memory.
inline bool lines8_layout(const Lines8Image& fwd, const Lines8Image* bwd, std::vector<uint8_t>& img, Lines8Blob& meta) {
  if (!fwd.available || (bwd && !bwd->available)) return false;
  const uint32_t fwd_bytes = (fwd.trans_bytes + 15) & ~15;
  const uint32_t bwd_bytes = bwd ? ((bwd->trans_bytes + 15) & ~15) : 0;
  if (fwd_bytes + bwd_bytes > kL8MaxTransBytes) return false;
  img.assign(kL8CmapBytes + fwd_bytes + bwd_bytes, 0);
  auto put = [&](uint32_t off, int32_t v) { std::memcpy(img.data() + off, &v, 4); };
  auto emit = [&](const Lines8Image& t, uint32_t trans_rel, int cm_half) {
    const int R = t.replicated;
    const int rows = static_cast<int>(t.row_off.size());
    const int col_bytes = 4 * R;
    const int32_t mid_abs = static_cast<int32_t>(kL8AbsTrans + trans_rel) + t.mid_off;
    for (int b = 0; b < 256; b++)
      for (int lane = 0; lane < 32; lane++) put(b * 256 + cm_half * 128 + lane * 4, mid_abs + t.cmap[b] * col_bytes);
    for (int s = 0; s < rows; s++)
      for (int c = 0; c < t.cols; c++) {
        const int target = t.trans[static_cast<size_t>(s) * t.cols + c];
        for (int lane = 0; lane < R; lane++)
          put(kL8CmapBytes + trans_rel + t.mid_off + t.row_off[s] + c * col_bytes + lane * 4,
              t.row_off[target] + (R == 32 ? lane * 4 : 0));
      }
  };
  emit(fwd, 0, 0);
  meta.fwd_root = fwd.row_off[0];
  meta.fwd_repl = fwd.replicated;
  meta.has_bwd = bwd != nullptr;
  if (bwd) {
    emit(*bwd, fwd_bytes, 1);
    meta.bwd_root = bwd->row_off[0];
    meta.bwd_repl = bwd->replicated;
  }
  meta.trans_bytes = fwd_bytes + bwd_bytes;
  return true;
}

struct Lines8Params {
  BatchParams g;         // buffers, mode, lengths, generic tables (slow-path tiles)
  const uint8_t* image;  // [cmap][trans]
  uint32_t trans_bytes;
  int32_t fwd_root, bwd_root;
  int fwd_repl, bwd_repl;
  int use_bwd_table;     // find with a table-driven reverse pass whose image is resident
};

// ---------------------------------------------------------------------------------------------
// device helpers (all shared-memory addresses are 32-bit shared-window addresses)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t phase) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(bar), "r"(phase)
      : "memory");
}
// TMA 1-D bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
// table lookups: constant after the image has landed, so a plain (movable) asm
__device__ __forceinline__ int32_t lds_tab(uint32_t addr) {
  int32_t v;
  asm("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
// haystack bytes: buffers are rewritten every tile
__device__ __forceinline__ uint4 lds_data16(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t lds_data8(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}

// One DFA step on byte K of `word`.  cmsel = 0x00010000 | lane*4 (forward map) or | 0x80 (backward map).
template <int K>
__device__ __forceinline__ void l8_step(uint32_t word, uint32_t cmsel, int32_t& e, uint32_t& mask) {
  // bytes: [0] <- cmsel.b0 (lane*4 [+128]), [1] <- word.bK, [2] <- cmsel.b2 (0x01), [3] <- cmsel.b3 (0x00)
  const uint32_t a = __byte_perm(word, cmsel, 0x7604u | (K << 4));
  e = lds_tab(static_cast<uint32_t>(e + lds_tab(a)));
  mask = __funnelshift_l(static_cast<uint32_t>(e), mask, 1);
}
__device__ __forceinline__ void l8_word(uint32_t w, uint32_t cmsel, int32_t& e, uint32_t& mask) {
  l8_step<0>(w, cmsel, e, mask);
  l8_step<1>(w, cmsel, e, mask);
  l8_step<2>(w, cmsel, e, mask);
  l8_step<3>(w, cmsel, e, mask);
}

// swizzled slot (in 16-byte units) of chunk `c` of tile-local line `line`; 2^log2cpl chunks per line
__device__ __forceinline__ uint32_t l8_slot(uint32_t line, uint32_t c, int log2cpl) {
  const uint32_t cpl = 1u << log2cpl;
  const uint32_t swz = (line >> (3 - log2cpl)) & (cpl - 1);
  return (line << log2cpl) + (c ^ swz);
}

// Generic per-thread walk of line i straight from global memory (irregular tiles).
__device__ __forceinline__ void l8_slow_line(const BatchParams& g, uint64_t i) {
  const uint64_t o0 = g.offsets[i], o1 = g.offsets[i + 1];
  const uint8_t* s = static_cast<const uint8_t*>(g.data) + o0;
  const int64_t len = static_cast<int64_t>(o1 - o0);
  if (g.mode == 0) {
    g.matched[i] = dev_matches<uint8_t>(g, s, len);
  } else if (g.mode == 1) {
    g.matched[i] = dev_contained_in<uint8_t>(g, s, len);
  } else {
    const int64_t e = dev_index_forwards<uint8_t>(g, s, len, 0);
    int64_t st = -1;
    if (e != -1) st = (g.reverse_mode == 2) ? e - g.min_length : dev_index_backwards<uint8_t>(g, s, e - 1, 0, 0x7fffffff);
    g.matched[i] = e != -1;
    g.start[i] = static_cast<int32_t>(st);
    g.end[i] = static_cast<int32_t>(e);
  }
}

__global__ void __launch_bounds__(kL8Threads, 1) lines8_kernel(const Lines8Params p) {
  extern __shared__ __align__(128) uint8_t l8_dyn_smem[];
  const uint32_t base = static_cast<uint32_t>(__cvta_generic_to_shared(l8_dyn_smem));
  const uint32_t tid = threadIdx.x;
  const uint32_t lane4 = (tid & 31) * 4;
  const BatchParams& g = p.g;
  const uint8_t* const data = static_cast<const uint8_t*>(g.data);

  // buffer A fills whatever lies between the start of dynamic shared memory and the class map
  const uint32_t buf_a = (base + 127) & ~127u;
  const bool layout_ok = buf_a <= 0x8000;
  const uint32_t cap = layout_ok ? min(kL8BufBytes, kL8AbsCmap - buf_a) : 0;

  // --- table image: two TMA bulk copies (cmap, trans) completing on one mbarrier
  if (tid == 0) {
    mbar_init(kL8AbsBar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  if (tid == 0 && layout_ok) {
    mbar_expect_tx(kL8AbsBar, kL8CmapBytes + p.trans_bytes);
    tma_bulk_g2s(kL8AbsCmap, p.image, kL8CmapBytes, kL8AbsBar);
    tma_bulk_g2s(kL8AbsTrans, p.image + kL8CmapBytes, p.trans_bytes, kL8AbsBar);
  }

  // --- line geometry: L from the first two offsets (uniform)
  const uint64_t L64 = g.offsets[1] - g.offsets[0];
  int log2cpl = -1;
  if (layout_ok) {
    if (L64 == 16) log2cpl = 0;
    else if (L64 == 32) log2cpl = 1;
    else if (L64 == 64) log2cpl = 2;
    else if (L64 == 128) log2cpl = 3;
  }
  const uint32_t L = static_cast<uint32_t>(L64);
  const uint32_t tile_lines = (log2cpl < 0) ? kL8Threads : min(static_cast<uint32_t>(kL8Threads), cap / L);
  const uint64_t n_tiles = (g.n + tile_lines - 1) / tile_lines;

  // Stage tile `t` into buffer `b`; returns (uniformly) whether the tile is regular.
  auto stage = [&](uint64_t t, int b) -> bool {
    const uint64_t first = t * tile_lines;
    const uint64_t i = first + tid;
    const uint64_t tile_off = g.offsets[first];
    int ok = log2cpl >= 0;
    if (ok && tid < tile_lines && i < g.n) {
      const uint64_t o0 = g.offsets[i], o1 = g.offsets[i + 1];
      ok = (o1 - o0 == L) && (o0 == tile_off + static_cast<uint64_t>(tid) * L) && ((reinterpret_cast<uintptr_t>(data) + o0) & 15) == 0;
    }
    const bool regular = __syncthreads_and(ok) != 0;
    if (regular) {
      const uint64_t lines_here = min(static_cast<uint64_t>(tile_lines), g.n - first);
      const uint32_t n_chunks = static_cast<uint32_t>(lines_here) << log2cpl;
      const uint8_t* src = data + tile_off;
      const uint32_t dst = b ? kL8AbsBufB : buf_a;
      for (uint32_t c = tid; c < n_chunks; c += kL8Threads) {
        const uint32_t line = c >> log2cpl, ch = c & ((1u << log2cpl) - 1);
        cp_async16(dst + (l8_slot(line, ch, log2cpl) << 4), src + (static_cast<uint64_t>(c) << 4));
      }
    }
    cp_async_commit();
    return regular;
  };

  uint64_t t = blockIdx.x;
  int b = 0;
  bool regular = false;
  if (t < n_tiles) regular = stage(t, 0);
  if (layout_ok) mbar_wait(kL8AbsBar, 0);  // table image has landed

  const uint32_t cm_f = 0x00010000u | lane4;
  const uint32_t cm_b = cm_f | 0x80u;
  const int32_t e_root_f = p.fwd_root + (p.fwd_repl == 32 ? static_cast<int32_t>(lane4) : 0);
  const int32_t e_root_b = p.bwd_root + (p.bwd_repl == 32 ? static_cast<int32_t>(lane4) : 0);
  // entry of the forward root row is reached by stepping from a virtual entry whose target is the root
  // (entries are offsets of the *target* row, so the walk starts with e = row_off[root] [+ lane*4])

  for (; t < n_tiles; t += gridDim.x) {
    const uint64_t t_next = t + gridDim.x;
    bool regular_next = false;
    if (t_next < n_tiles) regular_next = stage(t_next, b ^ 1);
    else cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();

    const uint64_t i = t * tile_lines + tid;
    if (tid < tile_lines && i < g.n) {
      if (regular) {
        const uint32_t buf = b ? kL8AbsBufB : buf_a;
        int32_t e = e_root_f;
        int32_t last = g.fwd.root_accepting ? 0 : -1;
        const uint32_t cpl = 1u << log2cpl;
        uint32_t mask = 0;
        for (uint32_t c = 0; c < cpl; c++) {
          const uint4 w = lds_data16(buf + (l8_slot(tid, c, log2cpl) << 4));
          l8_word(w.x, cm_f, e, mask);
          l8_word(w.y, cm_f, e, mask);
          l8_word(w.z, cm_f, e, mask);
          l8_word(w.w, cm_f, e, mask);
          if ((c & 1) || c + 1 == cpl) {  // mask holds <= 32 steps; bit 0 = the most recent one
            if (mask) last = static_cast<int32_t>((c + 1) * 16) - (__ffs(mask) - 1);
            mask = 0;
          }
        }
        if (g.mode == 0) {
          bool m = e < 0;
          if (g.min_length > 4 && static_cast<uint32_t>(g.min_length) > L) m = false;  // DFAMethodComponents.java:75-93
          if (g.max_length != -1 && L > static_cast<uint32_t>(g.max_length)) m = false;
          g.matched[i] = m;
        } else if (g.mode == 1) {
          g.matched[i] = e < 0;
        } else {
          int32_t st = -1;
          if (last != -1) {
            if (g.reverse_mode == 2) {  // start = end - minLength (DFAClassBuilder.java:640-646)
              st = last - g.min_length;
            } else if (g.reverse_mode == 1) {  // single-char reverse scan (:588-614)
              st = 0x7fffffff;
              for (int32_t idx = last - 1; idx >= 0; idx--) {
                const uint32_t sl = l8_slot(tid, static_cast<uint32_t>(idx) >> 4, log2cpl);
                if (lds_data8(buf + (sl << 4) + (idx & 15)) == static_cast<uint32_t>(g.reverse_char)) {
                  st = idx;
                  break;
                }
              }
            } else if (p.use_bwd_table) {  // indexBackwards (:529-586) over the resident BACKWARDS image
              int32_t eb = e_root_b;
              st = g.bwd.root_accepting ? 0 : 0x7fffffff;
              for (int32_t idx = last - 1; idx >= 0; idx--) {
                const uint32_t sl = l8_slot(tid, static_cast<uint32_t>(idx) >> 4, log2cpl);
                const uint32_t byte = lds_data8(buf + (sl << 4) + (idx & 15));
                eb = lds_tab(static_cast<uint32_t>(eb + lds_tab(cm_b | (byte << 8))));
                if (eb < 0) st = idx;
              }
            } else {
              st = static_cast<int32_t>(dev_index_backwards<uint8_t>(g, data + g.offsets[i], last - 1, 0, 0x7fffffff));
            }
          }
          g.matched[i] = last != -1;
          g.start[i] = st;
          g.end[i] = last;
        }
      } else {
        l8_slow_line(g, i);
      }
    }
    __syncthreads();  // everyone is done with buffer b before the stage after next overwrites it
    regular = regular_next;
    b ^= 1;
  }
  cp_async_wait<0>();
}

}  // namespace ndl
