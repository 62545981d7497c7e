// The tile kernels of the batch path: one haystack ("line") per lane, a tile of 32 lines per warp.
//
//   lines8_kernel   class map in shared memory (any pattern whose tables fit): pair tables or stride-1 tables
//   linesq_kernel   packed-compare ("SWAR") classifier, no class map: 4 or 2 chars per transition lookup
// and what they share: tile staging; the fixed-length walks - resident tiles for records of 16 .. 64 bytes (l8_run, l8_run_any),
// rounds of 64 bytes per line for every longer record length (l8_run_rounds, l8_run_rounds_unaligned; l8_pick_geometry chooses);
// the ragged walks - short lines in resident tiles, two per round, lines paired short + long (l8_run_ragged; its long lines
// streamed, l8_stream_group), longer lines sorted by length and streamed in batches of similar lines (ragged_rounds.cuh); iterated
// find on the staged tile (l8_find_all); the reverse passes on the staged tile (l8_reverse, l8_reverse_char) and the result glue
// (l8_finish).  long8.cuh builds the single-haystack kernel on the same pieces; layouts.h builds the table images on the host.
//
// Per char the generated Java loop does two dependent array loads (BYTE_CLASSES[c], then
// STATES[class + state*stride]) plus bookkeeping (DFAClassBuilder.java:438-465).
//
// lines8 ("L" images): the automaton is stepped TWO chars at a time: the class of each char is one shared-memory
// load, the transition for the pair is one more, so a char costs 1.5 loads and the dependent chain is one load per
// two chars.  Shared memory is addressed ABSOLUTELY (shared-window addresses), with a fixed map:
//     [base, 0x10000)      per-warp tile buffers, set 0 (2 KB each)
//     [0x10000, 0x20000)   cmap: 256 slots x 256 B.  Slot b holds, per lane, two int32 maps:
//                            first 128 B   CA[b][lane] = class(b) * C * colBytes          (first char of a pair)
//                            second 128 B  CB[b][lane] = trans base + class(b) * colBytes [+ lane*4]
//                          Lane l only reads word l of a half slot -> bank l -> conflict free, and because
//                          the region starts at 0x10000 the address of that word is ONE byte-permute:
//                          bytes {lane*4 (+128), haystack byte, 0x01, 0x00}.
//     [0x20000, 0x30000)   per-warp tile buffers, set 1
//     [0x30000, 0x30800)   tile buffer set 0 of the warp that did not fit below 0x10000
//     [0x30800, 0x387C0)   trans: rows x C*C entries E (uint32), replicated per lane (entry is 128 B, lane
//                          l reads word l) when that fits, else a single copy.  C = number of distinct
//                          classes the 256 byte values map to.  E = byte offset of the target row
//                          | accept(after 1st char) << 31 | accept(after 2nd char) << 30, so
//                              next entry address = (E & 0x3fffffff) + CA[b1] + CB[b2]   (LOP3 + IADD3)
//                              the two accept bits are shifted into a per-line bit mask  (one SHF)
//                          DEAD is an ordinary absorbing row (device_image.h): no per-char branch.
//                          The BACKWARDS rows follow the forward rows when find() needs the reverse pass.
//     [0x387C0, ...)       mbarrier for the TMA bulk copies that bring cmap + trans in (SASS: UBLKCP)
//   (the stride-1 "S1" variant and the Q images of linesq are described where they are defined below)
//
// Haystack bytes: every warp owns tiles of 32 consecutive lines.  It copies a tile with coalesced 16-byte cp.async
// (LDGSTS) into its own buffer - XOR-swizzling the 16-byte chunk index so that the per-lane 16-byte reads of the
// 32 lines are conflict free too - and double-buffers the next tile against the walk.  Warps never synchronise
// with each other after the tables have landed.
//
// The fixed-length walks need every line of a warp tile to have the same length (and, for the aligned ones, 16-byte alignment);
// this is checked per tile, irregular tiles are walked line by line, and any other batch takes a ragged walk - same results.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <type_traits>
#include <utility>
#include <vector>

#include "../device_image.h"
#include "generic.cuh"
#include "swar_plan.h"

namespace ndl {

constexpr int kL8Warps = 32;
constexpr int kL8Threads = kL8Warps * 32;
constexpr uint32_t kL8WarpBuf = 2048;  // bytes per warp tile buffer
constexpr uint32_t kL8AbsCmap = 0x10000;
constexpr uint32_t kL8CmapBytes = 0x10000;
constexpr uint32_t kL8AbsSet1 = 0x20000;
constexpr uint32_t kL8AbsSpill = 0x30000;  // set-0 buffers that do not fit below the class map
constexpr uint32_t kL8AbsTrans = 0x30800;
constexpr uint32_t kL8AbsBar = 0x387C0;
constexpr uint32_t kL8MaxTransBytes = kL8AbsBar - kL8AbsTrans;  // 32704
constexpr uint32_t kL8AbsEnd = kL8AbsBar + 16;
constexpr uint32_t kL8DynSmem = kL8AbsEnd;  // covers the map for any dynamic base in [0, 0x400]
constexpr uint32_t kL8FlagMask = 0x3fffffffu;
// Second layout ("S1", for automata whose pair table cannot be replicated per bank, e.g. 256-state DFAs):
// a 32 KB class map (128-byte slots) at 0x10000, then a stride-1 transition table of 16-bit entries,
// replicated per lane (entry (row, col) is 64 B, lane l reads halfword l), then the upper tile buffers.
// Two lanes share a bank word, so lanes 2j and 2j+1 conflict (2-way) when they are in different states:
// ncu measures about 1.3 wavefronts per transition lookup on a 256-state DFA over random text, against
// about 3.4 for the unreplicated pair table (profiles/r01_ncu_lines8_c4b_summary.txt).
constexpr uint32_t kS1CmapBytes = 0x8000;
constexpr uint32_t kS1AbsTrans = 0x18000;
constexpr uint32_t kS1MaxTransBytes = 0x10800;  // 67584: 258 rows x 4 classes x 64 B fits
constexpr uint32_t kS1UpperLo = kS1AbsTrans + kS1MaxTransBytes;  // 0x28800

// Char modes.  The walk consumes 32-bit words of the haystack; what a "char" is depends on the mode:
//   kCmBytes  char_width 1: four chars per word, class map indexed by the byte.
//   kCmHi     char_width 2 and the class of a char depends only on its HIGH byte (e.g. `[؀-ۿ]+`):
//             two chars per word, class map indexed by bytes 1 and 3 - the low bytes are never looked at.
//   kCmMixed  char_width 2 and exactly one 256-char page is non-uniform while every other page has one
//             common class (an ASCII pattern over UTF-16 text): class map indexed by the LOW byte, and a
//             select replaces the looked-up value by the common class when the high byte is another page.
// Any other UTF-16 class map is left to the generic kernel.
//   kCmBytes1 char_width 1, ONE char per step over the S1 layout (see kS1* above).
//   kCmBytesH char_width 1, as kCmBytes with one plain copy of the pair table in 16-bit entries (byte offset of the
//             target row / 2 | accept flags << 14): twice the automaton in the same 32 KB - keyword lists and the like.
enum { kCmBytes = 0, kCmHi = 1, kCmMixed = 2, kCmBytes1 = 3, kCmBytesH = 4 };
// SWAR modes ("Q" layout, linesq_kernel): no class map in shared memory at all.  The class of a char comes
// from packed compares (swar_plan.h), the automaton is stepped K = 2 or 4 chars per transition lookup, and
// the column offset of a K-char group is a dot product (IDP.4A) of the compare planes with per-position
// weights.  Encoded as 16 | u16 << 5 | (K == 4) << 3 | hi << 2 | planes:
//   u16 = 1 16-bit table entries (K = 2 only): twice the copies in the same space, see linesq_layout
//   wide (bit 6) char_width 2 with the compares done on 16-bit lanes (thresholds up to 0x8000, chars above share one
//           class): two chars per word, e.g. an ASCII pattern over UTF-16 text - a java.lang.String
//   hi = 0  char_width 1, four chars per word
//   hi = 1  char_width 2 with the class decided by the HIGH byte of a char (as kCmHi): the high bytes of
//           two words are gathered into one (PRMT) and then classified like bytes
__host__ __device__ constexpr int cm_swar(int k, int planes, bool hi, bool u16 = false) {
  return 16 | (u16 ? 32 : 0) | (k == 4 ? 8 : 0) | (hi ? 4 : 0) | planes;
}
__host__ __device__ constexpr bool cm_u16(int cm) { return cm >= 16 && (cm & 32) != 0; }
__host__ __device__ constexpr int cm_swar_wide(int k, int planes) { return 64 | 16 | (k == 4 ? 8 : 0) | planes; }
__host__ __device__ constexpr bool cm_wide(int cm) { return cm >= 16 && (cm & 64) != 0; }
__host__ __device__ constexpr bool cm_is_swar(int cm) { return cm >= 16; }
__host__ __device__ constexpr int cm_k(int cm) { return (cm & 8) ? 4 : 2; }
__host__ __device__ constexpr int cm_planes(int cm) { return cm & 3; }
__host__ __device__ constexpr bool cm_hi(int cm) { return (cm & 4) != 0; }
// Q layout: [kQAbsTrans, + trans_bytes) transition table, then 2 KB tile buffers up to the mbarrier.
constexpr uint32_t kQAbsTrans = 0x800;
constexpr uint32_t kQMaxTransBytes = 0x22000;  // 139264: leaves 43 tile buffers = 21 warps
#ifndef NDL_Q_WARPS
#define NDL_Q_WARPS 28
#endif
constexpr int kQWarps = NDL_Q_WARPS;
constexpr int kQThreads = kQWarps * 32;

// What the kernels need of a SwarPlan; a kernel parameter, so every field is a constant-bank operand.
struct SwarDev {
  uint32_t lo[3], hi[3];  // range bounds replicated into the four bytes of a word
  // IDP.4A weights per plane.  K = 4: [0] forward, [2] reverse.  K = 2: [0], [1] forward pairs (bytes 0,1 and
  // 2,3); [2], [3] reverse pairs (bytes 3,2 and 1,0).  The first char of a group is the most significant digit.
  uint32_t w[3][4];
  uint32_t kmul;          // 128-byte lines per table column: entry address = dot product * kmul + row part
                          // (16-bit entries at K = 4: BYTES per column, address = (dot product * kmul >> 7) + row part)
  uint32_t copy_mask;     // R - 1: lane l uses copy l & (R - 1)
  uint32_t copy_bytes;    // bytes of one copy inside a 128-byte line (4 * 32 / R)
};


// Device image for one (mode, char width): [cmap 64 KB][trans], plus what the kernel needs to start a walk.
struct Lines8Blob {
  uint8_t* dev = nullptr;
  uint32_t trans_bytes = 0;  // multiple of 16
  uint32_t root_entry = 0;   // E of "currently in the root state" (forward table)
  uint32_t bwd_root = 0;     // same for the BACKWARDS table, when resident
  uint32_t bwd_dead = 0;     // row offset of its DEAD row
  uint32_t fwd_dead = 0;     // DEAD row of the forward table, same encoding as root_entry
  uint32_t ua = 0, ub = 0;   // kCmMixed: CA / CB value of the class every other page has
  uint32_t xa = 0, xb = 0;   // UTF-16 modes: CA / CB value of the class of U+FFFF
  int mixed_page = 0;        // kCmMixed: the high byte of the non-uniform page
  uint32_t row_bytes = 0;    // bytes per table row
  int char_mode = kCmBytes;
  int replicated = 0;        // 32 or 1
  int n_cols = 0;            // C
  bool has_bwd = false;
  bool ok = false;
  SwarDev q = {};            // SWAR modes
};

struct Lines8Params {
  BatchParams g;         // buffers, mode, lengths, generic tables (irregular tiles, reverse pass)
  const uint8_t* image;  // [cmap][trans]
  uint32_t trans_bytes;
  uint32_t root_entry;
  uint32_t bwd_root, bwd_dead, fwd_dead;
  uint32_t ua, ub;       // kCmMixed
  uint32_t xa, xb;       // UTF-16 modes: U+FFFF
  int mixed_page;
  int replicated;
  uint32_t row_bytes;
  int char_mode;
  int has_bwd;  // BACKWARDS pair table resident: table-driven reverse pass runs on the staged tile
  SwarDev q;    // SWAR modes
  uint32_t no_rounds;  // experiments (NDL_NO_ROUNDS): fixed-length lines keep the resident-tile walks
  uint32_t rr_min_mean;     // mean line length (bytes) from which ragged batches take the sorted streaming walk: kRrMinMeanBytes (experiments: NDL_RR_MIN_MEAN)
  uint32_t rounds_max_cpl;  // longest record (in 16-byte chunks) walked in rounds: kMaxRoundsCpl (experiments: NDL_ROUNDS_MAX_CPL)
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
__device__ __forceinline__ uint32_t lds_tab(uint32_t addr) {
  uint32_t v;
  asm("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
// haystack bytes: buffers are rewritten every tile
__device__ __forceinline__ uint4 lds_data16(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}

// Per-lane constants of the walk.  sel_a = 0x00010000 | lane*4 selects the CA half of a class-map slot,
// sel_b = sel_a | 0x80 the CB half: bytes {lane*4 [+128], <haystack byte>, 0x01, 0x00} of the address.
struct L8Ctx {
  uint32_t sel_a, sel_b;
  uint32_t page1, page3;  // kCmMixed: mixed page number positioned at byte 1 / byte 3 of a word
  uint32_t ua, ub;        // kCmMixed: CA / CB value of the common class of all other pages
  uint32_t xa, xb;        // UTF-16 modes: CA / CB value of U+FFFF, whose class never follows its page
  uint32_t row_bytes;     // kCmBytes1
  uint32_t root;          // entry value of "in the root state" (SWAR modes: this lane's copy)
  uint32_t bwd_root, bwd_dead;  // the same for the BACKWARDS rows; row of its DEAD state
  uint32_t fwd_dead;            // DEAD row of the forward table
};

// How a table entry encodes state and accept flags, per char mode.
template <int CM>
struct L8Enc {
  static constexpr uint32_t kStateMask =
      CM == kCmBytes1 ? 0x7fffu : CM == kCmBytesH ? 0x3fffu : cm_u16(CM) ? (0xffffu >> cm_k(CM)) : cm_is_swar(CM) ? (0xffffffffu >> cm_k(CM)) : kL8FlagMask;
  // accept flag of the LAST char of a step (kCmBytes1 entries are sign-extended, so bit 30 works there too)
  static constexpr uint32_t kTailFlag =
      CM == kCmBytesH ? 0x4000u : cm_u16(CM) ? (0x8000u >> (cm_k(CM) - 1)) : cm_is_swar(CM) ? (0x80000000u >> (cm_k(CM) - 1)) : 0x40000000u;
};

// One 2-char step.  KA / KB: byte index (within the 32-bit word) that selects the class-map slot of the
// first / second char; HA / HB (kCmMixed only): byte index holding that char's high byte.
template <int CM, int KA, int KB, int HA, int HB>
__device__ __forceinline__ void l8_step2(uint32_t word, const L8Ctx& cx, uint32_t& e, uint32_t& mask) {
  uint32_t ca = lds_tab(__byte_perm(word, cx.sel_a, 0x7604u | (KA << 4)));
  uint32_t cb = lds_tab(__byte_perm(word, cx.sel_b, 0x7604u | (KB << 4)));
  if (CM == kCmMixed) {
    ca = ((word & (0xffu << (8 * HA))) == (HA == 1 ? cx.page1 : cx.page3)) ? ca : cx.ua;
    cb = ((word & (0xffu << (8 * HB))) == (HB == 1 ? cx.page1 : cx.page3)) ? cb : cx.ub;
  }
  if (CM != kCmBytes && CM != kCmBytesH) {  // the char in the low (H == 1) / high (H == 3) half of the word is U+FFFF
    ca = (HA == 1 ? (word & 0xffffu) == 0xffffu : word >= 0xffff0000u) ? cx.xa : ca;
    cb = (HB == 1 ? (word & 0xffffu) == 0xffffu : word >= 0xffff0000u) ? cx.xb : cb;
  }
  if (CM == kCmBytesH) {  // 16-bit entries: row offset / 2 in the low 14 bits, the two flags above
    uint32_t v;
    asm("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"((e & 0x3fffu) * 2u + (ca + cb)));
    e = v;
    mask = __funnelshift_l(e * 0x10000u, mask, 2);
  } else {
    e = lds_tab((e & kL8FlagMask) + ca + cb);
    mask = __funnelshift_l(e, mask, 2);  // mask = mask << 2 | accept(1st) << 1 | accept(2nd)
  }
}

__device__ __forceinline__ int32_t lds_tab_s16(uint32_t addr) {
  int32_t v;
  asm("ld.shared.s16 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
// One 1-char step over the S1 layout on byte K of `word`: class-map slot address = 0x10000 | byte << 7 | lane*4,
// entry = row index | accept << 15, loaded sign-extended so that the accept flag is the sign bit.
template <int K>
__device__ __forceinline__ void l8_step1(uint32_t word, const L8Ctx& cx, uint32_t& e, uint32_t& mask) {
  const uint32_t sh = K == 0 ? (word << 7) : (word >> (8 * K - 7));
  const uint32_t cm = lds_tab((sh & 0x7f80u) | cx.sel_a);
  e = static_cast<uint32_t>(lds_tab_s16((e & 0x7fffu) * cx.row_bytes + cm));
  mask = __funnelshift_l(e, mask, 1);
}

// All chars of one word, forwards / backwards.
template <int CM>
__device__ __forceinline__ void l8_word(uint32_t w, const L8Ctx& cx, uint32_t& e, uint32_t& mask) {
  if (CM == kCmBytes1) {
    l8_step1<0>(w, cx, e, mask);
    l8_step1<1>(w, cx, e, mask);
    l8_step1<2>(w, cx, e, mask);
    l8_step1<3>(w, cx, e, mask);
  } else if (CM == kCmBytes || CM == kCmBytesH) {
    l8_step2<CM, 0, 1, 0, 0>(w, cx, e, mask);
    l8_step2<CM, 2, 3, 0, 0>(w, cx, e, mask);
  } else if (CM == kCmHi) {
    l8_step2<CM, 1, 3, 1, 3>(w, cx, e, mask);
  } else {
    l8_step2<CM, 0, 2, 1, 3>(w, cx, e, mask);
  }
}
template <int CM>
__device__ __forceinline__ void l8_word_rev(uint32_t w, const L8Ctx& cx, uint32_t& e, uint32_t& mask) {
  if (CM == kCmBytes1) {
    l8_step1<3>(w, cx, e, mask);
    l8_step1<2>(w, cx, e, mask);
    l8_step1<1>(w, cx, e, mask);
    l8_step1<0>(w, cx, e, mask);
  } else if (CM == kCmBytes || CM == kCmBytesH) {
    l8_step2<CM, 3, 2, 0, 0>(w, cx, e, mask);
    l8_step2<CM, 1, 0, 0, 0>(w, cx, e, mask);
  } else if (CM == kCmHi) {
    l8_step2<CM, 3, 1, 3, 1>(w, cx, e, mask);
  } else {
    l8_step2<CM, 2, 0, 3, 1>(w, cx, e, mask);
  }
}
template <int CM>
struct L8Chars {
  static constexpr int kBytes = (CM == kCmBytes || CM == kCmBytes1 || CM == kCmBytesH || (cm_is_swar(CM) && !cm_hi(CM) && !cm_wide(CM))) ? 1 : 2;   // bytes per char
  static constexpr int kPerChunk = 16 / kBytes;           // chars (= accept bits) per 16-byte chunk
};

// SWAR modes: one word = four slot values (bytes, or gathered high bytes).  Packed compares give one plane per
// range (swar_plan.h), IDP.4A turns the planes into the column offset of the group, one lookup steps K chars.
__device__ __forceinline__ uint32_t lds_tab_u16(uint32_t addr) {
  uint32_t v;
  asm("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}

template <int CM, bool REV>
__device__ __forceinline__ void q_word(uint32_t w, const SwarDev& q, uint32_t& e, uint32_t& mask) {
  constexpr int P = cm_planes(CM), K = cm_k(CM);
  constexpr uint32_t kState = L8Enc<CM>::kStateMask;
  const uint32_t w80 = w | 0x80808080u;
  uint32_t nm;  // kept out of the compiler's hands: one LOP3 here, then one per plane
  asm("lop3.b32 %0, %1, 0x80808080, 0, 0x0c;" : "=r"(nm) : "r"(w));  // ~w & 0x80808080 (LUT: ~a & b)
  uint32_t pl[P];
#pragma unroll
  for (int i = 0; i < P; i++) pl[i] = ((w80 - q.lo[i]) ^ (w80 - q.hi[i])) & nm;
  if (K == 4) {
    uint32_t dp = 0;
#pragma unroll
    for (int i = 0; i < P; i++) dp = __dp4a(pl[i], q.w[i][REV ? 2 : 0], dp);
    if (cm_u16(CM)) {
      // 16-bit entries, few copies (large automata): the column stride is kmul BYTES (any multiple of 2), and the
      // dot product is 128 x column: (dp * kmul) >> 7, which folds into the add (LEA.HI).  Flags in bits 15:12.
      e = lds_tab_u16(((dp * q.kmul) >> 7) + (e & kState));
      mask = __funnelshift_l(e * 0x10000u, mask, 4);
    } else {
      e = lds_tab(dp * q.kmul + (e & kState));
      mask = __funnelshift_l(e, mask, 4);
    }
  } else {
    uint32_t da = 0, db = 0;
#pragma unroll
    for (int i = 0; i < P; i++) {
      da = __dp4a(pl[i], q.w[i][REV ? 2 : 0], da);
      db = __dp4a(pl[i], q.w[i][REV ? 3 : 1], db);
    }
    if (cm_u16(CM)) {  // 16-bit entries: flags in bits 15:14; the shift up is an IMAD (the less loaded pipe)
      e = lds_tab_u16(da * q.kmul + (e & kState));
      mask = __funnelshift_l(e * 0x10000u, mask, 2);
      e = lds_tab_u16(db * q.kmul + (e & kState));
      mask = __funnelshift_l(e * 0x10000u, mask, 2);
    } else {
      e = lds_tab(da * q.kmul + (e & kState));
      mask = __funnelshift_l(e, mask, 2);
      e = lds_tab(db * q.kmul + (e & kState));
      mask = __funnelshift_l(e, mask, 2);
    }
  }
}

// 16-bit lanes: two chars per word.  Same compares with 0x8000 in place of 0x80; a plane has bit 15 / bit 31 set,
// i.e. byte 1 / byte 3 = 0x80, so IDP.4A works on it unchanged with the weights in bytes 1 and 3.
// K = 4: one lookup per two words (weights [0], [1] forwards, [2], [3] backwards); K = 2: one per word.
template <int CM, bool REV>
__device__ __forceinline__ void q_wide(uint32_t w0, uint32_t w1, const SwarDev& q, uint32_t& e, uint32_t& mask) {
  constexpr int P = cm_planes(CM), K = cm_k(CM);
  constexpr uint32_t kState = L8Enc<CM>::kStateMask;
  uint32_t d0 = 0, d1 = 0;
  {
    const uint32_t w80 = w0 | 0x80008000u;
    uint32_t nm;
    asm("lop3.b32 %0, %1, 0x80008000, 0, 0x0c;" : "=r"(nm) : "r"(w0));
#pragma unroll
    for (int i = 0; i < P; i++) d0 = __dp4a(((w80 - q.lo[i]) ^ (w80 - q.hi[i])) & nm, q.w[i][REV ? 2 : 0], d0);
  }
  {
    const uint32_t w80 = w1 | 0x80008000u;
    uint32_t nm;
    asm("lop3.b32 %0, %1, 0x80008000, 0, 0x0c;" : "=r"(nm) : "r"(w1));
    if (K == 4) d1 = d0;  // one column offset for the four chars of both words
#pragma unroll
    for (int i = 0; i < P; i++) d1 = __dp4a(((w80 - q.lo[i]) ^ (w80 - q.hi[i])) & nm, q.w[i][K == 4 ? (REV ? 3 : 1) : (REV ? 2 : 0)], d1);
  }
  if (K == 4) {
    e = lds_tab(d1 * q.kmul + (e & kState));
    mask = __funnelshift_l(e, mask, 4);
  } else {
    e = lds_tab(d0 * q.kmul + (e & kState));  // (the arguments come in walk order)
    mask = __funnelshift_l(e, mask, 2);
    e = lds_tab(d1 * q.kmul + (e & kState));
    mask = __funnelshift_l(e, mask, 2);
  }
}

// All chars of one 16-byte chunk, forwards / backwards.  Afterwards the low L8Chars<CM>::kPerChunk bits of
// `mask` are the accept flags of these chars, bit 0 = the char walked last.
template <int CM>
__device__ __forceinline__ void l8_chunk(const uint4& w, const SwarDev& q, const L8Ctx& cx, uint32_t& e, uint32_t& mask) {
  if constexpr (cm_is_swar(CM)) {
    if constexpr (cm_wide(CM)) {
      q_wide<CM, false>(w.x, w.y, q, e, mask);
      q_wide<CM, false>(w.z, w.w, q, e, mask);
    } else if constexpr (cm_hi(CM)) {
      q_word<CM, false>(__byte_perm(w.x, w.y, 0x7531), q, e, mask);
      q_word<CM, false>(__byte_perm(w.z, w.w, 0x7531), q, e, mask);
    } else {
      q_word<CM, false>(w.x, q, e, mask);
      q_word<CM, false>(w.y, q, e, mask);
      q_word<CM, false>(w.z, q, e, mask);
      q_word<CM, false>(w.w, q, e, mask);
    }
  } else {
    l8_word<CM>(w.x, cx, e, mask);
    l8_word<CM>(w.y, cx, e, mask);
    l8_word<CM>(w.z, cx, e, mask);
    l8_word<CM>(w.w, cx, e, mask);
  }
}
template <int CM>
__device__ __forceinline__ void l8_chunk_rev(const uint4& w, const SwarDev& q, const L8Ctx& cx, uint32_t& e, uint32_t& mask) {
  if constexpr (cm_is_swar(CM)) {
    if constexpr (cm_wide(CM)) {
      q_wide<CM, true>(w.w, w.z, q, e, mask);
      q_wide<CM, true>(w.y, w.x, q, e, mask);
    } else if constexpr (cm_hi(CM)) {
      q_word<CM, true>(__byte_perm(w.z, w.w, 0x7531), q, e, mask);
      q_word<CM, true>(__byte_perm(w.x, w.y, 0x7531), q, e, mask);
    } else {
      q_word<CM, true>(w.w, q, e, mask);
      q_word<CM, true>(w.z, q, e, mask);
      q_word<CM, true>(w.y, q, e, mask);
      q_word<CM, true>(w.x, q, e, mask);
    }
  } else {
    l8_word_rev<CM>(w.w, cx, e, mask);
    l8_word_rev<CM>(w.z, cx, e, mask);
    l8_word_rev<CM>(w.y, cx, e, mask);
    l8_word_rev<CM>(w.x, cx, e, mask);
  }
}

// Realign 16 bytes that start `sh` bytes into chunk x (continuing in chunk y) into four words.
struct L8Align {
  bool q1, q2;
  uint32_t r8;
  __device__ __forceinline__ explicit L8Align(uint32_t sh) : q1((sh & 4u) != 0), q2((sh & 8u) != 0), r8((sh & 3u) * 8u) {}
  __device__ __forceinline__ uint4 apply(const uint4& x, const uint4& y) const {
    const uint32_t a0 = q2 ? x.z : x.x, a1 = q2 ? x.w : x.y, a2 = q2 ? y.x : x.z, a3 = q2 ? y.y : x.w;
    const uint32_t a4 = q2 ? y.z : y.x, a5 = q2 ? y.w : y.y;
    const uint32_t w0 = q1 ? a1 : a0, w1 = q1 ? a2 : a1, w2 = q1 ? a3 : a2, w3 = q1 ? a4 : a3, w4 = q1 ? a5 : a4;
    return make_uint4(__funnelshift_r(w0, w1, r8), __funnelshift_r(w1, w2, r8), __funnelshift_r(w2, w3, r8), __funnelshift_r(w3, w4, r8));
  }
};

// indexBackwards(end - 1, 0) (DFAClassBuilder.java:529-586) over the staged tile.  `chunk_addr(c)` gives the
// shared address of 16-byte chunk c of the byte space in which the line starts at byte `ps`; the match ends
// (exclusive) at char index `last`.  Walks 16-byte windows downwards; steps that fall before the start of
// the line only produce accept bits that are shifted out.  Returns the smallest accepting index, or INT_MAX.
template <int CM, typename ChunkAddr>
__device__ __forceinline__ int32_t l8_reverse(const Lines8Params& p, ChunkAddr chunk_addr, uint32_t ps, int32_t last, const L8Ctx& cx,
                                              bool bwd_root_accepting) {
  constexpr int kPer = L8Chars<CM>::kPerChunk;
  int32_t st = bwd_root_accepting ? 0 : 0x7fffffff;
  uint32_t e = cx.bwd_root;
  for (int32_t rem = last; rem > 0; rem -= kPer) {
    const uint32_t h = ps + static_cast<uint32_t>(rem) * L8Chars<CM>::kBytes;  // window = bytes [h - 16, h)
    const uint32_t q = h >> 4;
    const uint4 y = lds_data16(chunk_addr(q));
    const uint4 x = lds_data16(chunk_addr(q > 0 ? q - 1 : 0));
    const uint4 w = L8Align(h & 15u).apply(x, y);
    uint32_t mask = 0;
    l8_chunk_rev<CM>(w, p.q, cx, e, mask);
    const int32_t valid = rem < kPer ? rem : kPer;
    mask >>= (kPer - valid);  // drop the steps taken before the start of the line
    const int32_t cand = rem - valid + (__ffs(mask) - 1);
    st = mask ? cand : st;
    if ((e & L8Enc<CM>::kStateMask) == cx.bwd_dead) break;
  }
  return st;
}

// The single-char reverse scan (DFAClassBuilder.java:588-614) over the staged tile: the largest index below
// `last` (relative to the byte position ps) whose char equals `rc`, or INT_MAX.  Packed equality test per word:
// x = w ^ rc..rc has a zero lane exactly where the char matches; ((x & 0x7f..) + 0x7f..) | x has the lane's top bit
// clear exactly for zero lanes (no carries between lanes).
template <int CM, typename ChunkAddr>
__device__ __forceinline__ int32_t l8_reverse_char(ChunkAddr chunk_addr, uint32_t ps, int32_t last, int rc) {
  constexpr int kPer = L8Chars<CM>::kPerChunk, kBytes = L8Chars<CM>::kBytes;
  constexpr uint32_t kLow = kBytes == 1 ? 0x7f7f7f7fu : 0x7fff7fffu;
  if (kBytes == 1 && rc > 0xff) return 0x7fffffff;
  const uint32_t rep = static_cast<uint32_t>(rc) * (kBytes == 1 ? 0x01010101u : 0x00010001u);
  for (int32_t rem = last; rem > 0; rem -= kPer) {
    const uint32_t h = ps + static_cast<uint32_t>(rem) * kBytes;  // window = bytes [h - 16, h)
    const uint32_t q = h >> 4;
    const uint4 y = lds_data16(chunk_addr(q));
    const uint4 x = lds_data16(chunk_addr(q > 0 ? q - 1 : 0));
    const uint4 w = L8Align(h & 15u).apply(x, y);
    const uint32_t words[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
    for (int j = 3; j >= 0; j--) {
      const uint32_t d = words[j] ^ rep;
      const uint32_t z = ~((((d & kLow) + kLow) | d) | kLow);  // top bit of every lane whose char equals rc
      if (z) {
        const int32_t lane_in_word = (31 - __clz(z)) / (8 * kBytes);
        const int32_t cand = rem - kPer + j * (4 / kBytes) + lane_in_word;
        return cand >= 0 ? cand : 0x7fffffff;  // (a hit before the start of the line: every other hit lies further back)
      }
    }
  }
  return 0x7fffffff;
}

// swizzled slot (in 16-byte units) of chunk `c` of tile-local line `line`; 2^log2cpl chunks per line.
// For lines of >= 128 bytes (cpl >= 8) the low 3 chunk bits are XORed with the line number.
__device__ __forceinline__ uint32_t l8_slot(uint32_t line, uint32_t c, int log2cpl) {
  const uint32_t cpl = 1u << log2cpl;
  const uint32_t swz = (log2cpl <= 3) ? ((line >> (3 - log2cpl)) & (cpl - 1)) : (line & 7);
  return (line << log2cpl) + (c ^ swz);
}

// Generic per-thread walk of line i straight from global memory (irregular tiles).
template <typename CharT>
__device__ __forceinline__ void l8_slow_line(const BatchParams& g, uint64_t i) {
  const uint64_t o0 = batch_off(g, i), o1 = batch_off(g, i + 1);
  const CharT* s = static_cast<const CharT*>(g.data) + o0;
  const int64_t len = static_cast<int64_t>(o1 - o0);
  if (g.mode == kModeFindAll) {
    dev_find_all_line<CharT>(g, i);
  } else if (g.mode == 0) {
    g.matched[i] = dev_matches<CharT>(g, s, len);
  } else if (g.mode == 1) {
    g.matched[i] = dev_contained_in<CharT>(g, s, len);
  } else {
    const int64_t from = g.from ? g.from[i] : 0;  // find(from, to): DFAClassBuilder.java:625-659
    const int64_t e = dev_index_forwards<CharT>(g, s, len, from);
    int64_t st = -1;
    if (e != -1) st = (g.reverse_mode == 2) ? e - g.min_length : dev_index_backwards<CharT>(g, s, e - 1, from, 0x7fffffff);
    g.matched[i] = e != -1;
    g.start[i] = static_cast<int32_t>(st);
    g.end[i] = static_cast<int32_t>(e);
  }
}

// Results of one line.  `last`: char index after the last accepting step or -1; `tail_accept`: the state at
// the end of the line is accepting (matches()); ps / chunk_addr: where the line sits in shared memory.
template <int CM, typename CharT, typename ChunkAddr>
__device__ __forceinline__ void l8_finish(const Lines8Params& p, const L8Ctx& cx, uint64_t i, uint32_t len_chars, int32_t last,
                                          bool tail_accept, uint32_t ps, ChunkAddr chunk_addr, int32_t from = 0, bool resident = true,
                                          bool have_st = false, int32_t st_pre = 0) {
  // `from`: find(from, to) started `from` chars into the line (mode 2 only); ps, len_chars and last are relative to it,
  // and the reverse pass stops there (its lower bound, DFAClassBuilder.java:640-659).  `resident`: the whole line is in
  // the tile buffer that chunk_addr addresses (false for streamed lines: their reverse pass reads global memory).
  // `have_st`: the table-driven reverse pass has been run already (pooled across the warp), st_pre is its result.
  const BatchParams& g = p.g;
  if (g.mode == 0) {
    bool m = tail_accept;
    if (g.min_length > 4 && static_cast<uint32_t>(g.min_length) > len_chars) m = false;  // DFAMethodComponents.java:75-93
    if (g.max_length != -1 && len_chars > static_cast<uint32_t>(g.max_length)) m = false;
    g.matched[i] = m;
  } else if (g.mode == 1) {
    g.matched[i] = last != -1;  // accepting rows of the containedIn automaton are absorbing
  } else if (g.reverse_mode == 2) {  // start = end - minLength (DFAClassBuilder.java:640-646): no branch on the line's own result
    const bool hit = last != -1;
    g.matched[i] = hit;
    g.start[i] = hit ? last + from - g.min_length : -1;
    g.end[i] = hit ? last + from : -1;
  } else {
    int32_t st = -1;
    if (last != -1) {
      if (g.reverse_mode == 0 && p.has_bwd && resident) {  // indexBackwards (:529-586) on the staged tile
        st = have_st ? st_pre : l8_reverse<CM>(p, chunk_addr, ps, last, cx, g.bwd.root_accepting != 0);
        if (st != 0x7fffffff) st += from;
      } else if (g.reverse_mode == 1 && resident) {  // single-char reverse scan (:588-614) on the staged tile
        st = l8_reverse_char<CM>(chunk_addr, ps, last, g.reverse_char);
        if (st != 0x7fffffff) st += from;
      } else {  // no resident BACKWARDS table, or the line is not resident (streamed lines): global memory
        st = static_cast<int32_t>(dev_index_backwards<CharT>(g, static_cast<const CharT*>(g.data) + batch_off(g, i), last + from - 1, from,
                                                             0x7fffffff));
      }
      last += from;
    }
    g.matched[i] = last != -1;
    g.start[i] = st;
    g.end[i] = last;
  }
}

// ---------------------------------------------------------------------------------------------
// Fixed-length lines: the walk of regular warp tiles, specialised on the line length in BYTES
// (16 << LOG2CPL) and the char mode.
// ---------------------------------------------------------------------------------------------
template <int CM, bool kOffsets, bool kPartial = false>
__device__ __forceinline__ void l8_run_rounds(const Lines8Params& p, const L8Ctx& cx, const uint32_t buf0, const uint32_t buf1,
                                              const uint32_t lane, const uint32_t warp_global, const uint32_t n_warps, const uint32_t cpl,
                                              const uint32_t line_lo);  // (defined below)

template <int LOG2CPL>
struct L8Geom {
  static constexpr uint32_t kCpl = 1u << LOG2CPL;                                 // 16-byte chunks per line
  static constexpr uint32_t kL = 16u * kCpl;                                      // bytes per line
  static constexpr uint32_t kTileLines = (kL8WarpBuf / kL) < 32u ? (kL8WarpBuf / kL) : 32u;
  static constexpr uint32_t kCopies = kTileLines * kCpl / 32u;                    // cp.async per lane per tile
};

// kOffsets: the batch has an offsets array (ndl_match_batch); false: equally spaced records (ndl_match_lines) - a template
// parameter so that the per-tile code holds one of the two address computations, not both under predicates.
template <int LOG2CPL, int CM, bool kOffsets>
__device__ __forceinline__ void l8_run(const Lines8Params& p, const L8Ctx& cx, const uint32_t buf0, const uint32_t buf1,
                                       const uint32_t lane, const uint32_t warp_global, const uint32_t n_warps) {
  using G = L8Geom<LOG2CPL>;
  using CharT = typename std::conditional<L8Chars<CM>::kBytes == 1, uint8_t, uint16_t>::type;
  constexpr uint32_t kCharBytes = L8Chars<CM>::kBytes;
  constexpr uint32_t kPer = L8Chars<CM>::kPerChunk;
  constexpr uint32_t kLenChars = G::kL / kCharBytes;
  const BatchParams& g = p.g;
  const uint8_t* const data = static_cast<const uint8_t*>(g.data);
  const uint32_t n = static_cast<uint32_t>(g.n);  // the host only takes this path for n < 2^31
  const uint32_t n_full = n / G::kTileLines;
  const bool active = lane < G::kTileLines;

  // per-lane constants: where this lane's copies land
  uint32_t dst_off[G::kCopies];
#pragma unroll
  for (uint32_t k = 0; k < G::kCopies; k++) {
    const uint32_t c = lane + 32 * k;
    dst_off[k] = l8_slot(c >> LOG2CPL, c & (G::kCpl - 1), LOG2CPL) << 4;
  }
  const uint32_t lane_line = active ? lane : 0;

  // offsets (in chars) of line (tile * kTileLines + lane) and the next one
  auto load_offsets = [&](uint32_t tile, uint64_t& o0, uint64_t& o1) {
    const uint32_t i = tile * G::kTileLines + lane_line;
    if constexpr (kOffsets) {
      o0 = g.offsets[i];
      o1 = g.offsets[i + 1];
    } else {
      o0 = static_cast<uint64_t>(i) * g.line_chars;
      o1 = o0 + g.line_chars;
    }
  };
  // Issue the copies of a tile whose offsets are (o0, o1); returns whether the tile is regular.
  auto stage = [&](uint64_t o0, uint64_t o1, uint32_t buf) -> bool {
    const uint64_t tile_off = o0 * kCharBytes - static_cast<uint64_t>(lane_line) * G::kL;  // bytes; same on every lane iff regular
    const uint8_t* src = data + tile_off;
    const bool ok = (o1 - o0 == kLenChars) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
    const bool regular = __all_sync(0xffffffffu, ok) != 0;
    if (regular) {
      src += lane * 16;
#pragma unroll
      for (uint32_t k = 0; k < G::kCopies; k++) cp_async16(buf + dst_off[k], src + 512 * k);
    }
    cp_async_commit();
    return regular;
  };

  uint32_t t = warp_global;
  uint32_t cur = buf0, nxt = buf1;
  bool regular = false;
  uint64_t a0 = 0, a1 = 0;  // offsets of the tile after the current one
  const int32_t last0 = g.fwd.root_accepting ? 0 : -1;
  if (t < n_full) {
    load_offsets(t, a0, a1);
    regular = stage(a0, a1, cur);
    if (t + n_warps < n_full) load_offsets(t + n_warps, a0, a1);
  }
  for (; t < n_full; t += n_warps) {
    bool regular_next = false;
    if (t + n_warps < n_full) {
      regular_next = stage(a0, a1, nxt);
      if (t + 2 * n_warps < n_full) load_offsets(t + 2 * n_warps, a0, a1);  // prefetch: consumed next iteration
    } else {
      cp_async_commit();
    }
    cp_async_wait<1>();
    __syncwarp();

    const uint32_t i = t * G::kTileLines + lane;
    if (regular) {
      if (active) {
        uint32_t e = cx.root;
        int32_t last = last0;
        uint32_t mask = 0;
#pragma unroll
        for (uint32_t c = 0; c < G::kCpl; c++) {
          const uint4 w = lds_data16(cur + (l8_slot(lane, c, LOG2CPL) << 4));
          l8_chunk<CM>(w, p.q, cx, e, mask);
          constexpr uint32_t kFlush = 32 / kPer;  // chunks whose accept bits fit in the 32-bit mask
          if ((c % kFlush) == kFlush - 1 || c + 1 == G::kCpl) {  // bit 0 = the most recent char
            const int32_t cand = static_cast<int32_t>((c + 1) * kPer + 1) - __ffs(mask);
            last = mask ? cand : last;
            mask = 0;
          }
        }
        l8_finish<CM, CharT>(p, cx, i, kLenChars, last, (e & L8Enc<CM>::kTailFlag) != 0, 0u,
                             [&](uint32_t ch) { return cur + (l8_slot(lane, ch & (G::kCpl - 1), LOG2CPL) << 4); });
      }
    } else if (active) {
      l8_slow_line<CharT>(g, i);
    }
    __syncwarp();  // every lane is done with `cur` before the stage after next overwrites it
    regular = regular_next;
    const uint32_t tmp = cur;
    cur = nxt;
    nxt = tmp;
  }
  cp_async_wait<0>();
  // the partial last tile: the rounds walk takes it (one warp).  A per-thread walk from global memory costs about 0.2 us per
  // byte of line - 15 us for 64-byte lines, more than a tenth of a 10 M-line launch.
  if (n_full * G::kTileLines < n)
    l8_run_rounds<CM, kOffsets, true>(p, cx, buf0, buf1, lane, (warp_global + n_warps - n_full % n_warps) % n_warps, n_warps, G::kCpl,
                                      n_full * G::kTileLines);
}

// byte offset of 16-byte chunk c of a byte-contiguous tile: the chunk index XOR-swizzled by (chunk >> 3) (ragged tiles, below)
__device__ __forceinline__ uint32_t l8_rslot(uint32_t c) { return (c ^ ((c >> 3) & 7)) << 4; }

// ---------------------------------------------------------------------------------------------
// Fixed-length lines whose byte length is a multiple of 16 but not a power of two (48-, 80-, 96-, 112-byte records):
// the same tile walk with the chunk count per line as a run-time value.  A tile is min(32, 2048 / L) lines; chunk c of
// the tile (line-major) sits in the XOR-swizzled slot of the ragged layout.  Every tile re-checks that its lines are
// equally spaced and 16-byte aligned, like l8_run.
// ---------------------------------------------------------------------------------------------
template <int CM, bool kOffsets>
__device__ __forceinline__ void l8_run_any(const Lines8Params& p, const L8Ctx& cx, const uint32_t buf0, const uint32_t buf1,
                                           const uint32_t lane, const uint32_t warp_global, const uint32_t n_warps, const uint32_t cpl) {
  using CharT = typename std::conditional<L8Chars<CM>::kBytes == 1, uint8_t, uint16_t>::type;
  constexpr uint32_t kCharBytes = L8Chars<CM>::kBytes;
  constexpr uint32_t kPer = L8Chars<CM>::kPerChunk;
  const BatchParams& g = p.g;
  const uint8_t* const data = static_cast<const uint8_t*>(g.data);
  const uint32_t n = static_cast<uint32_t>(g.n);
  const uint32_t line_bytes = 16u * cpl, len_chars = line_bytes / kCharBytes;
  const uint32_t tile_lines = min(32u, kL8WarpBuf / line_bytes);
  const uint32_t tile_chunks = tile_lines * cpl;
  const uint32_t n_full = n / tile_lines;
  const bool active = lane < tile_lines;
  const uint32_t lane_line = active ? lane : 0;

  auto load_offsets = [&](uint32_t tile, uint64_t& o0, uint64_t& o1) {
    const uint32_t i = tile * tile_lines + lane_line;
    if constexpr (kOffsets) {
      o0 = g.offsets[i];
      o1 = g.offsets[i + 1];
    } else {
      o0 = static_cast<uint64_t>(i) * g.line_chars;
      o1 = o0 + g.line_chars;
    }
  };
  auto stage = [&](uint64_t o0, uint64_t o1, uint32_t buf) -> bool {
    const uint64_t tile_off = o0 * kCharBytes - static_cast<uint64_t>(lane_line) * line_bytes;  // same on every lane iff regular
    const uint8_t* src = data + tile_off;
    const bool ok = (o1 - o0 == len_chars) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
    const bool regular = __all_sync(0xffffffffu, ok) != 0;
    if (regular)
      for (uint32_t c = lane; c < tile_chunks; c += 32) cp_async16(buf + l8_rslot(c), src + 16 * c);
    cp_async_commit();
    return regular;
  };

  uint32_t t = warp_global;
  uint32_t cur = buf0, nxt = buf1;
  bool regular = false;
  uint64_t a0 = 0, a1 = 0;  // offsets of the tile after the current one
  if (t < n_full) {
    load_offsets(t, a0, a1);
    regular = stage(a0, a1, cur);
    if (t + n_warps < n_full) load_offsets(t + n_warps, a0, a1);
  }
  for (; t < n_full; t += n_warps) {
    bool regular_next = false;
    if (t + n_warps < n_full) {
      regular_next = stage(a0, a1, nxt);
      if (t + 2 * n_warps < n_full) load_offsets(t + 2 * n_warps, a0, a1);
    } else {
      cp_async_commit();
    }
    cp_async_wait<1>();
    __syncwarp();
    const uint32_t i = t * tile_lines + lane;
    if (regular) {
      if (active) {
        const uint32_t c0 = lane * cpl;
        uint32_t e = cx.root;
        int32_t last = g.fwd.root_accepting ? 0 : -1;
        uint32_t mask = 0;
        constexpr uint32_t kFlush = 32 / kPer;  // chunks whose accept bits fit in the 32-bit mask
        for (uint32_t c = 0; c < cpl; c++) {
          const uint4 w = lds_data16(cur + l8_rslot(c0 + c));
          l8_chunk<CM>(w, p.q, cx, e, mask);
          if ((c % kFlush) == kFlush - 1 || c + 1 == cpl) {  // bit 0 = the most recent char
            const int32_t cand = static_cast<int32_t>((c + 1) * kPer + 1) - __ffs(mask);
            last = mask ? cand : last;
            mask = 0;
          }
        }
        l8_finish<CM, CharT>(p, cx, i, len_chars, last, (e & L8Enc<CM>::kTailFlag) != 0, 0u,
                             [&](uint32_t ch) { return cur + l8_rslot(c0 + min(ch, cpl - 1)); });
      }
    } else if (active) {
      l8_slow_line<CharT>(g, i);
    }
    __syncwarp();  // every lane is done with `cur` before the stage after next overwrites it
    regular = regular_next;
    const uint32_t tmp = cur;
    cur = nxt;
    nxt = tmp;
  }
  cp_async_wait<0>();
  // the partial last tile: the rounds walk takes it (one warp)
  if (n_full * tile_lines < n)
    l8_run_rounds<CM, kOffsets, true>(p, cx, buf0, buf1, lane, (warp_global + n_warps - n_full % n_warps) % n_warps, n_warps, cpl, n_full * tile_lines);
}

// ---------------------------------------------------------------------------------------------
// Ragged lines.  A warp tile is the longest run of <= 32 consecutive lines whose bytes (from the 16-byte
// boundary below the first line) fit in a 2 KB buffer; chunks are stored XOR-swizzled by (chunk >> 3) so that
// lines of about 64 bytes still read conflict free.  A lane reads its line as aligned 16-byte chunks and
// realigns them in registers (word select by the start offset, then funnel shifts).  The walk runs over whole
// 16-byte windows: chars past the end of a line only produce accept bits past the end, which are shifted out,
// so no per-char bounds test is needed.
//
// One line per lane would leave half the lanes idle (a warp walks as long as its longest line: 51 % lane
// utilisation on lines of 8..120 bytes).  So a warp takes TWO tiles at a time, one per buffer, sorts each by
// walk length (bitonic sort over shuffles), and lane k walks the k-th shortest line of the first tile followed
// by the k-th longest of the second in ONE loop - the sums are nearly equal across lanes (82 %).  The copies
// of a pair of tiles are not overlapped with its walk; the other warps of the SM cover them.
// ---------------------------------------------------------------------------------------------

// ascending bitonic sort of one key per lane
__device__ __forceinline__ uint32_t warp_sort32(uint32_t key, uint32_t lane) {
#pragma unroll
  for (uint32_t k = 2; k <= 32; k <<= 1) {
#pragma unroll
    for (uint32_t j = k >> 1; j > 0; j >>= 1) {
      const uint32_t other = __shfl_xor_sync(0xffffffffu, key, j);
      const bool up = (lane & k) == 0;
      const bool low = (lane & j) == 0;
      key = (low == up) ? min(key, other) : max(key, other);
    }
  }
  return key;
}

// Long lines (more than about 256 bytes: fewer than 8 fit a tile, down to none).  One line per lane as before, but
// streamed: every round stages the next 64 bytes of each lane's own line (the lane's four aligned 16-byte chunks, in the
// swizzled slots of a 64-byte-line tile) while the previous 64 are walked, so lines of any length keep all 32 lanes
// busy and nothing depends on a line fitting a buffer.  Chunk k of a line is needed as the upper half of walk step
// k - 1 (the lower half is carried in registers), so a round performs the steps whose upper chunk it holds.
// The reverse pass of find() reads global memory (the line is not resident).
// `own`, `line`: whether this lane has a line, and which.  defer_rev (find() with the table-driven reverse pass resident): a
// line that matched gets matched / end written here and is handed back - return value true, *rev_end / *rev_from - for the
// caller's reverse queue (ragged_rounds.cuh) instead of running the generic reverse loop over global memory.
template <int CM, typename CharT>
__device__ __forceinline__ bool l8_stream_lines(const Lines8Params& p, const L8Ctx& cx, const uint32_t buf0, const uint32_t buf1,
                                                const uint32_t lane, const bool own, const uint32_t line, const bool use_from,
                                                const bool defer_rev, int32_t* rev_end, int32_t* rev_from) {
  constexpr uint32_t kCharBytes = L8Chars<CM>::kBytes;
  constexpr uint32_t kPer = L8Chars<CM>::kPerChunk;
  const BatchParams& g = p.g;
  const uint8_t* const data = static_cast<const uint8_t*>(g.data);
  uint64_t o0 = 0;
  uint32_t len = 0;
  int32_t from = 0;
  bool slow = false;
  if (own) {
    o0 = batch_off(g, line);
    const uint64_t l64 = batch_off(g, line + 1) - o0;
    slow = l64 >= (1ull << 31);
    len = slow ? 0u : static_cast<uint32_t>(l64);
    if (use_from && !slow) {
      const int32_t f = g.from[line];
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
  const uint64_t sb = o0 * kCharBytes, eb = sb + static_cast<uint64_t>(len) * kCharBytes;  // bytes walked: [sb, eb)
  const uint32_t a = static_cast<uint32_t>(reinterpret_cast<uintptr_t>(data) + sb) & 15u;
  const uint64_t ab = sb - a;
  const uint32_t iters = (len + kPer - 1) / kPer;
  // 16-byte aligned lines (fixed-length records, mostly): walk step k reads chunk k itself - no realignment, no carry
  const bool aligned = __all_sync(0xffffffffu, a == 0 || iters == 0);
  const uint32_t chunks = iters ? iters + (aligned ? 0u : 1u) : 0;  // chunk k = bytes [ab + 16 k, + 16)
  const uint32_t max_pieces = __reduce_max_sync(0xffffffffu, (chunks + 3) / 4);
  // chunks to copy: those that start before the end of the line (one 32-bit bound and one pointer per line, set up once)
  const uint32_t n_copy = static_cast<uint32_t>(min(static_cast<uint64_t>(chunks), (eb - ab + 15) >> 4));
  const uint8_t* const line_base = data + ab;
  uint32_t slot_off[4];
#pragma unroll
  for (uint32_t cc = 0; cc < 4; cc++) slot_off[cc] = l8_slot(lane, cc, 2) << 4;
  // Group staging: four lanes copy one line's 64 bytes of a round (whole 32-byte sectors per request) instead of every lane
  // copying 16 bytes of its own line four times (half a sector per request).  Lane l copies chunk (l & 3) of the lines of lanes
  // (l >> 2) + 8 k; their base (relative to the lowest base of the batch, 32 bits) and copy bound come by shuffle, once per batch.
  const bool has_copies = n_copy != 0;  // (lanes without a line, empty lines: nothing to copy, no say in the base)
  uint64_t base_min = has_copies ? reinterpret_cast<uint64_t>(line_base) : ~0ull;
#pragma unroll
  for (uint32_t d = 16; d > 0; d >>= 1) {
    const uint64_t other = __shfl_xor_sync(0xffffffffu, base_min, d);
    base_min = other < base_min ? other : base_min;
  }
  const uint64_t rel64 = has_copies ? reinterpret_cast<uint64_t>(line_base) - base_min : 0;
  const bool grouped = __all_sync(0xffffffffu, rel64 < (1ull << 32));  // (lines of a batch further apart than 4 GiB: own-line staging)
  uint32_t g_rel[4], g_ncopy[4], g_dst[4];
#pragma unroll
  for (uint32_t k = 0; k < 4; k++) {
    const uint32_t owner = (lane >> 2) + 8 * k;
    g_rel[k] = __shfl_sync(0xffffffffu, static_cast<uint32_t>(rel64), owner);
    g_ncopy[k] = __shfl_sync(0xffffffffu, n_copy, owner);
    g_dst[k] = l8_slot(owner, lane & 3, 2) << 4;
  }
  const uint8_t* const group_base = reinterpret_cast<const uint8_t*>(base_min);
  auto stage = [&](uint32_t j, uint32_t buf) {
    if (grouped) {
      const uint32_t q = 4 * j + (lane & 3);
#pragma unroll
      for (uint32_t k = 0; k < 4; k++)
        if (q < g_ncopy[k]) cp_async16(buf + g_dst[k], group_base + g_rel[k] + 16 * q);
    } else {
#pragma unroll
      for (uint32_t cc = 0; cc < 4; cc++) {
        const uint32_t k = 4 * j + cc;
        if (k < n_copy) cp_async16(buf + slot_off[cc], line_base + 16 * k);
      }
    }
    cp_async_commit();
  };
  uint32_t cur = buf0, nxt = buf1;
  if (max_pieces) stage(0, cur);
  const L8Align al(a);
  uint32_t e = cx.root, pos = 0;
  int32_t last = g.fwd.root_accepting ? 0 : -1;
  uint32_t tail_bit = g.fwd.root_accepting ? 1u : 0u;
  uint4 x = make_uint4(0, 0, 0, 0);
  auto walk_step = [&](const uint4& w) {  // chars [pos, pos + kPer)
    uint32_t mask = 0;
    l8_chunk<CM>(w, p.q, cx, e, mask);
    const uint32_t valid = min(kPer, len - pos);
    mask >>= (kPer - valid);
    const int32_t cand = static_cast<int32_t>(pos + valid + 1) - __ffs(mask);
    last = mask ? cand : last;
    tail_bit = mask & 1u;
    pos += kPer;
    if ((e & L8Enc<CM>::kStateMask) == cx.fwd_dead) pos = len;
  };
  for (uint32_t j = 0; j < max_pieces; j++) {
    if (j + 1 < max_pieces) stage(j + 1, nxt);
    else cp_async_commit();
    cp_async_wait<1>();
    __syncwarp();
    if (aligned) {
#pragma unroll
      for (uint32_t cc = 0; cc < 4; cc++)
        if (4 * j + cc < chunks && pos < len) walk_step(lds_data16(cur + slot_off[cc]));
    } else {
      // A FULL round: every lane that is still walking has all four chunks of the round and the steps they complete are whole
      // 16-byte steps - no per-step bounds, no end-of-line masks, one dead-state test for the round (a dead automaton stays dead
      // and never accepts, so walking on is harmless).  Batches of similar lines (ragged_rounds.cuh) are mostly full rounds.
      const bool act = pos < len;
      const bool full = !act || (4 * j + 4 <= chunks && static_cast<uint64_t>(4 * j + 3) * kPer <= len);
      if (__all_sync(0xffffffffu, full)) {
        if (act) {
#pragma unroll
          for (uint32_t cc = 0; cc < 4; cc++) {
            const uint4 y = lds_data16(cur + slot_off[cc]);
            if (j > 0 || cc > 0) {  // walk step 4 j + cc - 1
              uint32_t mask = 0;
              l8_chunk<CM>(al.apply(x, y), p.q, cx, e, mask);
              const int32_t cand = static_cast<int32_t>(pos + kPer + 1) - __ffs(mask);
              last = mask ? cand : last;
              tail_bit = mask & 1u;
              pos += kPer;
            }
            x = y;
          }
          if ((e & L8Enc<CM>::kStateMask) == cx.fwd_dead) pos = len;
        }
      } else {
#pragma unroll
        for (uint32_t cc = 0; cc < 4; cc++) {
          const uint32_t k = 4 * j + cc;
          if (k < chunks) {
            const uint4 y = lds_data16(cur + slot_off[cc]);
            if (k > 0 && pos < len) walk_step(al.apply(x, y));  // walk step k - 1
            x = y;
          }
        }
      }
    }
    __syncwarp();
    const uint32_t tmp = cur;
    cur = nxt;
    nxt = tmp;
  }
  cp_async_wait<0>();
  bool want_rev = false;
  if (own) {
    if (slow) {
      l8_slow_line<CharT>(g, line);
    } else if (defer_rev && last != -1) {
      g.matched[line] = 1;
      g.end[line] = last + from;
      *rev_end = last + from;
      *rev_from = from;
      want_rev = true;
    } else {
      l8_finish<CM, CharT>(p, cx, line, len, last, tail_bit != 0, 0u, [&](uint32_t) { return buf0; }, from, false);
    }
  }
  __syncwarp();
  return want_rev;
}

// 32 consecutive lines starting at c (the long-line path of the ragged tile walk)
template <int CM, typename CharT>
__device__ __forceinline__ void l8_stream_group(const Lines8Params& p, const L8Ctx& cx, const uint32_t buf0, const uint32_t buf1,
                                                const uint32_t lane, const uint32_t c, const uint32_t m, const bool use_from) {
  int32_t unused_end = 0, unused_from = 0;
  l8_stream_lines<CM, CharT>(p, cx, buf0, buf1, lane, lane < m, c + lane, use_from, false, &unused_end, &unused_from);
}

template <int CM>
__device__ __forceinline__ void l8_run_ragged(const Lines8Params& p, const L8Ctx& cx, const uint32_t buf0, const uint32_t buf1,
                                              const uint32_t lane, const uint32_t warp_global, const uint32_t n_warps) {
  using CharT = typename std::conditional<L8Chars<CM>::kBytes == 1, uint8_t, uint16_t>::type;
  constexpr uint32_t kCharBytes = L8Chars<CM>::kBytes;
  constexpr uint32_t kPer = L8Chars<CM>::kPerChunk;
  const BatchParams& g = p.g;
  const uint8_t* const data = static_cast<const uint8_t*>(g.data);
  const uint32_t n = static_cast<uint32_t>(g.n);
  const uint32_t per_warp = (n + n_warps - 1) / n_warps;
  const uint32_t lo = min(n, warp_global * per_warp), hi = min(n, lo + per_warp);
  constexpr uint32_t kCap = kL8WarpBuf - 16;  // the last 16 bytes stay free for the window that runs past the tile

  struct Plan {
    uint32_t count;   // lines in the tile (0: none left, or the first line alone does not fit)
    uint32_t start;   // this lane's line: first byte, relative to the tile buffer
    uint32_t len;     // this lane's line length in chars
    int32_t from;     // find(from, to): chars to skip at the start of the line; -1: out of range, walk the line from global memory
    bool stream;      // fewer than 8 of the remaining lines fit the buffer: long lines, nothing was staged (l8_stream_group)
  };
  const bool use_from = g.from != nullptr && g.mode == 2;
  // Plan the tile that starts at line c and issue its copies into buf.
  auto plan_and_stage = [&](uint32_t c, uint32_t buf) -> Plan {
    Plan pl;
    pl.count = 0; pl.start = 0; pl.len = 0; pl.from = 0; pl.stream = false;
    if (c < hi) {
      const uint64_t s0 = batch_off(g, c) * kCharBytes;  // bytes
      const uint32_t idx = min(c + lane + 1, hi);
      const uint64_t e = batch_off(g, idx) * kCharBytes;
      const uint32_t slack = static_cast<uint32_t>(reinterpret_cast<uintptr_t>(data) + s0) & 15u;
      const uint64_t rel_end = e - s0 + slack;  // end of this lane's line relative to the tile buffer
      const bool fits = (c + lane < hi) && rel_end <= kCap;
      const uint32_t ballot = __ballot_sync(0xffffffffu, fits);
      pl.count = __popc(ballot);  // offsets are non-decreasing, so `fits` is a prefix
      if (pl.count < 8 && pl.count < hi - c) {
        pl.stream = true;
        pl.count = 0;
        return pl;
      }
      const uint32_t end32 = static_cast<uint32_t>(rel_end);
      uint32_t prev = __shfl_up_sync(0xffffffffu, end32, 1);
      if (lane == 0) prev = slack;
      pl.start = prev;
      pl.len = (end32 - prev) / kCharBytes;
      if (use_from && lane < pl.count) {
        // the walk starts `from` chars in; from outside [0, len) keeps the reference's corner cases (generic walk)
        const int32_t f = g.from[c + lane];
        if (f < 0 || (f != 0 && static_cast<uint32_t>(f) >= pl.len)) {
          pl.from = -1;
        } else {
          pl.from = f;
          pl.start += static_cast<uint32_t>(f) * kCharBytes;
          pl.len -= static_cast<uint32_t>(f);
        }
      }
      if (pl.count) {
        const uint32_t total = __shfl_sync(0xffffffffu, end32, pl.count - 1);
        const uint32_t n_chunks = (total + 15) >> 4;
        const uint8_t* src = data + s0 - slack;
        for (uint32_t j = lane; j < n_chunks; j += 32) cp_async16(buf + l8_rslot(j), src + (static_cast<uint64_t>(j) << 4));
      }
    }
    return pl;
  };

  uint32_t c = lo;
  while (c < hi) {
    const Plan pa = plan_and_stage(c, buf0);
    if (pa.stream) {  // long lines: stream the next 32, one per lane
      const uint32_t m = min(32u, hi - c);
      l8_stream_group<CM, CharT>(p, cx, buf0, buf1, lane, c, m, use_from);
      c += m;
      continue;
    }
    if (pa.count == 0) break;  // (c == hi)
    const uint32_t cb = c + pa.count;
    const Plan pb = plan_and_stage(cb, buf1);  // count 0: nothing left, or a long line that the next round handles
    cp_async_commit();
    // pair the lines: lane k gets rank k of tile A and rank 31 - k of tile B (keys: walk iterations, has-line, lane)
    const bool own_a = lane < pa.count, own_b = lane < pb.count;
    const uint32_t walk_a = (own_a && pa.from >= 0) ? (pa.len + kPer - 1) / kPer : 0u, walk_b = (own_b && pb.from >= 0) ? (pb.len + kPer - 1) / kPer : 0u;
    const uint32_t key_a = warp_sort32(walk_a << 6 | (own_a ? 32u : 0u) | lane, lane);
    const uint32_t key_b = __shfl_sync(0xffffffffu, warp_sort32(walk_b << 6 | (own_b ? 32u : 0u) | lane, lane), 31 - lane);
    const bool has_a = (key_a & 32u) != 0, has_b = (key_b & 32u) != 0;
    const uint32_t start_a = __shfl_sync(0xffffffffu, pa.start, key_a & 31u);
    const uint32_t start_b = __shfl_sync(0xffffffffu, pb.start, key_b & 31u);
    int32_t from_a = 0, from_b = 0;
    if (use_from) {
      from_a = __shfl_sync(0xffffffffu, pa.from, key_a & 31u);
      from_b = __shfl_sync(0xffffffffu, pb.from, key_b & 31u);
    }
    // (a line whose `from` is out of range is not walked here: length 0 in the loop, generic walk afterwards)
    uint32_t len_a = __shfl_sync(0xffffffffu, pa.len, key_a & 31u), len_b = __shfl_sync(0xffffffffu, pb.len, key_b & 31u);
    len_a = from_a < 0 ? 0u : len_a;
    len_b = from_b < 0 ? 0u : len_b;
    cp_async_wait<0>();
    __syncwarp();

    // one loop over both lines of the lane
    uint32_t item = has_a ? 0u : has_b ? 1u : 2u;
    uint32_t base = item == 0 ? buf0 : buf1;
    uint32_t len = item == 0 ? len_a : len_b;
    uint32_t start = item == 0 ? start_a : start_b;
    L8Align al(start & 15u);
    uint32_t c0 = start >> 4, pos = 0;
    uint32_t e = cx.root;
    const int32_t last0 = g.fwd.root_accepting ? 0 : -1;
    const uint32_t tail0 = g.fwd.root_accepting ? 1u : 0u;  // accept flag exactly at the end of the line (matches())
    int32_t last = last0, last_a = last0, last_b = last0;
    uint32_t tail_bit = tail0, tail_a = tail0, tail_b = tail0;
    uint4 x = make_uint4(0, 0, 0, 0);
    if (item < 2) x = lds_data16(base + l8_rslot(c0));
    while (__ballot_sync(0xffffffffu, item < 2) != 0) {
      if (item < 2) {
        if (pos < len) {
          const uint4 y = lds_data16(base + l8_rslot(c0 + (pos / kPer) + 1));
          const uint4 w = al.apply(x, y);
          uint32_t mask = 0;
          l8_chunk<CM>(w, p.q, cx, e, mask);
          const uint32_t valid = min(kPer, len - pos);
          mask >>= (kPer - valid);  // drop the accept bits of chars past the end of the line
          const int32_t cand = static_cast<int32_t>(pos + valid + 1) - __ffs(mask);
          last = mask ? cand : last;
          tail_bit = mask & 1u;
          x = y;
          pos += kPer;
          // a dead automaton stays dead (and never accepts): the rest of the line cannot change the result
          if ((e & L8Enc<CM>::kStateMask) == cx.fwd_dead) pos = len;
        }
        if (pos >= len) {  // this line is done: keep its result, move on to the lane's second line
          if (item == 0) {
            last_a = last;
            tail_a = tail_bit;
          } else {
            last_b = last;
            tail_b = tail_bit;
          }
          item = (item == 0 && has_b) ? 1u : 2u;
          if (item == 1) {
            base = buf1;
            len = len_b;
            al = L8Align(start_b & 15u);
            c0 = start_b >> 4;
            pos = 0;
            e = cx.root;
            last = last0;
            tail_bit = tail0;
            x = lds_data16(base + l8_rslot(c0));
          }
        }
      }
    }
    // Table-driven reverse passes (find() of a variable-length pattern), pooled: only the lines that matched need one,
    // so when there are at most 32 of them in the two tiles they are dealt out one per lane - job j goes to lane j -
    // instead of every lane looping over its own two lines while most lanes idle.
    int32_t st_a = 0, st_b = 0;
    bool pooled = false;
    if (g.mode == 2 && g.reverse_mode == 0 && p.has_bwd != 0) {
      const bool need_a = has_a && from_a >= 0 && last_a != -1, need_b = has_b && from_b >= 0 && last_b != -1;
      const uint32_t jobs_a = __ballot_sync(0xffffffffu, need_a), jobs_b = __ballot_sync(0xffffffffu, need_b);
      const uint32_t n_a = __popc(jobs_a), total = n_a + __popc(jobs_b);
      pooled = total <= 32;  // (when most lines match, every lane has work anyway: each walks its own)
      if (pooled && total) {
        const uint32_t lt = (1u << lane) - 1u;
        const uint32_t my_job_a = __popc(jobs_a & lt), my_job_b = n_a + __popc(jobs_b & lt);  // job numbers of this lane's own lines
        const bool valid = lane < total, from_b_tile = lane >= n_a;
        const uint32_t src = valid ? __fns(from_b_tile ? jobs_b : jobs_a, 0, static_cast<int>(from_b_tile ? lane - n_a : lane) + 1) : 0u;
        const uint32_t ps_a = __shfl_sync(0xffffffffu, start_a, src), ps_b = __shfl_sync(0xffffffffu, start_b, src);
        const int32_t l_a = __shfl_sync(0xffffffffu, last_a, src), l_b = __shfl_sync(0xffffffffu, last_b, src);
        int32_t st = 0;
        if (valid) {
          const uint32_t buf = from_b_tile ? buf1 : buf0;
          st = l8_reverse<CM>(p, [&](uint32_t ch) { return buf + l8_rslot(ch); }, from_b_tile ? ps_b : ps_a, from_b_tile ? l_b : l_a, cx,
                              g.bwd.root_accepting != 0);
        }
        // hand the results back to the lanes that own the lines
        st_a = __shfl_sync(0xffffffffu, st, my_job_a & 31u);
        st_b = __shfl_sync(0xffffffffu, st, my_job_b & 31u);
      }
    }
    if (has_a) {
      if (from_a < 0) l8_slow_line<CharT>(g, c + (key_a & 31u));
      else l8_finish<CM, CharT>(p, cx, c + (key_a & 31u), len_a, last_a, tail_a != 0, start_a, [&](uint32_t ch) { return buf0 + l8_rslot(ch); }, from_a,
                                true, pooled, st_a);
    }
    if (has_b) {
      if (from_b < 0) l8_slow_line<CharT>(g, cb + (key_b & 31u));
      else l8_finish<CM, CharT>(p, cx, cb + (key_b & 31u), len_b, last_b, tail_b != 0, start_b, [&](uint32_t ch) { return buf1 + l8_rslot(ch); }, from_b,
                                true, pooled, st_b);
    }
    __syncwarp();  // every lane is done with both buffers before the next pair of tiles overwrites them
    c = cb + pb.count;
  }
}


// ---------------------------------------------------------------------------------------------
// Iterated find() on the staged tile: all non-overlapping matches of every line - the loop
// `while (m.find()) { m.start(); m.end(); }` (DFACompilerTest.java:678-699), find() resuming at nextStart = end
// of the previous match and the reverse pass bounded below by it (DFAClassBuilder.java:625-659).  One line per lane,
// tiles double-buffered as in l8_run; the lane loops over its line's matches, every forward scan starting wherever the
// last match ended (any byte offset: realigned windows as in the ragged walk), the reverse pass on the same staged bytes.
// A match that does not move nextStart forward would repeat forever in the reference (:634-635): reported once, ends the line.
// counts[i] = matches of line i; with match_offsets, match k is stored at match_offsets[i] + k (two-pass CSR).
// ---------------------------------------------------------------------------------------------
template <int CM>
__device__ __forceinline__ void l8_find_all(const Lines8Params& p, const L8Ctx& cx, const uint32_t buf0, const uint32_t buf1,
                                            const uint32_t lane, const uint32_t warp_global, const uint32_t n_warps) {
  using CharT = typename std::conditional<L8Chars<CM>::kBytes == 1, uint8_t, uint16_t>::type;
  constexpr uint32_t kCharBytes = L8Chars<CM>::kBytes;
  constexpr uint32_t kPer = L8Chars<CM>::kPerChunk;
  constexpr uint32_t kFull = 0xffffffffu;
  const BatchParams& g = p.g;
  const uint8_t* const data = static_cast<const uint8_t*>(g.data);
  const uint32_t n = static_cast<uint32_t>(g.n);
  const uint32_t per_warp = (n + n_warps - 1) / n_warps;
  const uint32_t lo = min(n, warp_global * per_warp), hi = min(n, lo + per_warp);
  constexpr uint32_t kCap = kL8WarpBuf - 16;  // the last 16 bytes stay free for the window that runs past the tile

  struct Plan {
    uint32_t count;  // lines in the tile; 0: the next line does not fit a buffer (generic walk)
    uint32_t start;  // this lane's line: first byte, relative to the tile buffer
    uint32_t len;    // its length in chars
  };
  auto plan_and_stage = [&](uint32_t c, uint32_t buf) -> Plan {
    Plan pl;
    pl.count = 0; pl.start = 0; pl.len = 0;
    if (c < hi) {
      const uint64_t s0 = batch_off(g, c) * kCharBytes;  // bytes
      const uint32_t idx = min(c + lane + 1, hi);
      const uint64_t e_off = batch_off(g, idx) * kCharBytes;
      const uint32_t slack = static_cast<uint32_t>(reinterpret_cast<uintptr_t>(data) + s0) & 15u;
      const uint64_t rel_end = e_off - s0 + slack;
      const bool fits = (c + lane < hi) && rel_end <= kCap;
      pl.count = __popc(__ballot_sync(kFull, fits));  // offsets are non-decreasing, so `fits` is a prefix
      if (pl.count) {
        const uint32_t end32 = static_cast<uint32_t>(rel_end);
        uint32_t prev = __shfl_up_sync(kFull, end32, 1);
        if (lane == 0) prev = slack;
        pl.start = prev;
        pl.len = (end32 - prev) / kCharBytes;
        const uint32_t total = __shfl_sync(kFull, end32, pl.count - 1);
        const uint32_t n_chunks = (total + 15) >> 4;
        const uint8_t* src = data + s0 - slack;
        for (uint32_t j = lane; j < n_chunks; j += 32) cp_async16(buf + l8_rslot(j), src + (static_cast<uint64_t>(j) << 4));
      }
    }
    cp_async_commit();
    return pl;
  };
  // matches of line i, handed to the caller's arrays
  auto emit = [&](uint32_t i, uint32_t k, uint64_t out, uint64_t cap, int32_t st, int32_t en) {
    (void)i;
    if (k < cap) {
      g.start[out + k] = st;
      g.end[out + k] = en;
    }
  };
  // (lines longer than a tile buffer run the whole loop straight from global memory: dev_find_all_line)
  auto generic_line = [&](uint32_t i) { dev_find_all_line<CharT>(g, i); };

  uint32_t c = lo, cur = buf0, nxt = buf1;
  Plan t = plan_and_stage(c, cur);
  while (c < hi) {
    if (t.count == 0) {  // a line that does not fit a buffer: it and up to 31 followers take the generic loop, one per lane
      cp_async_wait<0>();
      const uint32_t m = min(32u, hi - c);
      if (lane < m) generic_line(c + lane);
      __syncwarp();
      c += m;
      t = plan_and_stage(c, cur);
      continue;
    }
    const uint32_t cn = c + t.count;
    const Plan tn = plan_and_stage(cn, nxt);  // (commits a group even when nothing is staged)
    cp_async_wait<1>();
    __syncwarp();
    // every line of the tile starts on a 16-byte boundary (fixed-length records, mostly): the FIRST search of every line - the
    // only one for lines without a match - reads its chunks as they are, no realignment
    const bool tile_aligned = __all_sync(kFull, lane >= t.count || (t.start & 15u) == 0);
    if (lane < t.count) {
      const uint32_t i = c + lane;
      const uint32_t len = t.len;
      uint64_t out = 0, cap = 0;
      if (g.match_offsets) {
        out = g.match_offsets[i];
        cap = g.match_offsets[i + 1] - out;
      }
      auto chunk_addr = [&](uint32_t ch) { return cur + l8_rslot(ch); };
      uint32_t count = 0, from = 0;
      bool first = tile_aligned;
      for (;;) {
        // indexForwards(from): DFAClassBuilder.java:335-471
        int32_t last = g.fwd.root_accepting ? (from < len ? static_cast<int32_t>(from) : 0) : -1;
        const uint32_t ps = t.start + from * kCharBytes;
        uint32_t ch = ps >> 4, pos = from, e = cx.root;
        if (first) {
          first = false;
          while (pos < len) {
            const uint4 wv = lds_data16(chunk_addr(ch++));
            uint32_t mask = 0;
            l8_chunk<CM>(wv, p.q, cx, e, mask);
            const uint32_t valid = min(kPer, len - pos);
            mask >>= (kPer - valid);  // drop the accept bits of chars past the end of the line
            const int32_t cand = static_cast<int32_t>(pos + valid + 1) - __ffs(mask);
            last = mask ? cand : last;
            pos += kPer;
            if ((e & L8Enc<CM>::kStateMask) == cx.fwd_dead) break;
          }
        } else {
          const L8Align al(ps & 15u);
          uint4 x = lds_data16(chunk_addr(ch));
          while (pos < len) {
            const uint4 y = lds_data16(chunk_addr(++ch));
            const uint4 wv = al.apply(x, y);
            uint32_t mask = 0;
            l8_chunk<CM>(wv, p.q, cx, e, mask);
            const uint32_t valid = min(kPer, len - pos);
            mask >>= (kPer - valid);  // drop the accept bits of chars past the end of the line
            const int32_t cand = static_cast<int32_t>(pos + valid + 1) - __ffs(mask);
            last = mask ? cand : last;
            x = y;
            pos += kPer;
            if ((e & L8Enc<CM>::kStateMask) == cx.fwd_dead) break;
          }
        }
        if (last == -1) break;
        // start(): DFAClassBuilder.java:640-659, lower bound = from
        int32_t st;
        const int32_t rel = last - static_cast<int32_t>(from);
        if (g.reverse_mode == 2) {
          st = last - g.min_length;
        } else if (g.reverse_mode == 0 && p.has_bwd) {
          st = l8_reverse<CM>(p, chunk_addr, ps, rel, cx, g.bwd.root_accepting != 0);
          if (st != 0x7fffffff) st += static_cast<int32_t>(from);
        } else if (g.reverse_mode == 1) {
          st = l8_reverse_char<CM>(chunk_addr, ps, rel, g.reverse_char);
          if (st != 0x7fffffff) st += static_cast<int32_t>(from);
        } else {
          st = static_cast<int32_t>(dev_index_backwards<CharT>(g, static_cast<const CharT*>(g.data) + batch_off(g, i), last - 1, from, 0x7fffffff));
        }
        emit(i, count, out, cap, st, last);
        count++;
        if (last <= static_cast<int32_t>(from)) break;  // nextStart did not advance: the reference would repeat this match forever
        from = static_cast<uint32_t>(last);
      }
      g.counts[i] = count;
    }
    __syncwarp();  // every lane is done with `cur` before the stage after next overwrites it
    c = cn;
    const uint32_t tmp = cur;
    cur = nxt;
    nxt = tmp;
    t = tn;
  }
  cp_async_wait<0>();
}

}  // namespace ndl
#include "ragged_rounds.cuh"
namespace ndl {

// ---------------------------------------------------------------------------------------------
// Fixed-length lines of any multiple of 16 bytes, walked in ROUNDS: a tile is always 32 lines - one per lane - and a round
// stages the next 64 bytes of each of them (coalesced: four lanes copy one line's 64 bytes; the slots of a 64-byte-line tile)
// while the previous round is walked.  The resident-tile walks above hold 2048 / L lines per tile, so 96-byte records
// keep 21 lanes busy, 128-byte records 16 and 256-byte records 8; here every lane is busy for any record length.  The line
// is not resident when the walk ends, so a table-driven reverse pass reads global memory (l8_finish, resident = false).
// ---------------------------------------------------------------------------------------------
template <int CM, bool kOffsets, bool kPartial>
__device__ __forceinline__ void l8_run_rounds(const Lines8Params& p, const L8Ctx& cx, const uint32_t buf0, const uint32_t buf1,
                                              const uint32_t lane, const uint32_t warp_global, const uint32_t n_warps, const uint32_t cpl,
                                              const uint32_t line_lo) {
  // lines [line_lo, n) in tiles of 32.  kPartial = false: the full tiles (then the leftover lines through the kPartial = true
  // instance, whose tile may hold fewer than 32 lines - its other lanes sit out; kept apart so that the hot loop carries no
  // per-copy line-count tests)
  using CharT = typename std::conditional<L8Chars<CM>::kBytes == 1, uint8_t, uint16_t>::type;
  constexpr uint32_t kCharBytes = L8Chars<CM>::kBytes;
  constexpr uint32_t kPer = L8Chars<CM>::kPerChunk;
  constexpr uint32_t kFlush = 32 / kPer;  // chunks whose accept bits fit in the 32-bit mask (2 or 4: a round is a multiple)
  const BatchParams& g = p.g;
  const uint8_t* const data = static_cast<const uint8_t*>(g.data);
  const uint32_t n = static_cast<uint32_t>(g.n);
  const uint32_t line_bytes = 16u * cpl, len_chars = line_bytes / kCharBytes;
  const uint32_t rounds = (cpl + 3) / 4;
  const uint32_t n_tiles = kPartial ? (n - line_lo + 31) / 32 : (n - line_lo) / 32;
  // this lane's four copies of a round: chunk c = lane + 32 k is part (c & 3) of line (c >> 2)
  uint32_t dst_off[4], src_off[4];
#pragma unroll
  for (uint32_t k = 0; k < 4; k++) {
    const uint32_t c = lane + 32 * k;
    dst_off[k] = l8_slot(c >> 2, c & 3, 2) << 4;
    src_off[k] = (c >> 2) * line_bytes + (c & 3) * 16;
  }
  const uint32_t part = lane & 3;  // (the same for the lane's four copies)

  auto lines_in = [&](uint32_t tile) { return kPartial ? min(32u, n - line_lo - tile * 32) : 32u; };
  auto load_offsets = [&](uint32_t tile, uint64_t& o0, uint64_t& o1) {
    const uint32_t i = line_lo + tile * 32 + min(lane, lines_in(tile) - 1);  // (lanes without a line repeat the last one)
    if constexpr (kOffsets) {
      o0 = g.offsets[i];
      o1 = g.offsets[i + 1];
    } else {
      o0 = static_cast<uint64_t>(i) * g.line_chars;
      o1 = o0 + g.line_chars;
    }
  };
  // are the tile's lines equally spaced and 16-byte aligned?  src: first byte of the tile (the same on every lane if so)
  auto check = [&](uint32_t tile, uint64_t o0, uint64_t o1, const uint8_t*& src) -> bool {
    src = data + (o0 * kCharBytes - static_cast<uint64_t>(min(lane, lines_in(tile) - 1)) * line_bytes);
    const bool ok = (o1 - o0 == len_chars) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
    return __all_sync(0xffffffffu, ok) != 0;
  };
  auto stage = [&](const uint8_t* src, uint32_t count, uint32_t r, uint32_t buf) {
    if (4 * r + part < cpl) {
#pragma unroll
      for (uint32_t k = 0; k < 4; k++)
        if (((lane + 32 * k) >> 2) < count) cp_async16(buf + dst_off[k], src + 64 * r + src_off[k]);
    }
    cp_async_commit();
  };

  const bool defer_rev = g.mode == 2 && g.reverse_mode == 0 && p.has_bwd != 0;  // find() with the table-driven reverse pass
  L8RevQueue<CM> queue;
  uint32_t t = warp_global;
  uint32_t cur = buf0, nxt = buf1;
  bool regular = false;
  const uint8_t* src = data;
  uint64_t a0 = 0, a1 = 0;  // offsets of the tile after the current one
  if (t < n_tiles) {
    load_offsets(t, a0, a1);
    regular = check(t, a0, a1, src);
    if (regular) stage(src, lines_in(t), 0, cur);
    else cp_async_commit();
    if (t + n_warps < n_tiles) load_offsets(t + n_warps, a0, a1);
  }
  for (; t < n_tiles; t += n_warps) {
    const uint32_t count = lines_in(t);
    const uint32_t i = line_lo + t * 32 + lane;
    uint32_t e = cx.root, mask = 0;
    int32_t last = g.fwd.root_accepting ? 0 : -1;
    bool regular_next = false;
    const uint8_t* src_next = data;
    for (uint32_t r = 0; r < rounds; r++) {
      if (r + 1 < rounds) {
        if (regular) stage(src, count, r + 1, nxt);
        else cp_async_commit();
      } else if (t + n_warps < n_tiles) {  // the first round of this warp's next tile
        regular_next = check(t + n_warps, a0, a1, src_next);
        if (regular_next) stage(src_next, lines_in(t + n_warps), 0, nxt);
        else cp_async_commit();
        if (t + 2 * n_warps < n_tiles) load_offsets(t + 2 * n_warps, a0, a1);
      } else {
        cp_async_commit();
      }
      cp_async_wait<1>();
      __syncwarp();
      if (regular && lane < count) {
#pragma unroll
        for (uint32_t cc = 0; cc < 4; cc++) {
          const uint32_t c = 4 * r + cc;
          if (c < cpl) {
            const uint4 w = lds_data16(cur + (l8_slot(lane, cc, 2) << 4));
            l8_chunk<CM>(w, p.q, cx, e, mask);
            if ((c % kFlush) == kFlush - 1 || c + 1 == cpl) {  // bit 0 = the most recent char
              const int32_t cand = static_cast<int32_t>((c + 1) * kPer + 1) - __ffs(mask);
              last = mask ? cand : last;
              mask = 0;
            }
          }
        }
      }
      __syncwarp();  // every lane is done with `cur` before the stage after next overwrites it
      const uint32_t tmp = cur;
      cur = nxt;
      nxt = tmp;
    }
    bool want_rev = false;
    if (lane < count) {
      if (!regular) {
        l8_slow_line<CharT>(g, i);
      } else if (defer_rev && last != -1) {  // the reverse pass of a line that matched: queued, 32 at a time (L8RevQueue)
        g.matched[i] = 1;
        g.end[i] = last;
        want_rev = true;
      } else {
        l8_finish<CM, CharT>(p, cx, i, len_chars, last, (e & L8Enc<CM>::kTailFlag) != 0, 0u, [&](uint32_t) { return buf0; }, 0, false);
      }
    }
    if (defer_rev) {
      const uint32_t rv = __ballot_sync(0xffffffffu, want_rev);
      if (rv) queue.push(p, cx, lane, rv, i, last, 0);
    }
    regular = regular_next;
    src = src_next;
  }
  cp_async_wait<0>();
  queue.flush(p, cx, lane);
  if constexpr (!kPartial) {
    // the leftover lines go to the warp that is next in the round-robin of tiles (it has one tile fewer than the first warps)
    if ((n - line_lo) % 32 != 0)
      l8_run_rounds<CM, kOffsets, true>(p, cx, buf0, buf1, lane, (warp_global + n_warps - n_tiles % n_warps) % n_warps, n_warps, cpl,
                                        line_lo + n_tiles * 32);
  }
}

// The rounds walk for fixed-length lines of ANY byte length (100-byte records ...): the lines of a tile start at different
// offsets within their 16-byte chunks, so a round copies the aligned chunks that cover the next 64 bytes of every line (addresses
// by arithmetic - the spacing is known - still four lanes per line), and every lane realigns its own line in registers as the
// ragged walk does: walk step s needs chunks s and s + 1, so it runs when chunk s + 1 has arrived; the lower chunk is carried.
template <int CM, bool kOffsets, bool kPartial = false>
__device__ __forceinline__ void l8_run_rounds_unaligned(const Lines8Params& p, const L8Ctx& cx, const uint32_t buf0, const uint32_t buf1,
                                                        const uint32_t lane, const uint32_t warp_global, const uint32_t n_warps,
                                                        const uint32_t line_bytes, const uint32_t line_lo = 0) {
  using CharT = typename std::conditional<L8Chars<CM>::kBytes == 1, uint8_t, uint16_t>::type;
  constexpr uint32_t kCharBytes = L8Chars<CM>::kBytes;
  constexpr uint32_t kPer = L8Chars<CM>::kPerChunk;
  const BatchParams& g = p.g;
  const uint8_t* const data = static_cast<const uint8_t*>(g.data);
  const uint32_t n = static_cast<uint32_t>(g.n);
  const uint32_t len_chars = line_bytes / kCharBytes;
  const uint32_t steps = (len_chars + kPer - 1) / kPer;  // walk steps per line; step s reads chunks s and s + 1 of the line
  const uint32_t rounds = (steps + 1 + 3) / 4;
  const uint32_t n_tiles = kPartial ? (n - line_lo + 31) / 32 : (n - line_lo) / 32;  // (kPartial: as in l8_run_rounds)
  uint32_t dst_off[4];
#pragma unroll
  for (uint32_t k = 0; k < 4; k++) {
    const uint32_t c = lane + 32 * k;
    dst_off[k] = l8_slot(c >> 2, c & 3, 2) << 4;
  }
  const uint32_t part = lane & 3;

  auto lines_in = [&](uint32_t tile) { return kPartial ? min(32u, n - line_lo - tile * 32) : 32u; };
  auto load_offsets = [&](uint32_t tile, uint64_t& o0, uint64_t& o1) {
    const uint32_t i = line_lo + tile * 32 + min(lane, lines_in(tile) - 1);  // (lanes without a line repeat the last one)
    if constexpr (kOffsets) {
      o0 = g.offsets[i];
      o1 = g.offsets[i + 1];
    } else {
      o0 = static_cast<uint64_t>(i) * g.line_chars;
      o1 = o0 + g.line_chars;
    }
  };
  // are the tile's lines equally spaced?  base: first byte of the tile's first line (the same on every lane if so)
  auto check = [&](uint32_t tile, uint64_t o0, uint64_t o1, const uint8_t*& base) -> bool {
    base = data + (o0 * kCharBytes - static_cast<uint64_t>(min(lane, lines_in(tile) - 1)) * line_bytes);
    return __all_sync(0xffffffffu, o1 - o0 == len_chars) != 0;
  };
  auto stage = [&](const uint8_t* base, uint32_t count, uint32_t r, uint32_t buf) {
    const uint32_t q = 4 * r + part;  // chunk of the line this lane copies (for four lines)
#pragma unroll
    for (uint32_t k = 0; k < 4; k++) {
      const uint32_t j = (lane + 32 * k) >> 2;
      const uint8_t* line = base + j * line_bytes;
      const uint32_t a = static_cast<uint32_t>(reinterpret_cast<uintptr_t>(line)) & 15u;
      if (j < count && q < ((a + line_bytes + 15u) >> 4)) cp_async16(buf + dst_off[k], line - a + 16 * q);  // (the chunk starts before the line ends)
    }
    cp_async_commit();
  };

  const bool defer_rev = g.mode == 2 && g.reverse_mode == 0 && p.has_bwd != 0;  // find() with the table-driven reverse pass
  L8RevQueue<CM> queue;
  uint32_t t = warp_global;
  uint32_t cur = buf0, nxt = buf1;
  bool regular = false;
  const uint8_t* base = data;
  uint64_t a0 = 0, a1 = 0;  // offsets of the tile after the current one
  if (t < n_tiles) {
    load_offsets(t, a0, a1);
    regular = check(t, a0, a1, base);
    if (regular) stage(base, lines_in(t), 0, cur);
    else cp_async_commit();
    if (t + n_warps < n_tiles) load_offsets(t + n_warps, a0, a1);
  }
  for (; t < n_tiles; t += n_warps) {
    const uint32_t count = lines_in(t);
    const uint32_t i = line_lo + t * 32 + lane;
    const uint32_t a = static_cast<uint32_t>(reinterpret_cast<uintptr_t>(base + lane * line_bytes)) & 15u;
    const uint32_t chunks = (a + line_bytes + 15u) >> 4;  // chunks that hold bytes of this lane's line
    const L8Align al(a);
    uint32_t e = cx.root, tail_bit = g.fwd.root_accepting ? 1u : 0u;
    int32_t last = g.fwd.root_accepting ? 0 : -1;
    uint4 x = make_uint4(0, 0, 0, 0);
    bool regular_next = false;
    const uint8_t* base_next = data;
    for (uint32_t r = 0; r < rounds; r++) {
      if (r + 1 < rounds) {
        if (regular) stage(base, count, r + 1, nxt);
        else cp_async_commit();
      } else if (t + n_warps < n_tiles) {  // the first round of this warp's next tile
        regular_next = check(t + n_warps, a0, a1, base_next);
        if (regular_next) stage(base_next, lines_in(t + n_warps), 0, nxt);
        else cp_async_commit();
        if (t + 2 * n_warps < n_tiles) load_offsets(t + 2 * n_warps, a0, a1);
      } else {
        cp_async_commit();
      }
      cp_async_wait<1>();
      __syncwarp();
      if (regular && lane < count) {
#pragma unroll
        for (uint32_t cc = 0; cc < 4; cc++) {
          const uint32_t q = 4 * r + cc;
          if (q <= steps) {
            const uint4 y = q < chunks ? lds_data16(cur + (l8_slot(lane, cc, 2) << 4)) : make_uint4(0, 0, 0, 0);
            if (q > 0) {  // walk step q - 1: chars [(q - 1) kPer, q kPer)
              uint32_t mask = 0;
              l8_chunk<CM>(al.apply(x, y), p.q, cx, e, mask);
              const uint32_t pos = (q - 1) * kPer;
              const uint32_t valid = min(kPer, len_chars - pos);
              mask >>= (kPer - valid);  // drop the accept bits of chars past the end of the line
              const int32_t cand = static_cast<int32_t>(pos + valid + 1) - __ffs(mask);
              last = mask ? cand : last;
              tail_bit = mask & 1u;
            }
            x = y;
          }
        }
      }
      __syncwarp();  // every lane is done with `cur` before the stage after next overwrites it
      const uint32_t tmp = cur;
      cur = nxt;
      nxt = tmp;
    }
    bool want_rev = false;
    if (lane < count) {
      if (!regular) {
        l8_slow_line<CharT>(g, i);
      } else if (defer_rev && last != -1) {  // the reverse pass of a line that matched: queued, 32 at a time (L8RevQueue)
        g.matched[i] = 1;
        g.end[i] = last;
        want_rev = true;
      } else {
        l8_finish<CM, CharT>(p, cx, i, len_chars, last, tail_bit != 0, 0u, [&](uint32_t) { return buf0; }, 0, false);
      }
    }
    if (defer_rev) {
      const uint32_t rv = __ballot_sync(0xffffffffu, want_rev);
      if (rv) queue.push(p, cx, lane, rv, i, last, 0);
    }
    regular = regular_next;
    base = base_next;
  }
  cp_async_wait<0>();
  queue.flush(p, cx, lane);
  if constexpr (!kPartial) {
    if ((n - line_lo) % 32 != 0)
      l8_run_rounds_unaligned<CM, kOffsets, true>(p, cx, buf0, buf1, lane, (warp_global + n_warps - n_tiles % n_warps) % n_warps, n_warps, line_bytes,
                                                  line_lo + n_tiles * 32);
  }
}

// Which walk a batch takes, from the byte length L64 of its first line (every tile re-checks its own lines):
//   0..4                fixed-length, 16 << k bytes, whole tile resident (l8_run)
//   kCplAny + cpl       fixed-length, 16 * cpl bytes, whole tile resident with a run-time chunk count (l8_run_any)
//   kCplRounds + cpl    fixed-length, 16 * cpl bytes, 32 lines per tile walked in rounds of 64 bytes (l8_run_rounds)
//   kCplRoundsU + L     fixed-length, L bytes (not a multiple of 16), the same with per-line alignment (l8_run_rounds_unaligned)
//   -1                  ragged
// Lines of 16, 32, 48 and 64 bytes fill all 32 lanes of a resident tile.  Longer ones do not (2048 / L lines per tile), so they
// are walked in rounds - measured faster for every pattern kind, also when find() then runs its table-driven reverse pass from
// global memory instead of the tile (e-mail regex on 256-byte records: 3.75 TB/s against 1.28 with 8 lines per tile).
constexpr int kCplAny = 100;
constexpr int kCplRounds = 1000;
constexpr int kCplRoundsU = 100000;         // + line bytes: fixed-length lines of any byte length, rounds with per-line alignment
constexpr uint64_t kMinRoundsUnaligned = 17;
constexpr uint32_t kMaxRoundsCpl = 4096;  // records up to 64 KB (longer ones: the ragged walk streams them)
__device__ __forceinline__ int l8_pick_geometry(const Lines8Params& p, uint64_t L64) {
  const BatchParams& g = p.g;
  if (g.from != nullptr && g.mode == 2) return -1;  // find(from, to): the ragged walk takes the per-line start offsets
  if (L64 >= kMinRoundsUnaligned && L64 <= 16ull * p.rounds_max_cpl && (L64 & 15) != 0 && p.no_rounds == 0) return kCplRoundsU + static_cast<int>(L64);
  if (L64 < 16 || (L64 & 15) != 0 || (L64 >> 4) > p.rounds_max_cpl) return -1;
  const uint32_t cpl = static_cast<uint32_t>(L64 >> 4);
  if (cpl == 1 || cpl == 2 || cpl == 4) return 31 - __clz(cpl);
  if (cpl == 3) return kCplAny + 3;
  if (p.no_rounds != 0) return cpl < 8 ? kCplAny + static_cast<int>(cpl) : -1;  // (experiments, NDL_NO_ROUNDS: resident tiles / ragged)
  return kCplRounds + static_cast<int>(cpl);
}

template <int CM>
__device__ __forceinline__ void l8_dispatch(const Lines8Params& p, const L8Ctx& cx, int log2cpl, uint32_t buf0, uint32_t buf1,
                                            uint32_t lane, uint32_t warp_global, uint32_t n_warps) {
  if (p.g.mode == kModeFindAll) {
    l8_find_all<CM>(p, cx, buf0, buf1, lane, warp_global, n_warps);
    return;
  }
  const bool off = p.g.offsets != nullptr;
  if (log2cpl >= kCplRoundsU) {
    if (off) l8_run_rounds_unaligned<CM, true>(p, cx, buf0, buf1, lane, warp_global, n_warps, static_cast<uint32_t>(log2cpl - kCplRoundsU));
    else l8_run_rounds_unaligned<CM, false>(p, cx, buf0, buf1, lane, warp_global, n_warps, static_cast<uint32_t>(log2cpl - kCplRoundsU));
    return;
  }
  if (log2cpl >= kCplRounds) {
    if (off) l8_run_rounds<CM, true>(p, cx, buf0, buf1, lane, warp_global, n_warps, static_cast<uint32_t>(log2cpl - kCplRounds), 0u);
    else l8_run_rounds<CM, false>(p, cx, buf0, buf1, lane, warp_global, n_warps, static_cast<uint32_t>(log2cpl - kCplRounds), 0u);
    return;
  }
  if (log2cpl >= kCplAny) {
    if (off) l8_run_any<CM, true>(p, cx, buf0, buf1, lane, warp_global, n_warps, static_cast<uint32_t>(log2cpl - kCplAny));
    else l8_run_any<CM, false>(p, cx, buf0, buf1, lane, warp_global, n_warps, static_cast<uint32_t>(log2cpl - kCplAny));
    return;
  }
  switch (log2cpl) {
#define NDL_RUN(L)                                                               \
  case L:                                                                        \
    if (off) l8_run<L, CM, true>(p, cx, buf0, buf1, lane, warp_global, n_warps); \
    else l8_run<L, CM, false>(p, cx, buf0, buf1, lane, warp_global, n_warps);    \
    break;
    NDL_RUN(0) NDL_RUN(1) NDL_RUN(2)  // (128- and 256-byte records: the rounds walk)
#undef NDL_RUN
    default: {
      // ragged: longer lines (mean length from the first and the last offset) take the sorted streaming walk
      const uint64_t bytes = (batch_off(p.g, p.g.n) - batch_off(p.g, 0)) * L8Chars<CM>::kBytes;
      const bool p_from = p.g.from != nullptr && p.g.mode == 2;
      // (on shorter lines - mean 64 bytes - it is within +-6 % of the tile walk, depending on the pattern: they keep the tile walk)
      if (p.g.n >= kRrMinLines && bytes >= static_cast<uint64_t>(p.rr_min_mean) * p.g.n && p.no_rounds == 0 && !p_from)
        l8_run_ragged_rounds<CM>(p, cx, buf0, buf1, lane, warp_global, n_warps);
      else
        l8_run_ragged<CM>(p, cx, buf0, buf1, lane, warp_global, n_warps);
      break;
    }
  }
}

// Where a warp's two tile buffers live.  Buffer slots (2 KB) fill the space below the class map and an
// "upper" region that depends on the layout; warp w owns slots 2w and 2w+1.  Warps without a pair of
// slots (one warp in the S1 layout) do not take part.
struct L8Setup {
  bool layout_ok, warp_ok;
  uint32_t buf0, buf1, usable_warps, cmap_bytes, abs_trans;
  __device__ __forceinline__ L8Setup(bool s1, uint32_t warp) {
    extern __shared__ __align__(128) uint8_t l8_dyn_smem[];
    const uint32_t base = static_cast<uint32_t>(__cvta_generic_to_shared(l8_dyn_smem));
    const uint32_t buf_a = (base + 127) & ~127u;
    layout_ok = buf_a <= 0x8000;
    cmap_bytes = s1 ? kS1CmapBytes : kL8CmapBytes;
    abs_trans = s1 ? kS1AbsTrans : kL8AbsTrans;
    const uint32_t upper_lo = s1 ? kS1UpperLo : kL8AbsSet1;
    const uint32_t upper_hi = s1 ? kL8AbsBar : kL8AbsTrans;
    const uint32_t fit = layout_ok ? (kL8AbsCmap - buf_a) / kL8WarpBuf : 0;
    const uint32_t slots = fit + (upper_hi - upper_lo) / kL8WarpBuf;
    usable_warps = min(static_cast<uint32_t>(kL8Warps), slots / 2);
    warp_ok = layout_ok && warp < usable_warps;
    auto slot = [&](uint32_t k) { return k < fit ? buf_a + k * kL8WarpBuf : upper_lo + (k - fit) * kL8WarpBuf; };
    buf0 = slot(2 * warp);
    buf1 = slot(2 * warp + 1);
  }
};

#ifdef NDL_MAIN_TU  // non-template kernels are compiled by capi_device.cu only (the inst_*.cu files share this header)
__global__ void __launch_bounds__(kL8Threads, 1) lines8_kernel(const Lines8Params p) {
  const uint32_t tid = threadIdx.x;
  const uint32_t lane = tid & 31, warp = tid >> 5;
  const BatchParams& g = p.g;

  L8Setup su(p.char_mode == kCmBytes1, warp);
  const bool layout_ok = su.layout_ok, warp_ok = su.warp_ok;
  const uint32_t buf0 = su.buf0, buf1 = su.buf1;

  // --- table image: two TMA bulk copies (cmap, trans) completing on one mbarrier
  if (tid == 0) {
    mbar_init(kL8AbsBar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  if (tid == 0 && layout_ok) {
    mbar_expect_tx(kL8AbsBar, su.cmap_bytes + p.trans_bytes);
    tma_bulk_g2s(kL8AbsCmap, p.image, su.cmap_bytes, kL8AbsBar);
    tma_bulk_g2s(su.abs_trans, p.image + su.cmap_bytes, p.trans_bytes, kL8AbsBar);
  }

  // --- line geometry: byte length from the first two offsets (uniform); every tile re-checks its own lines
  const uint32_t char_bytes = (p.char_mode == kCmBytes || p.char_mode == kCmBytes1 || p.char_mode == kCmBytesH) ? 1u : 2u;
  const uint64_t l_chars = batch_off(g, 1) - batch_off(g, 0);
  const uint64_t L64 = l_chars * char_bytes;
  int log2cpl = l8_pick_geometry(p, L64);
  if (layout_ok) mbar_wait(kL8AbsBar, 0);  // table image has landed

  const uint32_t warp_global = blockIdx.x * su.usable_warps + warp;
  const uint32_t n_warps = gridDim.x * su.usable_warps;
  // The fixed-length path is taken when the batch starts with 33 equally spaced offsets of a supported
  // length (its tiles still re-check themselves); everything else goes down the ragged path.
  if (log2cpl >= 0) {
    const uint32_t probe = static_cast<uint32_t>(min(static_cast<uint64_t>(lane) + 1, g.n - 1));
    const bool same = batch_off(g, probe + 1) - batch_off(g, probe) == l_chars;
    if (!__all_sync(0xffffffffu, same)) log2cpl = -1;
  }
  if (!layout_ok) {
    // unexpected shared-memory base: generic walk, one line per thread
    for (uint64_t i = (static_cast<uint64_t>(blockIdx.x) * kL8Warps + warp) * 32 + lane; i < g.n;
         i += static_cast<uint64_t>(gridDim.x) * kL8Threads) {
      if (char_bytes == 1) l8_slow_line<uint8_t>(g, i);
      else l8_slow_line<uint16_t>(g, i);
    }
    return;
  }
  if (!warp_ok) return;  // no pair of tile buffers for this warp in this layout
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
  cx.root = p.root_entry;
  cx.bwd_root = p.bwd_root;
  cx.bwd_dead = p.bwd_dead;
  cx.fwd_dead = p.fwd_dead;
  if (p.char_mode == kCmBytesH) l8_dispatch<kCmBytesH>(p, cx, log2cpl, buf0, buf1, lane, warp_global, n_warps);
  else if (p.char_mode == kCmBytes1) l8_dispatch<kCmBytes1>(p, cx, log2cpl, buf0, buf1, lane, warp_global, n_warps);
  else if (p.char_mode == kCmBytes) l8_dispatch<kCmBytes>(p, cx, log2cpl, buf0, buf1, lane, warp_global, n_warps);
  else if (p.char_mode == kCmHi) l8_dispatch<kCmHi>(p, cx, log2cpl, buf0, buf1, lane, warp_global, n_warps);
  else l8_dispatch<kCmMixed>(p, cx, log2cpl, buf0, buf1, lane, warp_global, n_warps);
}

#endif  // NDL_MAIN_TU

// ---------------------------------------------------------------------------------------------
// linesq_kernel: the SWAR modes.  Same tiles, same walks (l8_run / l8_run_ragged), other table image:
// [kQAbsTrans, +trans_bytes) transition table, brought in by TMA bulk copies, then 2 KB tile buffers.
// One instantiation per char mode; 28 warps x 72 registers (a full-size table leaves tile buffers for 21-23 of them).
// ---------------------------------------------------------------------------------------------
template <int CM>
__global__ void __launch_bounds__(kQThreads, 1) linesq_kernel(const Lines8Params p) {
  const uint32_t tid = threadIdx.x;
  const uint32_t lane = tid & 31, warp = tid >> 5;
  const BatchParams& g = p.g;
  using CharT = typename std::conditional<L8Chars<CM>::kBytes == 1, uint8_t, uint16_t>::type;

  extern __shared__ __align__(128) uint8_t l8_dyn_smem[];
  const uint32_t base = static_cast<uint32_t>(__cvta_generic_to_shared(l8_dyn_smem));
  const bool layout_ok = base <= kQAbsTrans;
  const uint32_t tiles_lo = (kQAbsTrans + p.trans_bytes + 127u) & ~127u;
  const uint32_t slots = (kL8AbsBar - tiles_lo) / kL8WarpBuf;
  const uint32_t usable_warps = min(static_cast<uint32_t>(kQWarps), slots / 2);
  const uint32_t buf0 = tiles_lo + 2 * warp * kL8WarpBuf, buf1 = buf0 + kL8WarpBuf;

  if (tid == 0) {
    mbar_init(kL8AbsBar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  if (tid == 0 && layout_ok) {
    mbar_expect_tx(kL8AbsBar, p.trans_bytes);
    for (uint32_t off = 0; off < p.trans_bytes; off += 0x8000u)
      tma_bulk_g2s(kQAbsTrans + off, p.image + off, min(0x8000u, p.trans_bytes - off), kL8AbsBar);
  }
  const uint64_t l_chars = batch_off(g, 1) - batch_off(g, 0);
  const uint64_t L64 = l_chars * L8Chars<CM>::kBytes;
  int log2cpl = l8_pick_geometry(p, L64);
  if (log2cpl >= 0) {
    const uint32_t probe = static_cast<uint32_t>(min(static_cast<uint64_t>(lane) + 1, g.n - 1));
    const bool same = batch_off(g, probe + 1) - batch_off(g, probe) == l_chars;
    if (!__all_sync(0xffffffffu, same)) log2cpl = -1;
  }
  if (!layout_ok) {  // unexpected shared-memory base: generic walk, one line per thread
    for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * kQThreads + tid; i < g.n; i += static_cast<uint64_t>(gridDim.x) * kQThreads)
      l8_slow_line<CharT>(g, i);
    return;
  }
  mbar_wait(kL8AbsBar, 0);  // table image has landed
  if (warp >= usable_warps) return;
  const uint32_t lane_off = (lane & p.q.copy_mask) * p.q.copy_bytes;
  L8Ctx cx;
  cx.sel_a = cx.sel_b = cx.page1 = cx.page3 = cx.ua = cx.ub = cx.xa = cx.xb = cx.row_bytes = 0;
  cx.root = p.root_entry + lane_off;
  cx.bwd_root = p.bwd_root + lane_off;
  cx.bwd_dead = p.bwd_dead + lane_off;
  cx.fwd_dead = p.fwd_dead + lane_off;
  l8_dispatch<CM>(p, cx, log2cpl, buf0, buf1, lane, blockIdx.x * usable_warps + warp, gridDim.x * usable_warps);
}

}  // namespace ndl
