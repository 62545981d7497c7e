// Host side of the tile kernels: builds the shared-memory table images (lines8.cuh: "L" images with a class map,
// pair / stride-1 tables; "Q" images of the packed-compare modes) from the device automaton of a pattern
// (device_image.h).  Used by ndl_pattern_create and by the host-only image emulation hook (capi_device.cu).
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <utility>
#include <vector>

#include "../device_image.h"
#include "lines8.cuh"
#include "swar_plan.h"

namespace ndl {

// Build the shared-memory image for 2-char steps: the forward automaton of the mode and, optionally, the
// BACKWARDS automaton (find() of a variable-length pattern) behind it.  Both share one pair of class maps:
// columns are the distinct (forward class, backward class) combinations the class-map slots take.
// Returns false when the class map has no supported char mode or the pair tables do not fit (the generic
// kernel handles the pattern then).
inline bool lines8_layout(const HostDeviceTable& f, const HostDeviceTable* b, int char_width, bool allow_unreplicated,
                          std::vector<uint8_t>& img, Lines8Blob& meta, bool entries16 = false) {
  // entries16 (byte haystacks, one plain copy): 16-bit entries = byte offset of the target row / 2 | flags << 14 (kCmBytesH)
  using Key = std::pair<int, int>;
  auto key_of = [&](int c) { return Key{f.cmap[c], b ? b->cmap[c] : 0}; };
  Key slot_key[256];
  Key uniform_key{0, 0};
  int char_mode = kCmBytes, mixed_page = -1;
  if (char_width == 1) {
    for (int v = 0; v < 256; v++) slot_key[v] = key_of(v);
  } else {
    std::vector<int> mixed;
    for (int hi = 0; hi < 256; hi++) {
      // U+FFFF is always treated as an exception char (its class is 0 in every reference class map,
      // DFA.java:451), so it does not make page 0xFF "mixed"
      bool uniform = true;
      for (int lo = 1; lo < 256 && uniform; lo++)
        if ((hi << 8 | lo) != 0xFFFF) uniform = key_of(hi << 8 | lo) == key_of(hi << 8);
      if (!uniform) mixed.push_back(hi);
    }
    if (mixed.empty()) {
      char_mode = kCmHi;
      for (int hi = 0; hi < 256; hi++) slot_key[hi] = key_of(hi << 8);
    } else if (mixed.size() == 1) {
      mixed_page = mixed[0];
      const int other = mixed_page == 0 ? 1 : 0;
      uniform_key = key_of(other << 8);
      for (int hi = 0; hi < 256; hi++)
        if (hi != mixed_page && !(key_of(hi << 8) == uniform_key)) return false;
      char_mode = kCmMixed;
      for (int lo = 0; lo < 256; lo++) slot_key[lo] = key_of(mixed_page << 8 | lo);
    } else {
      return false;
    }
  }
  std::vector<Key> col_classes;
  auto col_for = [&](const Key& k) {
    for (size_t i = 0; i < col_classes.size(); i++)
      if (col_classes[i] == k) return static_cast<int>(i);
    col_classes.push_back(k);
    return static_cast<int>(col_classes.size() - 1);
  };
  int col_of_slot[256];
  for (int v = 0; v < 256; v++) col_of_slot[v] = col_for(slot_key[v]);
  const int col_uniform = char_mode == kCmMixed ? col_for(uniform_key) : 0;
  const int col_exc = char_mode != kCmBytes ? col_for(key_of(0xFFFF)) : 0;
  const int C = static_cast<int>(col_classes.size());
  const int rows_f = f.n_states + 1, rows_b = b ? b->n_states + 1 : 0;
  const long pairs = static_cast<long>(rows_f + rows_b) * C * C;
  int R;
  if (entries16) {
    if (char_mode != kCmBytes || !allow_unreplicated || pairs * 2 > static_cast<long>(kL8MaxTransBytes) || pairs * 2 > 0x8000) return false;
    R = 1;
  } else if (pairs * 128 <= static_cast<long>(kL8MaxTransBytes)) {
    R = 32;
  } else if (allow_unreplicated && pairs * 4 <= static_cast<long>(kL8MaxTransBytes)) {
    R = 1;
  } else {
    return false;
  }
  const uint32_t col_bytes = entries16 ? 2u : 4u * R;
  const uint32_t row_bytes = static_cast<uint32_t>(C) * C * col_bytes;
  const uint32_t trans_bytes = (static_cast<uint32_t>(rows_f + rows_b) * row_bytes + 15) & ~15u;
  img.assign(kL8CmapBytes + trans_bytes, 0);
  auto put = [&](uint32_t off, uint32_t v) { std::memcpy(img.data() + off, &v, 4); };
  for (int v = 0; v < 256; v++) {
    const uint32_t c = static_cast<uint32_t>(col_of_slot[v]);
    for (uint32_t lane = 0; lane < 32; lane++) {
      put(v * 256 + lane * 4, c * C * col_bytes);                                              // CA
      put(v * 256 + 128 + lane * 4, kL8AbsTrans + c * col_bytes + (R == 32 ? lane * 4 : 0));   // CB
    }
  }
  auto emit = [&](const HostDeviceTable& t, int row0, bool backward) {
    const int rows = t.n_states + 1;
    for (int s = 0; s < rows; s++)
      for (int c1 = 0; c1 < C; c1++) {
        const int k1 = backward ? col_classes[c1].second : col_classes[c1].first;
        const int s1 = t.trans[static_cast<size_t>(s) * t.n_classes + k1];
        for (int c2 = 0; c2 < C; c2++) {
          const int k2 = backward ? col_classes[c2].second : col_classes[c2].first;
          const int s2 = t.trans[static_cast<size_t>(s1) * t.n_classes + k2];
          const uint32_t at = kL8CmapBytes + static_cast<uint32_t>(row0 + s) * row_bytes + (static_cast<uint32_t>(c1) * C + c2) * col_bytes;
          if (entries16) {
            const uint16_t e16 = static_cast<uint16_t>((static_cast<uint32_t>(row0 + s2) * row_bytes) >> 1 | (t.accept[s1] ? 0x8000u : 0) |
                                                       (t.accept[s2] ? 0x4000u : 0));
            std::memcpy(img.data() + at, &e16, 2);
            continue;
          }
          const uint32_t e = static_cast<uint32_t>(row0 + s2) * row_bytes | (t.accept[s1] ? 0x80000000u : 0) | (t.accept[s2] ? 0x40000000u : 0);
          for (int lane = 0; lane < R; lane++) put(at + lane * 4, e);
        }
      }
  };
  emit(f, 0, false);
  const uint32_t state_unit = entries16 ? row_bytes / 2 : row_bytes;  // what a table entry holds per row
  meta.root_entry = 0;  // forward root = row 0
  meta.fwd_dead = static_cast<uint32_t>(f.n_states) * state_unit;
  meta.has_bwd = b != nullptr;
  if (b) {
    emit(*b, rows_f, true);
    meta.bwd_root = static_cast<uint32_t>(rows_f) * state_unit;
    meta.bwd_dead = static_cast<uint32_t>(rows_f + b->n_states) * state_unit;
  }
  meta.trans_bytes = trans_bytes;
  meta.replicated = R;
  meta.n_cols = C;
  meta.row_bytes = row_bytes;
  meta.char_mode = entries16 ? static_cast<int>(kCmBytesH) : char_mode;
  meta.mixed_page = mixed_page < 0 ? 0 : mixed_page;
  meta.ua = static_cast<uint32_t>(col_uniform) * C * col_bytes;
  meta.ub = kL8AbsTrans + static_cast<uint32_t>(col_uniform) * col_bytes;  // + lane*4 in the kernel when replicated
  meta.xa = static_cast<uint32_t>(col_exc) * C * col_bytes;
  meta.xb = kL8AbsTrans + static_cast<uint32_t>(col_exc) * col_bytes;
  return true;
}

// The S1 image (byte haystacks): [cmap 32 KB][trans u16], stride-1 steps.  `copies` = 32: every entry replicated
// per lane (64 bytes per entry, lane l reads halfword l).  `copies` = 1: a plain [row][column] array of 16-bit
// entries - about 3.5 wavefronts per lookup (32 lanes over 32 banks at random), but it holds automata of thousands
// of states (the large-table path of SURVEY.md 8f-3) and is still 2-3 times the generic kernel, whose tables live
// in global memory.  With `b`, the BACKWARDS rows follow the forward rows and share the class map (joint columns),
// so the reverse pass of find() runs on the staged tile as well.  The kernel code is the same for both: the class
// map yields the address of (row 0, column) for the lane, the entry the row index | accept << 15.
inline bool lines8_layout_s1(const HostDeviceTable& f, const HostDeviceTable* b, int copies, std::vector<uint8_t>& img, Lines8Blob& meta) {
  using Key = std::pair<int, int>;
  std::vector<Key> col_classes;
  int col_of_byte[256];
  for (int v = 0; v < 256; v++) {
    const Key k{f.cmap[v], b ? b->cmap[v] : 0};
    int c = -1;
    for (size_t i = 0; i < col_classes.size(); i++)
      if (col_classes[i] == k) c = static_cast<int>(i);
    if (c < 0) {
      col_classes.push_back(k);
      c = static_cast<int>(col_classes.size() - 1);
    }
    col_of_byte[v] = c;
  }
  const int C = static_cast<int>(col_classes.size());
  const int rows_f = f.n_states + 1, rows_b = b ? b->n_states + 1 : 0, rows = rows_f + rows_b;
  if (rows > 0x7fff) return false;
  const uint32_t entry_bytes = copies == 32 ? 64u : 2u;  // bytes per (row, column)
  const uint32_t row_bytes = static_cast<uint32_t>(C) * entry_bytes;
  const uint64_t want = static_cast<uint64_t>(rows) * row_bytes;
  if (want > kS1MaxTransBytes) return false;
  const uint32_t trans_bytes = (static_cast<uint32_t>(want) + 15) & ~15u;
  img.assign(kS1CmapBytes + trans_bytes, 0);
  for (int v = 0; v < 256; v++)
    for (uint32_t lane = 0; lane < 32; lane++) {
      const uint32_t cm = kS1AbsTrans + static_cast<uint32_t>(col_of_byte[v]) * entry_bytes + (copies == 32 ? lane * 2u : 0u);
      std::memcpy(img.data() + v * 128 + lane * 4, &cm, 4);
    }
  auto emit = [&](const HostDeviceTable& t, int row0, bool backward) {
    for (int s = 0; s <= t.n_states; s++)
      for (int c = 0; c < C; c++) {
        const int k = backward ? col_classes[c].second : col_classes[c].first;
        const int nx = t.trans[static_cast<size_t>(s) * t.n_classes + k];
        const uint16_t e = static_cast<uint16_t>((row0 + nx) | (t.accept[nx] ? 0x8000 : 0));
        const uint32_t at = kS1CmapBytes + static_cast<uint32_t>(row0 + s) * row_bytes + static_cast<uint32_t>(c) * entry_bytes;
        for (uint32_t lane = 0; lane < (copies == 32 ? 32u : 1u); lane++) std::memcpy(img.data() + at + lane * 2u, &e, 2);
      }
  };
  emit(f, 0, false);
  meta = Lines8Blob();
  if (b) {
    emit(*b, rows_f, true);
    meta.has_bwd = true;
    meta.bwd_root = static_cast<uint32_t>(rows_f);
    meta.bwd_dead = static_cast<uint32_t>(rows_f + b->n_states);
  }
  meta.trans_bytes = trans_bytes;
  meta.root_entry = 0;
  meta.fwd_dead = static_cast<uint32_t>(f.n_states);
  meta.replicated = copies;
  meta.n_cols = C;
  meta.row_bytes = row_bytes;
  meta.char_mode = kCmBytes1;
  return true;
}

// The Q image (SWAR modes): only a transition table.  Entry (row, column) of copy q lives at
//   column * kmul * 128 + (row / W) * 128 + q * 4W + (row % W) * 4,      W = 32 / R banks per copy
// (16-bit entries: 2 instead of 4 bytes each, W = 64 / R entries per copy and line)
// so that with R = 32 every lane has a private bank (conflict free) and with fewer copies the 32 / R lanes
// that share a copy spread over its W banks by row number.  An entry is the absolute shared-memory address
// of the target row's slot in the same copy, with the accept flags of the K steps in the top K bits
// (bit 31 = after the first char).  Columns are numbered c1 * n^(K-1) + ... + cK over the n codes of the plan.
// Preference order = estimated wavefronts per char (lookups per char x expected bank-conflict degree).
inline bool linesq_layout(const HostDeviceTable& f, const HostDeviceTable* b, int char_width, std::vector<uint8_t>& img,
                          Lines8Blob& meta) {
  using Key = std::pair<int, int>;
  // classes whose table columns are identical (e.g. "in no range" and "above maxChar", both dead) are one class
  auto canonical = [](const HostDeviceTable& t) {
    std::vector<int> canon(t.n_classes);
    for (int k = 0; k < t.n_classes; k++) {
      canon[k] = k;
      for (int j = 0; j < k && canon[k] == k; j++) {
        bool same = true;
        for (int s = 0; s <= t.n_states && same; s++)
          same = t.trans[static_cast<size_t>(s) * t.n_classes + k] == t.trans[static_cast<size_t>(s) * t.n_classes + j];
        if (same) canon[k] = canon[j];
      }
    }
    return canon;
  };
  const std::vector<int> canon_f = canonical(f), canon_b = b ? canonical(*b) : std::vector<int>();
  auto key_of = [&](int c) { return Key{canon_f[f.cmap[c]], b ? canon_b[b->cmap[c]] : 0}; };
  std::vector<Key> classes;
  auto id_of = [&](const Key& k) {
    for (size_t i = 0; i < classes.size(); i++)
      if (classes[i] == k) return static_cast<int>(i);
    classes.push_back(k);
    return static_cast<int>(classes.size() - 1);
  };
  // slot domain: byte values; for UTF-16 the high byte when every 256-char page is uniform, else the code unit itself
  // (16-bit lanes: thresholds below 0x8000, everything above - U+FFFF included, whose class is always 0 - one class)
  std::vector<int> key;
  bool hi = false, wide = false;
  if (char_width == 1) {
    key.resize(256);
    for (int v = 0; v < 256; v++) key[v] = id_of(key_of(v));
  } else {
    hi = true;
    for (int h = 0; h < 256 && hi; h++)
      for (int lo = 1; lo < 256 && hi; lo++) hi = key_of(h << 8 | lo) == key_of(h << 8);
    if (hi) {
      key.resize(256);
      for (int h = 0; h < 256; h++) key[h] = id_of(key_of(h << 8));
    } else {
      wide = true;
      key.resize(65536);
      for (int c = 0; c < 65536; c++) key[c] = id_of(key_of(c));
    }
  }
  auto char_of_slot = [&](int slot) { return hi ? slot << 8 : slot; };
  SwarPlan plan;
  if (!swar_solve(key, 6, plan)) return false;
  const int n = plan.n_codes;
  const int rows_f = f.n_states + 1, rows_b = b ? b->n_states + 1 : 0, rows = rows_f + rows_b;
  // candidates in order of estimated shared-memory wavefronts per char = (expected conflict degree of one
  // lookup: 1 with 32 copies, about 2 with 16, about 3 with 8, measured) / K
  struct Cand { int k, r, bytes; };
  // Preference, measured on 10 M x 64-byte lines (exp/layout_ab.py, exp/warps_ab.sh):
  //  * every 4-char table beats every 2-char table: the 2-char walk is issue-bound (23 instructions per 4 chars
  //    against 15-16): a[ab]{k}c 4.85 TB/s with (4 chars, 16-bit entries, 1 copy) against 3.90 / 3.85 / 3.72 TB/s
  //    with (2 chars, 32 / 16 / 8 copies);
  //  * a smaller table leaves room for more warps, which matters more than conflict-free lookups: C2 with 28 warps
  //    5.27 TB/s on 16 copies (2 wavefronts per lookup) against 5.02 on 32 copies (1 wavefront, but only 21 warps
  //    have tile buffers), and 5.14 on a single copy of 16-bit entries (about 3.5 wavefronts);
  //  * 16-bit entries hold a (16 - K)-bit row address + K flags, so at K = 4 one copy of a 1000-row automaton fits;
  //    four copies of them measured worse than one on small automata (all rows of a column share one 128-byte line).
  static const Cand kCands[] = {{4, 16, 4}, {4, 8, 4}, {4, 1, 2}, {2, 32, 4}, {2, 16, 4}, {2, 8, 4}, {2, 8, 2}};
  constexpr int kNCands = sizeof(kCands) / sizeof(kCands[0]);
  int K = 0, R = 0, W = 0, EB = 4, lines_per_col = 0;
  uint32_t n_cols = 0;
  uint32_t col_stride = 0;
  // experiments only: NDL_Q_FORCE="k,copies,entry bytes" restricts the choice to one candidate
  int force_k = 0, force_r = 0, force_b = 0;
  if (const char* f = std::getenv("NDL_Q_FORCE")) std::sscanf(f, "%d,%d,%d", &force_k, &force_r, &force_b);
  for (int ci = 0; ci < kNCands; ci++) {
    const Cand& c = kCands[ci];
    if (force_k && (c.k != force_k || c.r != force_r || c.bytes != force_b)) continue;
    if (c.bytes == 2 && (hi || wide)) continue;  // 16-bit entries: byte haystacks only (kernel instantiations)
    long cols = 1;
    for (int i = 0; i < c.k; i++) cols *= n;
    int vmax = 0;
    for (int p = 0; p < plan.planes; p++) vmax = plan.val[p] > vmax ? plan.val[p] : vmax;
    if (vmax * (cols / n) > 255) continue;  // IDP.4A weights are bytes
    const int w = 128 / c.bytes / c.r;      // entries of one copy per 128-byte line
    const long lpc = (rows + w - 1) / w;
    // bytes per column: whole 128-byte lines, except for a single copy, whose rows are simply contiguous
    const long stride = c.r == 1 ? ((static_cast<long>(rows) * c.bytes + 3) & ~3L) : lpc * 128;
    if (cols * stride > static_cast<long>(kQMaxTransBytes)) continue;
    // a 16-bit entry holds the row address in 16 - K bits
    if (c.bytes == 2 && kQAbsTrans + (c.r == 1 ? stride : lpc * 128) > (0x10000L >> c.k)) continue;
    K = c.k; R = c.r; W = w; EB = c.bytes; lines_per_col = static_cast<int>(lpc); n_cols = static_cast<uint32_t>(cols);
    col_stride = static_cast<uint32_t>(stride);
    break;
  }
  if (K == 0) return false;
  const uint32_t trans_bytes = (n_cols * col_stride + 15u) & ~15u;
  img.assign(trans_bytes, 0);
  auto slot_off = [&](uint32_t row, uint32_t copy) {  // offset of a row's slot inside a column
    return (row / W) * 128u + copy * static_cast<uint32_t>(EB * W) + (row % W) * static_cast<uint32_t>(EB);
  };
  auto emit = [&](const HostDeviceTable& t, int row0, bool backward) {
    const int rows_t = t.n_states + 1;
    std::vector<int> cls(n);  // table class of every code (unused codes behave like code 0)
    for (int c = 0; c < n; c++) {
      const int slot = plan.slot_of_code[c] >= 0 ? plan.slot_of_code[c] : plan.slot_of_code[0] >= 0 ? plan.slot_of_code[0] : 128;
      cls[c] = t.cmap[char_of_slot(slot)];
    }
    (void)backward;
    for (int s = 0; s < rows_t; s++)
      for (uint32_t col = 0; col < n_cols; col++) {
        int st = s;
        uint32_t flags = 0, rem = col, div = n_cols;
        for (int i = 0; i < K; i++) {
          div /= n;
          const int code = static_cast<int>(rem / div);
          rem %= div;
          st = t.trans[static_cast<size_t>(st) * t.n_classes + cls[code]];
          if (t.accept[st]) flags |= 0x80000000u >> i;
        }
        for (int q = 0; q < R; q++) {
          const uint32_t target = kQAbsTrans + slot_off(static_cast<uint32_t>(row0 + st), q);
          const uint32_t at = col * col_stride + slot_off(static_cast<uint32_t>(row0 + s), q);
          if (EB == 4) {
            const uint32_t v = target | flags;
            std::memcpy(img.data() + at, &v, 4);
          } else {
            const uint16_t v = static_cast<uint16_t>(target | flags >> 16);
            std::memcpy(img.data() + at, &v, 2);
          }
        }
      }
  };
  emit(f, 0, false);
  meta = Lines8Blob();
  meta.has_bwd = b != nullptr;
  if (b) {
    emit(*b, rows_f, true);
    meta.bwd_root = kQAbsTrans + slot_off(static_cast<uint32_t>(rows_f), 0);
    meta.bwd_dead = kQAbsTrans + slot_off(static_cast<uint32_t>(rows_f + b->n_states), 0);
  }
  meta.root_entry = kQAbsTrans;
  meta.fwd_dead = kQAbsTrans + slot_off(static_cast<uint32_t>(f.n_states), 0);
  meta.trans_bytes = trans_bytes;
  meta.replicated = R;
  meta.n_cols = n;
  meta.row_bytes = 0;
  meta.char_mode = wide ? cm_swar_wide(K, plan.planes) : cm_swar(K, plan.planes, hi, EB == 2);
  SwarDev& q = meta.q;
  for (int p = 0; p < 3; p++) {
    const bool on = p < plan.planes;
    const uint32_t rep = wide ? 0x00010001u : 0x01010101u;
    q.lo[p] = on ? static_cast<uint32_t>(plan.lo[p]) * rep : 0;
    q.hi[p] = on ? static_cast<uint32_t>(plan.hi[p]) * rep : 0;
    const uint32_t v = on ? static_cast<uint32_t>(plan.val[p]) : 0;
    const uint32_t un = static_cast<uint32_t>(n);
    if (wide && K == 4) {  // two chars per word (bytes 1 and 3), two words per lookup
      q.w[p][0] = (v * un * un * un) << 8 | (v * un * un) << 24;
      q.w[p][1] = (v * un) << 8 | v << 24;
      q.w[p][2] = (v * un * un * un) << 24 | (v * un * un) << 8;
      q.w[p][3] = (v * un) << 24 | v << 8;
    } else if (wide) {
      q.w[p][0] = (v * un) << 8 | v << 24;
      q.w[p][2] = (v * un) << 24 | v << 8;
      q.w[p][1] = q.w[p][3] = 0;
    } else if (K == 4) {
      q.w[p][0] = v * un * un * un | (v * un * un) << 8 | (v * un) << 16 | v << 24;
      q.w[p][2] = v | (v * un) << 8 | (v * un * un) << 16 | (v * un * un * un) << 24;
      q.w[p][1] = q.w[p][3] = 0;
    } else {
      q.w[p][0] = v * un | v << 8;
      q.w[p][1] = (v * un) << 16 | v << 24;
      q.w[p][2] = v << 16 | (v * un) << 24;
      q.w[p][3] = v | (v * un) << 8;
    }
  }
  q.kmul = (EB == 2 && K == 4) ? col_stride : static_cast<uint32_t>(lines_per_col);  // bytes / 128-byte lines per column
  q.copy_mask = static_cast<uint32_t>(R - 1);
  q.copy_bytes = static_cast<uint32_t>(EB * W);
  return true;
}

}  // namespace ndl
