// Host side of the SWAR classifier: decides whether a char -> class map can be evaluated with packed
// compares instead of table lookups, and if so with which constants.
//
// The reference looks the class of every char up in BYTE_CLASSES (DFAClassBuilder.java:269-305).  On the GPU
// that lookup is one shared-memory wavefront per char and is what bounds the table-driven kernels
// (profiles/r01_ncu_lines8_v3_summary.txt: 1.85 wavefronts per warp-char, LSU pipe 79 % busy).  Regex class
// maps are unions of a few char ranges (DFA.java:438-463 builds them from sorted RangeGroups), so for many
// patterns a class is decided by a handful of comparisons, and comparisons on four packed bytes cost the same
// as on one:
//     ge(t)   = ((w | 0x80808080) - t * 0x01010101) : bit 7 of each byte = (byte & 0x7f) >= t,  t in [0, 128]
//     plane_p = (ge(lo_p) ^ ge(hi_p)) & ~w & 0x80808080 : byte in [lo_p, hi_p) and byte < 0x80
// A "plan" is up to three such ranges with small integer values v_p; the column code of a byte is the sum of
// the values of the ranges it lies in (ranges may overlap), bytes >= 0x80 have code 0.  The kernels never
// materialise the code: a dot product (IDP.4A) of each plane with per-position weights gives the column
// offset of a whole group of 2 or 4 chars directly.
#pragma once
#include <cstdint>
#include <vector>

namespace ndl {

struct SwarPlan {
  int planes = 0;          // 1..3
  int lo[3] = {0, 0, 0};   // range p is [lo, hi), 0 <= lo < hi <= span
  int hi[3] = {0, 0, 0};
  int val[3] = {0, 0, 0};  // value added to the code inside the range
  int n_codes = 0;         // codes are 0 .. n_codes - 1 (some may be unused)
  int slot_of_code[8];     // a representative slot value per code, -1 when the code is unused
};

// key[v]: an arbitrary class id per slot value v (byte value, high byte of a UTF-16 char, or - 16-bit lanes -
// the UTF-16 code unit itself); key.size() = 2 * span with span = 128 or 32768.  Succeeds when all slots
// >= span (the top bit of the lane set) share one key and the slots below fall into at most `max_codes`
// classes expressible with <= 3 ranges.  Picks the plan with the fewest codes, then the fewest planes.
inline bool swar_solve(const std::vector<int>& key, int max_codes, SwarPlan& out) {
  const int span = static_cast<int>(key.size() / 2);
  const int z = key[span];
  for (int v = span + 1; v < 2 * span; v++)
    if (key[v] != z) return false;
  // maximal runs of equal key over [0, span)
  std::vector<int> bound{0};  // run j is [bound[j], bound[j + 1])
  std::vector<int> run_key;
  for (int v = 1; v < span; v++)
    if (key[v] != key[v - 1]) {
      bound.push_back(v);
      if (bound.size() > 10) return false;
    }
  bound.push_back(span);
  const int m = static_cast<int>(bound.size()) - 1;
  for (int j = 0; j < m; j++) run_key.push_back(key[bound[j]]);
  if (m > 9) return false;

  struct Range { int s, t; };
  std::vector<Range> ranges;
  for (int s = 0; s < m; s++)
    for (int t = s + 1; t <= m; t++) ranges.push_back({s, t});
  bool found = false;
  int best_codes = 0, best_planes = 0;
  SwarPlan best;
  int code[16];
  auto consider = [&](const Range* rs, const int* vals, int planes) {
    int max_code = 0;
    for (int j = 0; j < m; j++) {
      int c = 0;
      for (int p = 0; p < planes; p++)
        if (rs[p].s <= j && j < rs[p].t) c += vals[p];
      code[j] = c;
      if (c > max_code) max_code = c;
    }
    if (max_code + 1 > max_codes) return;
    for (int j = 0; j < m; j++) {
      if ((run_key[j] == z) != (code[j] == 0)) return;
      for (int i = 0; i < j; i++)
        if ((run_key[i] == run_key[j]) != (code[i] == code[j])) return;
    }
    const int n_codes = max_code + 1;
    if (found && (n_codes > best_codes || (n_codes == best_codes && planes >= best_planes))) return;
    found = true;
    best_codes = n_codes;
    best_planes = planes;
    best = SwarPlan();
    best.planes = planes;
    best.n_codes = n_codes;
    for (int p = 0; p < planes; p++) {
      best.lo[p] = bound[rs[p].s];
      best.hi[p] = bound[rs[p].t];
      best.val[p] = vals[p];
    }
    for (int c = 0; c < 8; c++) best.slot_of_code[c] = -1;
    for (int j = 0; j < m; j++)
      if (best.slot_of_code[code[j]] < 0) best.slot_of_code[code[j]] = bound[j];
    if (best.slot_of_code[0] < 0) best.slot_of_code[0] = span;  // code 0 is at least the class of the top half
  };
  // a map with a single class: one plane that never fires
  if (m == 1 && run_key[0] == z) {
    Range r{0, 1};
    // [0, 128) would fire on every ASCII byte; use an empty test instead: lo == hi is not allowed, so
    // give the range value 0 semantics by choosing code 0 for it through val = 0
    int v0 = 0;
    consider(&r, &v0, 1);
    if (found) {
      out = best;
      return true;
    }
  }
  const int nr = static_cast<int>(ranges.size());
  Range rs[3];
  int vals[3];
  for (int a = 0; a < nr; a++)
    for (int va = 1; va <= 3; va++) {
      rs[0] = ranges[a];
      vals[0] = va;
      consider(rs, vals, 1);
      for (int b = a + 1; b < nr; b++)
        for (int vb = 1; vb <= 3; vb++) {
          rs[1] = ranges[b];
          vals[1] = vb;
          consider(rs, vals, 2);
        }
    }
  if (!found)  // three planes only when two do not suffice (the search is 50x larger)
    for (int a = 0; a < nr; a++)
      for (int b = a + 1; b < nr; b++)
        for (int c = b + 1; c < nr; c++)
          for (int va = 1; va <= 3; va++)
            for (int vb = 1; vb <= 3; vb++)
              for (int vc = 1; vc <= 3; vc++) {
                rs[0] = ranges[a]; rs[1] = ranges[b]; rs[2] = ranges[c];
                vals[0] = va; vals[1] = vb; vals[2] = vc;
                consider(rs, vals, 3);
              }
  if (!found) return false;
  out = best;
  return true;
}

}  // namespace ndl
