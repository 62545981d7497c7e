// Device automaton: what the kernels walk.  Derived from the reference-layout CompiledPattern once, at
// ndl_pattern_create time (the analogue of the generated class's <clinit>, DFAClassBuilder.java:184-333).
//
// The generated Java loops interleave three kinds of special-casing with the table walk; each is folded
// into the tables here so the per-char step on the GPU is a pure lookup `state = T[state][class(c)]`:
//
//  * dead state.  The reference stores -1 and tests `state == -1` after every step.  Here DEAD is a real,
//    absorbing, non-accepting row (index n_states), so lanes that died keep stepping harmlessly.
//  * `c > maxChar` checks (matches :899-901, containedIn :1012-1016, indexForwards :451-457,
//    indexBackwards :573-575).  Chars above the table's maxChar map to one extra class column K:
//      MATCHES / FORWARDS / BACKWARDS:  T[s][K] = DEAD        ("return false" / "return lastMatch")
//      CONTAINEDIN:                     T[s][K] = 0           ("state = 0; index++; break")
//  * containedIn's restart and early return (:1000-1009): a dead entry means "continue from state 0 at
//    the next char", and an accepting state returns true before the next char is read, so dead entries
//    become 0 and accepting rows are made absorbing; the answer is accept[final state].
//
// The useMaxStart loop bound (index <= length - minLength, :360-364, :975-979) only stops scans that can
// no longer fit a match; with exact tables it cannot change a result and is not represented.
#pragma once
#include <cstdint>
#include <vector>

#include "host/pattern.h"

namespace ndl {

struct HostDeviceTable {
  int n_states = 0;     // real states; DEAD == n_states
  int n_classes = 0;    // reference stride + 1 (the extra "above maxChar" column is index n_classes - 1)
  bool root_accepting = false;
  std::vector<uint16_t> cmap;    // 65536: char -> class column
  std::vector<uint16_t> trans;   // (n_states + 1) * n_classes -> next state
  std::vector<uint8_t> accept;   // n_states + 1
};

inline HostDeviceTable build_device_table(const CompiledPattern& p, int table_id) {
  const Table& t = p.tables[table_id];
  HostDeviceTable d;
  d.n_states = t.n_states;
  d.n_classes = p.stride + 1;
  const int K = p.stride;
  const int dead = t.n_states;
  const bool contained = table_id == kContainedIn;
  d.root_accepting = t.accepting[0] != 0;
  d.cmap.resize(65536);
  for (int c = 0; c < 65536; c++) d.cmap[c] = (c > t.max_char) ? static_cast<uint16_t>(K) : p.class_map[c];
  d.accept.assign(t.n_states + 1, 0);
  for (int s = 0; s < t.n_states; s++) d.accept[s] = t.accepting[s];
  d.trans.assign(static_cast<size_t>(t.n_states + 1) * d.n_classes, static_cast<uint16_t>(dead));
  for (int s = 0; s < t.n_states; s++) {
    for (int k = 0; k <= K; k++) {
      int next;
      if (contained && t.accepting[s]) {
        next = s;
      } else if (k == K) {
        next = contained ? 0 : dead;
      } else {
        int e = t.entries[static_cast<size_t>(s) * p.stride + k];
        next = (e == -1) ? (contained ? 0 : dead) : e;
      }
      d.trans[static_cast<size_t>(s) * d.n_classes + k] = static_cast<uint16_t>(next);
    }
  }
  if (contained) {
    // DEAD is unreachable for containedIn; keep it well-formed anyway
    for (int k = 0; k <= K; k++) d.trans[static_cast<size_t>(dead) * d.n_classes + k] = 0;
  }
  return d;
}

}  // namespace ndl
