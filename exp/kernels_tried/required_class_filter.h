// Host half of the required-class filter used by chain.cuh (see README.md here).  Was part of kernels/layouts.h.
#pragma once
// Required-class filter of the chain walk (chain.cuh).  A class of chars is REQUIRED when every path of the automaton from
// the root to an accepting state takes a transition on it - no haystack without such a char can match (the `@` of the
// e-mail regex, the `-` of the SSN regex; the reference finds such factors in the AST, Factorization.java, and seeks them
// with indexOf, DFAClassBuilder.java:394-399).  Among the required classes whose bytes form one range below 0x80 the one
// with the fewest bytes is chosen; lo / hi receive the range [lo, hi) replicated into the four bytes of a word.
inline bool required_class_filter(const HostDeviceTable& f, uint32_t& lo, uint32_t& hi) {
  if (f.root_accepting || f.n_states > 4096) return false;
  const int rows = f.n_states + 1, C = f.n_classes;
  int best = -1, best_lo = 0, best_hi = 0;
  std::vector<int> stack;
  std::vector<char> seen(rows);
  for (int k = 0; k < C; k++) {
    int first = -1, last = -1, count = 0;
    for (int v = 0; v < 256; v++)
      if (f.cmap[v] == k) {
        if (first < 0) first = v;
        last = v;
        count++;
      }
    if (count == 0 || last >= 128 || last - first + 1 != count) continue;
    if (best >= 0 && count >= best_hi - best_lo) continue;
    // can an accepting state be reached without class k?
    std::fill(seen.begin(), seen.end(), 0);
    stack.assign(1, 0);
    seen[0] = 1;
    bool reachable = false;
    while (!stack.empty() && !reachable) {
      const int s = stack.back();
      stack.pop_back();
      if (f.accept[s]) {
        reachable = true;
        break;
      }
      for (int c = 0; c < C; c++) {
        if (c == k) continue;
        const int t = f.trans[static_cast<size_t>(s) * C + c];
        if (!seen[t]) {
          seen[t] = 1;
          stack.push_back(t);
        }
      }
    }
    if (!reachable) {
      best = k;
      best_lo = first;
      best_hi = last + 1;
    }
  }
  if (best < 0) return false;
  lo = static_cast<uint32_t>(best_lo) * 0x01010101u;
  hi = static_cast<uint32_t>(best_hi) * 0x01010101u;
  return true;
}

