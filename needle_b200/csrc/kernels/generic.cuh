// Generic batch kernel: one haystack per thread, any char width, any table size, tables read from
// global memory through the read-only path.  This is the shape-agnostic path (UTF-16 input, DFAs too
// large for shared memory, arbitrarily long haystacks); the tuned byte-input kernels live in
// lines8.cuh.  Same device automaton (device_image.h), same results.
#pragma once
#include <cstdint>

namespace ndl {

struct DevTable {
  const uint16_t* cmap;   // 65536 entries: char -> class column
  const uint16_t* trans;  // (n_states + 1) * n_classes
  const uint8_t* accept;  // n_states + 1
  int n_states;           // DEAD == n_states
  int n_classes;
  int root_accepting;
};

struct BatchParams {
  const void* data;
  const uint64_t* offsets;  // n + 1, in chars; NULL: fixed-length lines, haystack i starts at i * line_chars
  uint64_t line_chars;      // (ndl_match_lines)
  const int32_t* from;      // nullable
  uint8_t* matched;
  int32_t* start;           // nullable unless mode == find
  int32_t* end;
  uint64_t n;
  int mode;
  int min_length, max_length;
  int reverse_mode, reverse_char;
  DevTable fwd;  // the table the mode walks forwards: MATCHES / CONTAINEDIN / FORWARDS
  DevTable bwd;  // BACKWARDS (find only)
  // mode kModeFindAll (ndl_find_all_batch on the tile kernels): counts[n]; with match_offsets, match k of haystack i is
  // stored in start / end at match_offsets[i] + k while that is below match_offsets[i + 1]
  uint32_t* counts;
  const uint64_t* match_offsets;
};
constexpr int kModeFindAll = 3;  // internal: iterated find() (`while (m.find())`), forward table = FORWARDS

// offset (in chars) of haystack i: from the offsets array, or computed for fixed-length lines - which saves the
// 8 bytes per line of HBM traffic (and of PCIe traffic on the host path) that reading them would cost
__device__ __forceinline__ uint64_t batch_off(const BatchParams& p, uint64_t i) { return p.offsets ? p.offsets[i] : i * p.line_chars; }

template <typename CharT>
__device__ __forceinline__ int dev_step(const DevTable& t, int state, CharT c) {
  return __ldg(t.trans + state * t.n_classes + __ldg(t.cmap + c));
}

// matches(): DFAClassBuilder.java:854-912.
template <typename CharT>
__device__ __forceinline__ bool dev_matches(const BatchParams& p, const CharT* s, int64_t len) {
  if (p.min_length > 4 && p.min_length > len) return false;  // DFAMethodComponents.java:75-93
  if (p.max_length != -1 && len > p.max_length) return false;
  const int dead = p.fwd.n_states;
  int state = 0;
  for (int64_t i = 0; i < len; i++) {
    state = dev_step(p.fwd, state, s[i]);
    if (state == dead) return false;
  }
  return __ldg(p.fwd.accept + state) != 0;
}

// containedIn(): DFAClassBuilder.java:956-1025 (accepting rows are absorbing in the device table).
template <typename CharT>
__device__ __forceinline__ bool dev_contained_in(const BatchParams& p, const CharT* s, int64_t len) {
  int state = 0;
  for (int64_t i = 0; i < len; i++) {
    if (__ldg(p.fwd.accept + state)) return true;
    state = dev_step(p.fwd, state, s[i]);
  }
  return __ldg(p.fwd.accept + state) != 0;
}

// indexForwards(from, _): DFAClassBuilder.java:335-471.
template <typename CharT>
__device__ __forceinline__ int64_t dev_index_forwards(const BatchParams& p, const CharT* s, int64_t len, int64_t from) {
  const int dead = p.fwd.n_states;
  int state = 0;
  // root accepting: `lastMatch = 0`, then the top-of-loop wasAccepted check of the first iteration sets it to `from`
  int64_t last = p.fwd.root_accepting ? (from < len ? from : 0) : -1;
  for (int64_t i = from; i < len; i++) {
    state = dev_step(p.fwd, state, s[i]);
    if (state == dead) return last;
    if (__ldg(p.fwd.accept + state)) last = i + 1;
  }
  return last;
}

// indexBackwards(index = end - 1, lowerBound = from): DFAClassBuilder.java:529-614.
template <typename CharT>
__device__ __forceinline__ int64_t dev_index_backwards(const BatchParams& p, const CharT* s, int64_t index, int64_t lower,
                                                       int64_t int_max) {
  if (p.reverse_mode == 1) {  // single-char reverse scan
    for (; index >= lower; index--)
      if (static_cast<int>(s[index]) == p.reverse_char) return index;
    return int_max;
  }
  const int dead = p.bwd.n_states;
  int64_t last = p.bwd.root_accepting ? lower : int_max;
  int state = 0;
  for (; index >= lower; index--) {
    state = dev_step(p.bwd, state, s[index]);
    if (state == dead) return last;
    if (__ldg(p.bwd.accept + state)) last = index;
  }
  return last;
}

template <typename CharT>
__global__ void __launch_bounds__(256) generic_batch_kernel(const BatchParams p) {
  const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
  for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < p.n; i += stride) {
    const uint64_t o0 = batch_off(p, i), o1 = batch_off(p, i + 1);
    const CharT* s = static_cast<const CharT*>(p.data) + o0;
    const int64_t len = static_cast<int64_t>(o1 - o0);
    if (p.mode == 0) {
      p.matched[i] = dev_matches(p, s, len);
    } else if (p.mode == 1) {
      p.matched[i] = dev_contained_in(p, s, len);
    } else {
      const int64_t from = p.from ? p.from[i] : 0;
      const int64_t e = dev_index_forwards(p, s, len, from);  // find(from, to): DFAClassBuilder.java:625-659
      int64_t st = -1;
      if (e != -1) st = (p.reverse_mode == 2) ? e - p.min_length : dev_index_backwards(p, s, e - 1, from, 0x7fffffff);
      p.matched[i] = e != -1;
      p.start[i] = static_cast<int32_t>(st);
      p.end[i] = static_cast<int32_t>(e);
    }
  }
}

// All non-overlapping matches of every haystack: the loop `while (m.find()) { m.start(); m.end(); }` of
// DFACompilerTest.java:678-699.  find() resumes at nextStart = end of the previous match
// (DFAClassBuilder.java:634-635) and its reverse pass is bounded below by that index (:640-659).  A match
// that does not move nextStart forward (an empty match) would be reported forever by the
// reference: here it is reported once and ends the loop, so `from` strictly increases and the loop terminates.
// counts[i] = number of matches; when match_offsets != NULL, match k of haystack i is stored at
// match_offsets[i] + k as long as that is below match_offsets[i + 1] (two-pass CSR: count, scan, fill).
// The loop for one haystack, straight from global memory.
template <typename CharT>
__device__ __forceinline__ void dev_find_all_line(const BatchParams& p, uint64_t i) {
  const uint64_t o0 = batch_off(p, i), o1 = batch_off(p, i + 1);
  const CharT* s = static_cast<const CharT*>(p.data) + o0;
  const int64_t len = static_cast<int64_t>(o1 - o0);
  uint64_t out = 0, cap = 0;
  if (p.match_offsets) {
    out = p.match_offsets[i];
    cap = p.match_offsets[i + 1] - out;
  }
  uint32_t count = 0;
  int64_t from = 0;
  for (;;) {
    const int64_t e = dev_index_forwards<CharT>(p, s, len, from);
    if (e == -1) break;
    const int64_t st = (p.reverse_mode == 2) ? e - p.min_length : dev_index_backwards<CharT>(p, s, e - 1, from, 0x7fffffff);
    if (count < cap) {
      p.start[out + count] = static_cast<int32_t>(st);
      p.end[out + count] = static_cast<int32_t>(e);
    }
    count++;
    if (e <= from) break;  // nextStart did not advance: the reference would repeat this match forever
    from = e;
  }
  p.counts[i] = count;
}

struct FindAllParams {
  BatchParams b;                  // data, offsets, n, tables, reverse mode (matched/start/end/from unused)
  uint32_t* counts;
  const uint64_t* match_offsets;  // nullable
  int32_t* starts;
  int32_t* ends;
};

template <typename CharT>
__global__ void __launch_bounds__(256) find_all_kernel(const FindAllParams q) {
  const BatchParams& p = q.b;
  const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
  for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < p.n; i += stride) {
    const uint64_t o0 = batch_off(p, i), o1 = batch_off(p, i + 1);
    const CharT* s = static_cast<const CharT*>(p.data) + o0;
    const int64_t len = static_cast<int64_t>(o1 - o0);
    uint64_t out = 0, cap = 0;
    if (q.match_offsets) {
      out = q.match_offsets[i];
      cap = q.match_offsets[i + 1] - out;
    }
    uint32_t count = 0;
    int64_t from = 0;
    for (;;) {
      const int64_t e = dev_index_forwards<CharT>(p, s, len, from);
      if (e == -1) break;
      const int64_t st = (p.reverse_mode == 2) ? e - p.min_length : dev_index_backwards<CharT>(p, s, e - 1, from, 0x7fffffff);
      if (count < cap) {
        q.starts[out + count] = static_cast<int32_t>(st);
        q.ends[out + count] = static_cast<int32_t>(e);
      }
      count++;
      if (e <= from) break;  // nextStart did not advance: the reference would repeat this match forever
      from = e;
    }
    q.counts[i] = count;
  }
}

}  // namespace ndl
