// Host half of the NDL_MEM_HOST path of ndl_match_batch: how a batch of haystacks in host memory is cut into
// pipeline chunks (H2D / kernel / D2H overlap, capi_device.cu) and how a chunk's offsets are checked.
// No CUDA in here: tests/test_host_chunks.py drives it through ndl_debug_plan_chunks without a GPU.
#pragma once
#include <cstddef>
#include <cstdint>

namespace ndl {

// Number of pipeline chunks for `data_bytes` of haystack: about `chunk_bytes` each, at most `max_chunks`, at most one per line.
inline int host_chunk_count(size_t data_bytes, uint64_t n, size_t chunk_bytes, int max_chunks) {
  if (n == 0) return 0;
  int k = static_cast<int>(data_bytes / (chunk_bytes ? chunk_bytes : 1)) + 1;
  if (k > max_chunks) k = max_chunks;
  if (static_cast<uint64_t>(k) > n) k = static_cast<int>(n);
  return k < 1 ? 1 : k;
}

// End (exclusive line index) of chunk k of n_chunks that starts at line i0 < n.  Chunks are cut at line boundaries
// near equal shares of the chars; every chunk but possibly the last is non-empty, the result never exceeds n, and
// the caller stops as soon as it returns n (so a dominant last line cannot push later chunks past the arrays).
// `offsets` NULL: fixed-length lines.
inline uint64_t host_chunk_end(const uint64_t* offsets, uint64_t n, uint64_t i0, int k, int n_chunks) {
  if (k + 1 >= n_chunks || i0 + 1 >= n) return n;
  if (!offsets) {
    uint64_t i1 = n / static_cast<uint64_t>(n_chunks) * static_cast<uint64_t>(k + 1) +
                  n % static_cast<uint64_t>(n_chunks) * static_cast<uint64_t>(k + 1) / static_cast<uint64_t>(n_chunks);
    if (i1 <= i0) i1 = i0 + 1;
    return i1 > n ? n : i1;
  }
  const uint64_t base = offsets[0], total = offsets[n] - base;
  // target = base + total * (k + 1) / n_chunks without overflowing 64 bits
  const uint64_t kk = static_cast<uint64_t>(k + 1), nc = static_cast<uint64_t>(n_chunks);
  const uint64_t target = base + total / nc * kk + total % nc * kk / nc;
  uint64_t a = i0 + 1, b = n;  // first line index in [i0 + 1, n] whose offset reaches the target
  while (a < b) {
    const uint64_t m = a + (b - a) / 2;
    if (offsets[m] < target) a = m + 1; else b = m;
  }
  return a;
}

// One pass over the cnt + 1 offsets of a chunk (o[0 .. cnt]): are they non-decreasing, and equally spaced
// (then the kernel computes them and they need not cross the link)?  Plain differences, so the loop vectorises.
struct OffsetScan {
  bool monotonic, uniform;
  uint64_t stride;
};
inline OffsetScan scan_offsets(const uint64_t* o, uint64_t cnt) {
  OffsetScan r{true, false, 0};
  if (cnt == 0) return r;
  const uint64_t stride = o[1] - o[0];
  uint64_t differs = 0, negative = 0;
  for (uint64_t j = 1; j <= cnt; j++) {
    const uint64_t d = o[j] - o[j - 1];
    differs |= d ^ stride;
    negative |= d;
  }
  r.monotonic = (negative >> 63) == 0;  // a line of 2^63 chars does not exist; a wrapped difference does
  r.uniform = differs == 0 && stride < (1ull << 31);
  r.stride = stride;
  return r;
}

}  // namespace ndl
