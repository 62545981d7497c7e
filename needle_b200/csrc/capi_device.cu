// C ABI, device half: pattern upload, batch launch, host<->device staging.  See include/needle_b200.h.
//
// There is deliberately no CPU matcher behind these entry points: without a usable CUDA device they
// return NDL_ECUDA.
#include <cuda_runtime.h>

#include <atomic>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "capi_internal.h"
#include "device_image.h"
#include "host/pattern.h"
#include "kernels/generic.cuh"
#include "kernels/lines8.cuh"
#include "needle_b200.h"

namespace ndl {

static std::atomic<uint64_t> g_launches{0};

#define NDL_CUDA(expr)                                                                             \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess)                                                                         \
      return fail(NDL_ECUDA, std::string(#expr) + " failed: " + cudaGetErrorString(_e));           \
  } while (0)

struct DeviceTableStorage {
  HostDeviceTable host;
  uint16_t* cmap = nullptr;
  uint16_t* trans = nullptr;
  uint8_t* accept = nullptr;
  DevTable view() const {
    DevTable v;
    v.cmap = cmap;
    v.trans = trans;
    v.accept = accept;
    v.n_states = host.n_states;
    v.n_classes = host.n_classes;
    v.root_accepting = host.root_accepting ? 1 : 0;
    return v;
  }
};

// Grow-only device staging for NDL_MEM_HOST calls.
struct Workspace {
  void* data = nullptr;
  size_t data_cap = 0;
  uint64_t* offsets = nullptr;
  int32_t* from = nullptr;
  uint8_t* matched = nullptr;
  int32_t* start = nullptr;
  int32_t* end = nullptr;
  size_t n_cap = 0;
};

}  // namespace ndl

using namespace ndl;

struct ndl_pattern {
  CompiledPattern cp;
  int device = 0;
  int sm_count = 0;
  DeviceTableStorage tables[4];
  Lines8Blob l8[3];  // per mode: shared-memory image of the byte-input kernel
  std::mutex ws_mutex;
  Workspace ws;
};

namespace ndl {

static int upload_table(DeviceTableStorage& t) {
  NDL_CUDA(cudaMalloc(&t.cmap, t.host.cmap.size() * sizeof(uint16_t)));
  NDL_CUDA(cudaMalloc(&t.trans, t.host.trans.size() * sizeof(uint16_t)));
  NDL_CUDA(cudaMalloc(&t.accept, t.host.accept.size()));
  NDL_CUDA(cudaMemcpy(t.cmap, t.host.cmap.data(), t.host.cmap.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
  NDL_CUDA(cudaMemcpy(t.trans, t.host.trans.data(), t.host.trans.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
  NDL_CUDA(cudaMemcpy(t.accept, t.host.accept.data(), t.host.accept.size(), cudaMemcpyHostToDevice));
  return NDL_OK;
}

static void free_pattern(ndl_pattern* p) {
  if (!p) return;
  int prev = 0;
  cudaGetDevice(&prev);
  cudaSetDevice(p->device);
  for (auto& t : p->tables) {
    cudaFree(t.cmap);
    cudaFree(t.trans);
    cudaFree(t.accept);
  }
  for (auto& b : p->l8) cudaFree(b.dev);
  cudaFree(p->ws.data);
  cudaFree(p->ws.offsets);
  cudaFree(p->ws.from);
  cudaFree(p->ws.matched);
  cudaFree(p->ws.start);
  cudaFree(p->ws.end);
  cudaSetDevice(prev);
  delete p;
}

static int ensure_workspace(Workspace& ws, size_t data_bytes, uint64_t n, bool want_from, bool want_pos) {
  if (data_bytes > ws.data_cap) {
    cudaFree(ws.data);
    ws.data = nullptr;
    ws.data_cap = 0;
    size_t cap = data_bytes + data_bytes / 8 + 256;
    NDL_CUDA(cudaMalloc(&ws.data, cap));
    ws.data_cap = cap;
  }
  if (n + 1 > ws.n_cap || (want_from && !ws.from) || (want_pos && !ws.start)) {
    size_t cap = n + n / 8 + 16;
    if (cap < ws.n_cap) cap = ws.n_cap;
    cudaFree(ws.offsets); cudaFree(ws.from); cudaFree(ws.matched); cudaFree(ws.start); cudaFree(ws.end);
    ws.offsets = nullptr; ws.from = nullptr; ws.matched = nullptr; ws.start = nullptr; ws.end = nullptr;
    ws.n_cap = 0;
    NDL_CUDA(cudaMalloc(&ws.offsets, (cap + 1) * sizeof(uint64_t)));
    NDL_CUDA(cudaMalloc(&ws.from, cap * sizeof(int32_t)));
    NDL_CUDA(cudaMalloc(&ws.matched, cap));
    NDL_CUDA(cudaMalloc(&ws.start, cap * sizeof(int32_t)));
    NDL_CUDA(cudaMalloc(&ws.end, cap * sizeof(int32_t)));
    ws.n_cap = cap;
  }
  return NDL_OK;
}

// Launch the kernels for one batch whose buffers are all on the device.
static int launch_batch(ndl_pattern* p, const BatchParams& bp, int char_width, uint64_t total_chars, cudaStream_t stream) {
  if (bp.n == 0) return NDL_OK;
  (void)total_chars;
  if (char_width == 1 && bp.from == nullptr && p->l8[bp.mode].ok && bp.n >= 2 && bp.n < (1ull << 31)) {
    const Lines8Blob& img = p->l8[bp.mode];
    Lines8Params lp;
    lp.g = bp;
    lp.image = img.dev;
    lp.trans_bytes = img.trans_bytes;
    lp.root_entry = img.root_entry;
    uint64_t max_tiles = (bp.n + 1023) / 1024;  // a CTA's 32 warps take 32 lines each per round
    int blocks = static_cast<int>(max_tiles < static_cast<uint64_t>(p->sm_count) ? max_tiles : p->sm_count);
    lines8_kernel<<<blocks, kL8Threads, kL8DynSmem, stream>>>(lp);
    g_launches.fetch_add(1);
    NDL_CUDA(cudaGetLastError());
    return NDL_OK;
  }
  const int threads = 256;
  uint64_t blocks64 = (bp.n + threads - 1) / threads;
  const uint64_t max_blocks = static_cast<uint64_t>(p->sm_count) * 32;
  int blocks = static_cast<int>(blocks64 < max_blocks ? blocks64 : max_blocks);
  if (char_width == 1)
    generic_batch_kernel<uint8_t><<<blocks, threads, 0, stream>>>(bp);
  else
    generic_batch_kernel<uint16_t><<<blocks, threads, 0, stream>>>(bp);
  g_launches.fetch_add(1);
  NDL_CUDA(cudaGetLastError());
  return NDL_OK;
}

}  // namespace ndl

extern "C" {

int ndl_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

uint64_t ndl_kernel_launches(void) { return g_launches.load(); }

int ndl_pattern_device(const ndl_pattern* p) { return p ? p->device : -1; }

int ndl_pattern_create(const uint8_t* blob, size_t blob_len, int device, ndl_pattern** out) {
  if (!out) return fail(NDL_EINVAL, "out must not be NULL");
  *out = nullptr;
  CompiledPattern cp;
  try {
    cp = deserialize_pattern(blob, blob_len);
  } catch (const std::exception& e) {
    return fail(NDL_EBLOB, e.what());
  }
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
    cudaGetLastError();
    return fail(NDL_ECUDA, "no CUDA device available (needle_b200 has no CPU fallback)");
  }
  if (device < 0 || device >= count) return fail(NDL_EINVAL, "device ordinal out of range");
  NDL_CUDA(cudaSetDevice(device));
  ndl_pattern* p = new ndl_pattern();
  p->cp = std::move(cp);
  p->device = device;
  cudaDeviceProp prop;
  cudaError_t e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) {
    delete p;
    return fail(NDL_ECUDA, std::string("cudaGetDeviceProperties failed: ") + cudaGetErrorString(e));
  }
  p->sm_count = prop.multiProcessorCount;
  for (int k = 0; k < 4; k++) {
    p->tables[k].host = build_device_table(p->cp, k);
    int rc = upload_table(p->tables[k]);
    if (rc != NDL_OK) {
      free_pattern(p);
      return rc;
    }
  }
  // shared-memory images of the byte-input kernel, one per mode
  cudaError_t ae = cudaFuncSetAttribute(lines8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kL8DynSmem);
  for (int mode = 0; mode < 3 && ae == cudaSuccess; mode++) {
    std::vector<uint8_t> img;
    Lines8Blob& b = p->l8[mode];
    const bool ok = lines8_layout(p->tables[mode == NDL_MODE_FIND ? kForwards : mode].host, img, b);
    if (!ok) continue;
    if (cudaMalloc(&b.dev, img.size()) != cudaSuccess ||
        cudaMemcpy(b.dev, img.data(), img.size(), cudaMemcpyHostToDevice) != cudaSuccess) {
      cudaGetLastError();
      free_pattern(p);
      return fail(NDL_ECUDA, "uploading the shared-memory table image failed");
    }
    b.ok = true;
  }
  if (ae != cudaSuccess) cudaGetLastError();  // device cannot give the kernel its shared memory: generic path only
  *out = p;
  return NDL_OK;
}

void ndl_pattern_destroy(ndl_pattern* p) { free_pattern(p); }

int ndl_match_batch(ndl_pattern* p, int mode, const void* data, const uint64_t* offsets, uint64_t n, int char_width,
                    const int32_t* from, uint8_t* matched, int32_t* start, int32_t* end, int mem_kind, void* stream_) {
  if (!p) return fail(NDL_EINVAL, "pattern must not be NULL");
  if (mode < 0 || mode > 2) return fail(NDL_EINVAL, "mode must be NDL_MODE_MATCHES, _CONTAINEDIN or _FIND");
  if (char_width != 1 && char_width != 2) return fail(NDL_EINVAL, "char_width must be 1 or 2");
  if (mem_kind != NDL_MEM_HOST && mem_kind != NDL_MEM_DEVICE) return fail(NDL_EINVAL, "mem_kind must be NDL_MEM_HOST or NDL_MEM_DEVICE");
  if (n == 0) return NDL_OK;
  if (!offsets || !matched) return fail(NDL_EINVAL, "offsets and matched must not be NULL");
  if (mode == NDL_MODE_FIND && (!start || !end)) return fail(NDL_EINVAL, "start and end are required for NDL_MODE_FIND");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  NDL_CUDA(cudaSetDevice(p->device));

  BatchParams bp;
  std::memset(&bp, 0, sizeof(bp));
  bp.n = n;
  bp.mode = mode;
  bp.min_length = p->cp.min_length;
  bp.max_length = p->cp.max_length;
  bp.reverse_mode = p->cp.reverse_mode;
  bp.reverse_char = p->cp.reverse_char;
  bp.fwd = p->tables[mode == NDL_MODE_FIND ? kForwards : mode].view();
  bp.bwd = p->tables[kBackwards].view();

  if (mem_kind == NDL_MEM_DEVICE) {
    bp.data = data;
    bp.offsets = offsets;
    bp.from = from;
    bp.matched = matched;
    bp.start = start;
    bp.end = end;
    // total chars are only a sizing hint for the fast path; it reads the real offsets on the device
    return launch_batch(p, bp, char_width, 0, stream);
  }

  // Host buffers: stage in, launch, stage out, all on `stream`, then wait.
  if (offsets[0] > offsets[n]) return fail(NDL_EINVAL, "offsets must be non-decreasing");
  const uint64_t base = offsets[0];
  const uint64_t total_chars = offsets[n] - base;
  const size_t data_bytes = static_cast<size_t>(total_chars) * char_width;
  std::lock_guard<std::mutex> lock(p->ws_mutex);
  Workspace& ws = p->ws;
  int rc = ensure_workspace(ws, data_bytes + 64, n, from != nullptr, mode == NDL_MODE_FIND);
  if (rc != NDL_OK) return rc;
  if (data_bytes)
    NDL_CUDA(cudaMemcpyAsync(ws.data, static_cast<const uint8_t*>(data) + base * char_width, data_bytes, cudaMemcpyHostToDevice, stream));
  NDL_CUDA(cudaMemcpyAsync(ws.offsets, offsets, (n + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, stream));
  if (from) NDL_CUDA(cudaMemcpyAsync(ws.from, from, n * sizeof(int32_t), cudaMemcpyHostToDevice, stream));
  // the staged copy starts at offsets[0]; bias the data pointer instead of rewriting the offsets
  bp.data = static_cast<const uint8_t*>(ws.data) - base * char_width;
  bp.offsets = ws.offsets;
  bp.from = from ? ws.from : nullptr;
  bp.matched = ws.matched;
  bp.start = ws.start;
  bp.end = ws.end;
  rc = launch_batch(p, bp, char_width, total_chars, stream);
  if (rc != NDL_OK) return rc;
  NDL_CUDA(cudaMemcpyAsync(matched, ws.matched, n, cudaMemcpyDeviceToHost, stream));
  if (mode == NDL_MODE_FIND) {
    NDL_CUDA(cudaMemcpyAsync(start, ws.start, n * sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
    NDL_CUDA(cudaMemcpyAsync(end, ws.end, n * sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
  }
  NDL_CUDA(cudaStreamSynchronize(stream));
  return NDL_OK;
}

int ndl_find_long(ndl_pattern* p, const void* data, uint64_t n_chars, int char_width, int64_t from, uint8_t* matched,
                  int64_t* start, int64_t* end, int mem_kind, void* stream) {
  (void)p; (void)data; (void)n_chars; (void)char_width; (void)from; (void)matched; (void)start; (void)end; (void)mem_kind; (void)stream;
  return fail(NDL_EINVAL, "ndl_find_long: chunk-parallel single-haystack path is not built yet");
}

}  // extern "C"
