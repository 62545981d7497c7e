// C ABI, device half: pattern upload, batch launch, host<->device staging.  See include/needle_b200.h.
//
// There is deliberately no CPU matcher behind these entry points: without a usable CUDA device they
// return NDL_ECUDA.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <atomic>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#define NDL_MAIN_TU  // this file compiles the non-template kernels of the shared headers
#include "capi_internal.h"
#include "device_image.h"
#include "host/pattern.h"
#include "host_chunks.h"
#include "host_staging.h"
#include "kernels/generic.cuh"
#include "kernels/lines8.cuh"
#include "kernels/long8.cuh"
#include "kernels/layouts.h"
#include "kernels/instances.h"
#include "needle_b200.h"

namespace ndl {

static std::atomic<uint64_t> g_launches{0};
// Test hook: how the last chunk-parallel ndl_find_long settled - 1 = the first pass verified, k > 1 = after k - 1 further
// passes (one repeat that records exits + refinement passes), -1 = gave up: exact sequential walk.
static std::atomic<int> g_debug_replicas{0};
static std::atomic<int> g_long_passes{0};

#define NDL_CUDA(expr)                                                                             \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess)                                                                         \
      return fail(NDL_ECUDA, std::string(#expr) + " failed: " + cudaGetErrorString(_e));           \
  } while (0)

// Makes `device` current for the scope and restores the caller's device afterwards: a host that drives several
// GPUs from one thread (PyTorch, the multi-device pattern below) must not find its current device switched.
struct DeviceGuard {
  int prev = -1;
  cudaError_t err = cudaSuccess;
  explicit DeviceGuard(int device) {
    if (cudaGetDevice(&prev) != cudaSuccess) {
      cudaGetLastError();
      prev = -1;
    }
    if (prev != device) err = cudaSetDevice(device);
    else prev = -1;  // nothing to restore
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
  DeviceGuard(const DeviceGuard&) = delete;
  DeviceGuard& operator=(const DeviceGuard&) = delete;
};
#define NDL_DEVICE(guard)                                                                                        \
  do {                                                                                                           \
    if ((guard).err != cudaSuccess) return fail(NDL_ECUDA, std::string("cudaSetDevice failed: ") + cudaGetErrorString((guard).err)); \
  } while (0)

struct DeviceTableStorage {
  HostDeviceTable host;
  uint16_t* cmap = nullptr;  // (pointers into the pattern's device arena)
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

// Grow-only staging for NDL_MEM_HOST calls: the device side, and - for callers whose buffers are pageable - a pinned
// bounce ring for the input and pinned mirrors of the result arrays (an asynchronous DMA needs page-locked memory).
struct Workspace {
  void* data = nullptr;
  size_t data_cap = 0;
  uint64_t* offsets = nullptr;
  int32_t* from = nullptr;
  uint8_t* matched = nullptr;
  int32_t* start = nullptr;
  int32_t* end = nullptr;
  size_t n_cap = 0;
  // ndl_find_all_batch on host buffers
  uint32_t* counts = nullptr;
  uint64_t* match_offsets = nullptr;
  size_t counts_cap = 0;
  int32_t* all_starts = nullptr;
  int32_t* all_ends = nullptr;
  size_t all_cap = 0;
  // ndl_find_long: per-tile seam arrays and the scratch record
  uint32_t* seam_guess = nullptr;
  uint32_t* seam_exit = nullptr;
  uint32_t* seam_acc = nullptr;
  size_t seam_cap = 0;
  uint32_t* exit_a = nullptr;  // per-segment exit states of the refinement passes (ping-pong)
  uint32_t* exit_b = nullptr;
  size_t exit_cap = 0;
  void* long_scratch = nullptr;
  // pageable callers
  static constexpr int kRing = 3;
  static constexpr size_t kRingBytes = 32u << 20;
  uint8_t* ring[kRing] = {};
  cudaEvent_t ring_done[kRing] = {};
  int ring_next = 0;
  uint8_t* h_matched = nullptr;  // pinned mirrors of matched / start / end
  int32_t* h_start = nullptr;
  int32_t* h_end = nullptr;
  size_t h_cap = 0;
};

// Everything a pattern puts on a device lives in ONE allocation ("arena"): the four generic tables and the up to twelve
// shared-memory images.  One upload per device - or, for a multi-device pattern, one upload to the first device and one
// NCCL broadcast of the arena to the others.
struct Arena {
  std::vector<uint8_t> host;
  size_t add(const void* src, size_t bytes) {
    const size_t off = (host.size() + 255) & ~static_cast<size_t>(255);
    host.resize(off + bytes);
    if (bytes) std::memcpy(host.data() + off, src, bytes);
    return off;
  }
};
constexpr size_t kNoImage = ~static_cast<size_t>(0);

// Host half of a pattern: built once from the blob, shared by the replicas of a multi-device pattern.
struct HostPattern {
  CompiledPattern cp;
  HostDeviceTable tables[4];
  size_t off_cmap[4], off_trans[4], off_accept[4];
  Lines8Blob l8[3], l16[3], q8[3], q16[3];  // metadata (dev == nullptr); ok = an image exists in the arena
  size_t off_l8[3], off_l16[3], off_q8[3], off_q16[3];
  Arena arena;
};

}  // namespace ndl

using namespace ndl;

struct ndl_pattern {
  CompiledPattern cp;
  int device = 0;  // -1: multi-device pattern, `replicas` holds one pattern per GPU and this object only routes
  int sm_count = 0;
  DeviceTableStorage tables[4];
  Lines8Blob l8[3];   // per mode: shared-memory image of the lines8 kernel for byte haystacks
  Lines8Blob l16[3];  // per mode: same for UTF-16 haystacks (when the class map has a supported char mode)
  Lines8Blob q8[3];   // per mode: SWAR image (linesq_kernel) for byte haystacks, when the class map has a plan
  Lines8Blob q16[3];  // per mode: same for UTF-16 haystacks
  uint8_t* arena = nullptr;
  size_t arena_bytes = 0;
  std::mutex ws_mutex;
  Workspace ws;
  bool long8_ready = false, long16_ready = false;  // long8_kernel's shared-memory attribute is set (byte / UTF-16 kernel)
  bool long_refines = false; // the last chunk-parallel find needed refinement passes: the next one records exits from its first pass
  // host-buffer calls are pipelined in chunks: H2D on s_h2d, kernels on the caller's stream, D2H on s_d2h
  static constexpr int kMaxChunks = 16;
  cudaStream_t s_h2d = nullptr, s_d2h = nullptr, s_own = nullptr;
  cudaEvent_t ev_start = nullptr, ev_h2d[kMaxChunks] = {}, ev_kernel[kMaxChunks] = {};
  std::vector<ndl_pattern*> replicas;
  std::mutex multi_mutex;  // a multi-device pattern runs one call at a time (its replicas' workspaces hold the call's data)
};

namespace ndl {

static void free_pattern(ndl_pattern* p) {
  if (!p) return;
  for (ndl_pattern* r : p->replicas) free_pattern(r);
  if (p->device >= 0) {
    DeviceGuard guard(p->device);
    cudaFree(p->arena);
    if (p->s_h2d) cudaStreamDestroy(p->s_h2d);
    if (p->s_d2h) cudaStreamDestroy(p->s_d2h);
    if (p->s_own) cudaStreamDestroy(p->s_own);
    if (p->ev_start) cudaEventDestroy(p->ev_start);
    for (auto& e : p->ev_h2d) if (e) cudaEventDestroy(e);
    for (auto& e : p->ev_kernel) if (e) cudaEventDestroy(e);
    Workspace& ws = p->ws;
    cudaFree(ws.data);
    cudaFree(ws.offsets);
    cudaFree(ws.from);
    cudaFree(ws.matched);
    cudaFree(ws.start);
    cudaFree(ws.end);
    cudaFree(ws.counts);
    cudaFree(ws.match_offsets);
    cudaFree(ws.all_starts);
    cudaFree(ws.all_ends);
    cudaFree(ws.seam_guess);
    cudaFree(ws.seam_exit);
    cudaFree(ws.seam_acc);
    cudaFree(ws.exit_a);
    cudaFree(ws.exit_b);
    cudaFree(ws.long_scratch);
    for (int k = 0; k < Workspace::kRing; k++) {
      if (ws.ring[k]) cudaFreeHost(ws.ring[k]);
      if (ws.ring_done[k]) cudaEventDestroy(ws.ring_done[k]);
    }
    if (ws.h_matched) cudaFreeHost(ws.h_matched);
    if (ws.h_start) cudaFreeHost(ws.h_start);
    if (ws.h_end) cudaFreeHost(ws.h_end);
  }
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

// Is `p` page-locked (cudaHostAlloc / cudaHostRegister / managed) - i.e. can it be the end of an asynchronous DMA?
static bool host_is_pinned(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}

static int ensure_ring(Workspace& ws) {
  for (int k = 0; k < Workspace::kRing; k++) {
    if (!ws.ring[k]) NDL_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&ws.ring[k]), Workspace::kRingBytes, cudaHostAllocDefault));
    if (!ws.ring_done[k]) NDL_CUDA(cudaEventCreateWithFlags(&ws.ring_done[k], cudaEventDisableTiming));
  }
  return NDL_OK;
}

static int ensure_result_mirrors(Workspace& ws, uint64_t n, bool want_pos) {
  if (n > ws.h_cap || (want_pos && !ws.h_start)) {
    if (ws.h_matched) cudaFreeHost(ws.h_matched);
    if (ws.h_start) cudaFreeHost(ws.h_start);
    if (ws.h_end) cudaFreeHost(ws.h_end);
    ws.h_matched = nullptr; ws.h_start = nullptr; ws.h_end = nullptr;
    ws.h_cap = 0;
    const size_t cap = n + n / 8 + 16;
    NDL_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&ws.h_matched), cap, cudaHostAllocDefault));
    NDL_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&ws.h_start), cap * sizeof(int32_t), cudaHostAllocDefault));
    NDL_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&ws.h_end), cap * sizeof(int32_t), cudaHostAllocDefault));
    ws.h_cap = cap;
  }
  return NDL_OK;
}

// Host -> device copy on `s`.  Pinned source: one asynchronous copy.  Pageable source: pieces of it are copied into
// the pinned ring by the copy threads and sent from there, so the link stays busy while the next piece is being staged.
static int h2d_copy(Workspace& ws, void* dst, const void* src, size_t bytes, bool pinned, cudaStream_t s) {
  if (bytes == 0) return NDL_OK;
  if (pinned) {
    NDL_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, s));
    return NDL_OK;
  }
  for (size_t off = 0; off < bytes; off += Workspace::kRingBytes) {
    const size_t piece = bytes - off < Workspace::kRingBytes ? bytes - off : Workspace::kRingBytes;
    const int slot = ws.ring_next;
    ws.ring_next = (ws.ring_next + 1) % Workspace::kRing;
    NDL_CUDA(cudaEventSynchronize(ws.ring_done[slot]));  // the copy that last used this slot has left it
    CopyPool::instance().copy(ws.ring[slot], static_cast<const uint8_t*>(src) + off, piece);
    NDL_CUDA(cudaMemcpyAsync(static_cast<uint8_t*>(dst) + off, ws.ring[slot], piece, cudaMemcpyHostToDevice, s));
    NDL_CUDA(cudaEventRecord(ws.ring_done[slot], s));
  }
  return NDL_OK;
}

static void fill_lines8_params(Lines8Params& lp, const BatchParams& bp, const Lines8Blob& img) {
  std::memset(&lp, 0, sizeof(lp));
  lp.g = bp;
  lp.image = img.dev;
  lp.trans_bytes = img.trans_bytes;
  lp.root_entry = img.root_entry;
  lp.bwd_root = img.bwd_root;
  lp.bwd_dead = img.bwd_dead;
  lp.fwd_dead = img.fwd_dead;
  lp.ua = img.ua;
  lp.ub = img.ub;
  lp.xa = img.xa;
  lp.xb = img.xb;
  lp.mixed_page = img.mixed_page;
  lp.replicated = img.replicated;
  lp.row_bytes = img.row_bytes;
  lp.char_mode = img.char_mode;
  lp.has_bwd = img.has_bwd ? 1 : 0;
  lp.q = img.q;
  static const bool no_rounds = std::getenv("NDL_NO_ROUNDS") != nullptr;  // (experiments only)
  lp.no_rounds = no_rounds ? 1u : 0u;
  static const char* rr_min = std::getenv("NDL_RR_MIN_MEAN");  // (experiments only)
  lp.rr_min_mean = rr_min ? static_cast<uint32_t>(std::atoi(rr_min)) : kRrMinMeanBytes;
  static const char* rounds_max = std::getenv("NDL_ROUNDS_MAX_CPL");  // (experiments only)
  lp.rounds_max_cpl = rounds_max ? static_cast<uint32_t>(std::atoi(rounds_max)) : kMaxRoundsCpl;
}

// Launch the kernels for one batch whose buffers are all on the device.
static int launch_batch(ndl_pattern* p, const BatchParams& bp, int char_width, uint64_t total_chars, cudaStream_t stream) {
  if (bp.n == 0) return NDL_OK;
  (void)total_chars;
  const Lines8Blob& qimg = char_width == 1 ? p->q8[bp.mode] : p->q16[bp.mode];
  const Lines8Blob& limg = char_width == 1 ? p->l8[bp.mode] : p->l16[bp.mode];
  const bool tiles = (qimg.ok || limg.ok) && bp.n >= 2 && bp.n < (1ull << 31);
  if (tiles) {
    const Lines8Blob& img = qimg.ok ? qimg : limg;
    Lines8Params lp;
    fill_lines8_params(lp, bp, img);
    if (qimg.ok) {
      const uint64_t per_cta = 32ull * 21;  // lines a CTA's warps take per round
      const uint64_t want = (bp.n + per_cta - 1) / per_cta;
      const int blocks = static_cast<int>(want < static_cast<uint64_t>(p->sm_count) ? want : p->sm_count);
      linesq_kernel_for(qimg.char_mode)<<<blocks, kQThreads, kL8DynSmem, stream>>>(lp);
    } else {
      uint64_t max_tiles = (bp.n + 1023) / 1024;  // a CTA's 32 warps take 32 lines each per round
      int blocks = static_cast<int>(max_tiles < static_cast<uint64_t>(p->sm_count) ? max_tiles : p->sm_count);
      lines8_kernel<<<blocks, kL8Threads, kL8DynSmem, stream>>>(lp);
    }
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

// Test hook (not in include/needle_b200.h): see g_long_passes.
int ndl_debug_long_passes(void) { return g_long_passes.load(); }
// Tests on a one-GPU box: ndl_pattern_create(device = -1) builds `n` replicas on GPU 0 (n < 2: off), so the sharding code of the
// multi-device paths runs there too.
void ndl_debug_force_replicas(int n) { g_debug_replicas.store(n); }

int ndl_pattern_device(const ndl_pattern* p) { return p ? p->device : -1; }

// Test hook (not in include/needle_b200.h): which kernel a (mode, char_width) batch without `from` offsets
// takes.  -1: generic_batch_kernel; otherwise lines8 with char_mode | replicated << 8 | has_bwd << 16 | n_cols << 24.
int ndl_debug_fast_path(const ndl_pattern* p, int mode, int char_width) {
  if (!p || mode < 0 || mode > 2) return -1;
  const Lines8Blob& qb = char_width == 1 ? p->q8[mode] : p->q16[mode];
  const Lines8Blob& b = qb.ok ? qb : char_width == 1 ? p->l8[mode] : p->l16[mode];
  if (!b.ok) return -1;
  return b.char_mode | (b.replicated << 8) | ((b.has_bwd ? 1 : 0) << 16) | (b.n_cols << 24);
}

// Test hook (not in include/needle_b200.h; host only, no device needed): builds the SWAR image of
// (mode, char_width) exactly as ndl_pattern_create does and walks `data` through it on the host with the
// kernel's own integer arithmetic (packed compares, IDP.4A, entry decoding) as lane `lane`, forwards or -
// with the BACKWARDS rows - backwards from the end.  Compares the accept flag after every char and the
// final state with a plain walk of the device table.  Returns the number of differences (0 = identical),
// -1 when the class map has no SWAR plan, -2 on bad arguments.  info[0..3] = char mode, copies, codes, image bytes.
int ndl_debug_swar_emulate(const uint8_t* blob, size_t blob_len, int mode, int char_width, int backward, int lane,
                           const uint8_t* data, uint64_t n_chars, int32_t* info) {
  if (!blob || mode < 0 || mode > 2 || (char_width != 1 && char_width != 2) || lane < 0 || lane > 31) return -2;
  CompiledPattern cp;
  try {
    cp = deserialize_pattern(blob, blob_len);
  } catch (const std::exception&) {
    return -2;
  }
  const HostDeviceTable f = build_device_table(cp, mode == NDL_MODE_FIND ? kForwards : mode);
  const HostDeviceTable bt = build_device_table(cp, kBackwards);
  const bool want_bwd = mode == NDL_MODE_FIND && cp.reverse_mode == kReverseTable;
  if (backward && !want_bwd) return -2;
  std::vector<uint8_t> img;
  Lines8Blob b;
  bool ok = want_bwd && linesq_layout(f, &bt, char_width, img, b);
  if (!ok) {
    if (backward) return -1;
    ok = linesq_layout(f, nullptr, char_width, img, b);
  }
  if (!ok) return -1;
  if (info) {
    info[0] = b.char_mode;
    info[1] = b.replicated;
    info[2] = b.n_cols;
    info[3] = static_cast<int32_t>(img.size());
  }
  const int K = cm_k(b.char_mode), P = cm_planes(b.char_mode);
  const bool u16 = cm_u16(b.char_mode), wide = cm_wide(b.char_mode);
  const uint32_t state_mask = u16 ? 0xffffu >> K : 0xffffffffu >> K;
  const uint32_t lane_off = (static_cast<uint32_t>(lane) & b.q.copy_mask) * b.q.copy_bytes;
  const HostDeviceTable& t = backward ? bt : f;
  auto char_at = [&](uint64_t i) -> uint32_t { return char_width == 1 ? data[i] : (data[2 * i] | data[2 * i + 1] << 8); };
  auto slot_at = [&](uint64_t i) -> uint32_t {  // slot value of char i
    return char_width == 1 ? data[i] : wide ? char_at(i) : data[2 * i + 1];
  };
  auto dp4a = [](uint32_t a, uint32_t w, uint32_t c) {
    for (int i = 0; i < 4; i++) c += ((a >> (8 * i)) & 0xff) * ((w >> (8 * i)) & 0xff);
    return c;
  };
  auto lds = [&](uint32_t addr) -> uint32_t {
    uint32_t v = 0;
    if (addr < kQAbsTrans || addr - kQAbsTrans + (u16 ? 2 : 4) > img.size()) return 0xdeadbeefu;
    std::memcpy(&v, img.data() + (addr - kQAbsTrans), u16 ? 2 : 4);
    return u16 ? v << 16 | v : v;  // 16-bit entry: flags seen in the top bits, row address in the low 14
  };
  const uint32_t top = wide ? 0x80008000u : 0x80808080u;
  auto planes_dp = [&](uint32_t w, int widx, uint32_t acc) {  // compare planes of one word -> IDP.4A with weight set widx
    const uint32_t w80 = w | top, nm = ~w & top;
    for (int p = 0; p < P; p++) acc = dp4a(((w80 - b.q.lo[p]) ^ (w80 - b.q.hi[p])) & nm, b.q.w[p][widx], acc);
    return acc;
  };
  int diffs = 0;
  uint32_t e = (backward ? b.bwd_root : b.root_entry) + lane_off;
  int st = 0;
  const uint64_t groups = (n_chars + 3) / 4;
  for (uint64_t gi = 0; gi < groups; gi++) {
    // four slot values in walk order (backwards: from the last char down); past the ends: 0
    uint64_t idx[4];
    uint32_t c[4];
    for (int j = 0; j < 4; j++) {
      idx[j] = backward ? n_chars - 1 - (4 * gi + j) : 4 * gi + j;  // (wraps past 0: then >= n_chars)
      c[j] = idx[j] < n_chars ? slot_at(idx[j]) : 0;
    }
    uint32_t da, db;  // column offsets of the first / second pair (K = 4: da is the whole group's)
    const int w_first = backward ? 2 : 0;
    if (wide) {
      // a word holds two chars; walking backwards the char walked first sits in the high half
      const uint32_t wa = backward ? (c[1] | c[0] << 16) : (c[0] | c[1] << 16);
      const uint32_t wb = backward ? (c[3] | c[2] << 16) : (c[2] | c[3] << 16);
      if (K == 4) {
        da = planes_dp(wb, backward ? 3 : 1, planes_dp(wa, w_first, 0));
        db = 0;
      } else {
        da = planes_dp(wa, w_first, 0);
        db = planes_dp(wb, w_first, 0);
      }
    } else {
      uint32_t w = 0;  // forwards char j sits in byte j; the kernel's reverse step reads byte 3 first
      for (int j = 0; j < 4; j++) w |= c[j] << (8 * (backward ? 3 - j : j));
      da = planes_dp(w, w_first, 0);
      db = K == 4 ? 0 : planes_dp(w, backward ? 3 : 1, 0);
    }
    uint32_t flags4 = 0;
    if (K == 4) {
      e = lds((u16 ? (da * b.q.kmul) >> 7 : da * b.q.kmul) + (e & state_mask));
      flags4 = e >> 28;
    } else {
      e = lds(da * b.q.kmul + (e & state_mask));
      flags4 = (e >> 30) << 2;
      e = lds(db * b.q.kmul + (e & state_mask));
      flags4 |= e >> 30;
    }
    for (int j = 0; j < 4; j++) {
      if (idx[j] >= n_chars) break;
      st = t.trans[static_cast<size_t>(st) * t.n_classes + t.cmap[char_at(idx[j])]];
      const bool got = (flags4 >> (3 - j)) & 1;
      if (got != (t.accept[st] != 0)) diffs++;
    }
    if (n_chars % 4 == 0 || gi + 1 < groups) {  // whole groups only: the state is comparable
      const uint32_t row = static_cast<uint32_t>((backward ? f.n_states + 1 : 0) + st);
      const uint32_t EB = u16 ? 2 : 4, W = 128 / EB / b.replicated;
      const uint32_t want = kQAbsTrans + (row / W) * 128u + (static_cast<uint32_t>(lane) & b.q.copy_mask) * EB * W + (row % W) * EB;
      if ((e & state_mask) != want) diffs++;
    }
  }
  return diffs;
}

// Bench / test hook (not in include/needle_b200.h): the name of the kernel a (mode, char_width) batch without
// `from` offsets is scanned by, for the roofline record.
const char* ndl_debug_kernel_name(const ndl_pattern* p, int mode, int char_width) {
  static thread_local std::string name;
  if (!p || mode < 0 || mode > 2) return "";
  const Lines8Blob& qb = char_width == 1 ? p->q8[mode] : p->q16[mode];
  const Lines8Blob& lb = char_width == 1 ? p->l8[mode] : p->l16[mode];
  if (qb.ok) {
    name = "linesq_kernel<" + std::to_string(cm_k(qb.char_mode)) + " chars/lookup, " + std::to_string(cm_planes(qb.char_mode)) +
           " compare planes, " + std::to_string(qb.replicated) + (cm_u16(qb.char_mode) ? " table copies of 16-bit entries" : " table copies") +
           (cm_wide(qb.char_mode) ? ", UTF-16 on 16-bit lanes>" : cm_hi(qb.char_mode) ? ", UTF-16 high byte>" : ">");
  } else if (lb.ok) {
    static const char* kModes[] = {"pair table", "UTF-16 high byte", "UTF-16 mixed page", "stride-1 table", "pair table, 16-bit entries"};
    name = std::string("lines8_kernel<") + kModes[lb.char_mode <= 4 ? lb.char_mode : 0] + ">";
  } else {
    name = char_width == 1 ? "generic_batch_kernel<uint8_t>" : "generic_batch_kernel<uint16_t>";
  }
  return name.c_str();
}

}  // extern "C"

namespace ndl {

// Blob -> host half of a pattern: the generic tables and every shared-memory image, laid out in one arena.
static void build_host_pattern(HostPattern& hp) {
  for (int k = 0; k < 4; k++) {
    hp.tables[k] = build_device_table(hp.cp, k);
    const HostDeviceTable& t = hp.tables[k];
    hp.off_cmap[k] = hp.arena.add(t.cmap.data(), t.cmap.size() * sizeof(uint16_t));
    hp.off_trans[k] = hp.arena.add(t.trans.data(), t.trans.size() * sizeof(uint16_t));
    hp.off_accept[k] = hp.arena.add(t.accept.data(), t.accept.size());
  }
  for (int cw = 1; cw <= 2; cw++)
    for (int mode = 0; mode < 3; mode++) {
      const HostDeviceTable& fwd_t = hp.tables[mode == NDL_MODE_FIND ? kForwards : mode];
      const bool want_bwd = mode == NDL_MODE_FIND && hp.cp.reverse_mode == kReverseTable;
      const HostDeviceTable* bwd_t = want_bwd ? &hp.tables[kBackwards] : nullptr;
      std::vector<uint8_t> img;
      // --- "L" image (class map in shared memory, lines8_kernel)
      // Preference (measured, exp/large_table.py): bank-replicated pair tables, unreplicated pair tables (two dependent
      // lookups per char cost more than bank conflicts: [Ss]herlock 3.07 against 2.40 TB/s), the stride-1 table in 32
      // copies, then one plain copy of it (large tables: thousands of states still fit in shared memory).  When find()
      // needs the table-driven reverse pass, every layout that also holds the BACKWARDS rows comes first: the reverse pass
      // then runs on the staged tile instead of global memory ((Holmes|Watson|...)+ find: 2.14 against 0.88 TB/s).
      {
        Lines8Blob& b = cw == 1 ? hp.l8[mode] : hp.l16[mode];
        bool ok = false;
        if (want_bwd) {
          ok = lines8_layout(fwd_t, bwd_t, cw, false, img, b);
          if (!ok) ok = lines8_layout(fwd_t, bwd_t, cw, true, img, b);
          if (!ok && cw == 1) ok = lines8_layout(fwd_t, bwd_t, cw, true, img, b, true);
          if (!ok && cw == 1) ok = lines8_layout_s1(fwd_t, bwd_t, 32, img, b);
          if (!ok && cw == 1) ok = lines8_layout_s1(fwd_t, bwd_t, 1, img, b);
        }
        if (!ok) ok = lines8_layout(fwd_t, nullptr, cw, false, img, b);
        if (!ok) ok = lines8_layout(fwd_t, nullptr, cw, true, img, b);
        if (!ok && cw == 1) ok = lines8_layout(fwd_t, nullptr, cw, true, img, b, true);
        if (!ok && cw == 1) ok = lines8_layout_s1(fwd_t, nullptr, 32, img, b);
        if (!ok && cw == 1) ok = lines8_layout_s1(fwd_t, nullptr, 1, img, b);
        b.ok = ok;
        b.dev = nullptr;
        (cw == 1 ? hp.off_l8[mode] : hp.off_l16[mode]) = ok ? hp.arena.add(img.data(), img.size()) : kNoImage;
      }
      // --- "Q" image (packed-compare classifier, linesq_kernel), where the class map has a plan
      {
        Lines8Blob b;
        bool ok = want_bwd && linesq_layout(fwd_t, bwd_t, cw, img, b);
        if (!ok) ok = linesq_layout(fwd_t, nullptr, cw, img, b);
        if (ok && !linesq_kernel_for(b.char_mode)) ok = false;
        b.ok = ok;
        b.dev = nullptr;
        (cw == 1 ? hp.q8[mode] : hp.q16[mode]) = b;
        (cw == 1 ? hp.off_q8[mode] : hp.off_q16[mode]) = ok ? hp.arena.add(img.data(), img.size()) : kNoImage;
      }
    }
}

// Device half on `device`: allocates the arena (its content arrives by upload or broadcast), sets the kernel attributes.
static int instantiate_pattern(const HostPattern& hp, int device, ndl_pattern** out) {
  DeviceGuard guard(device);
  NDL_DEVICE(guard);
  cudaDeviceProp prop;
  NDL_CUDA(cudaGetDeviceProperties(&prop, device));
  ndl_pattern* p = new ndl_pattern();
  p->cp = hp.cp;
  p->device = device;
  p->sm_count = prop.multiProcessorCount;
  p->arena_bytes = hp.arena.host.size();
  if (cudaMalloc(&p->arena, p->arena_bytes) != cudaSuccess) {
    cudaGetLastError();
    delete p;
    return fail(NDL_ECUDA, "cudaMalloc of the pattern's device arena failed");
  }
  for (int k = 0; k < 4; k++) {
    p->tables[k].host = hp.tables[k];
    p->tables[k].cmap = reinterpret_cast<uint16_t*>(p->arena + hp.off_cmap[k]);
    p->tables[k].trans = reinterpret_cast<uint16_t*>(p->arena + hp.off_trans[k]);
    p->tables[k].accept = p->arena + hp.off_accept[k];
  }
  // a device that cannot give the kernels their shared memory runs the generic path only
  const bool l_ok = cudaFuncSetAttribute(lines8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kL8DynSmem) == cudaSuccess;
  if (!l_ok) cudaGetLastError();
  for (int mode = 0; mode < 3; mode++) {
    auto place = [&](Lines8Blob& dst, const Lines8Blob& src, size_t off, bool swar) {
      dst = src;
      dst.dev = nullptr;
      dst.ok = false;
      if (!src.ok || off == kNoImage || !l_ok) return;
      if (swar && cudaFuncSetAttribute(linesq_kernel_for(src.char_mode), cudaFuncAttributeMaxDynamicSharedMemorySize, kL8DynSmem) != cudaSuccess) {
        cudaGetLastError();
        return;
      }
      dst.dev = p->arena + off;
      dst.ok = true;
    };
    place(p->l8[mode], hp.l8[mode], hp.off_l8[mode], false);
    place(p->l16[mode], hp.l16[mode], hp.off_l16[mode], false);
    place(p->q8[mode], hp.q8[mode], hp.off_q8[mode], true);
    place(p->q16[mode], hp.q16[mode], hp.off_q16[mode], true);
  }
  *out = p;
  return NDL_OK;
}

// --- NCCL, bound at run time (libnccl.so.2): only a multi-device pattern needs it, and a host process that already
// carries an NCCL (PyTorch) must not get a second copy linked in.
struct NcclApi {
  typedef struct ncclComm* comm_t;
  int (*CommInitAll)(comm_t*, int, const int*) = nullptr;
  int (*CommDestroy)(comm_t) = nullptr;
  int (*Broadcast)(const void*, void*, size_t, int, int, comm_t, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool ok = false;
  std::string why;
  static const NcclApi& get() {
    static NcclApi api = load();
    return api;
  }
  static NcclApi load() {
    NcclApi a;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) {
      a.why = std::string("cannot load libnccl.so.2: ") + dlerror();
      return a;
    }
    a.CommInitAll = reinterpret_cast<decltype(a.CommInitAll)>(dlsym(h, "ncclCommInitAll"));
    a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(dlsym(h, "ncclCommDestroy"));
    a.Broadcast = reinterpret_cast<decltype(a.Broadcast)>(dlsym(h, "ncclBroadcast"));
    a.GroupStart = reinterpret_cast<decltype(a.GroupStart)>(dlsym(h, "ncclGroupStart"));
    a.GroupEnd = reinterpret_cast<decltype(a.GroupEnd)>(dlsym(h, "ncclGroupEnd"));
    a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(dlsym(h, "ncclGetErrorString"));
    a.ok = a.CommInitAll && a.CommDestroy && a.Broadcast && a.GroupStart && a.GroupEnd && a.GetErrorString;
    if (!a.ok) a.why = "libnccl.so.2 lacks an expected symbol";
    return a;
  }
};

// One NCCL broadcast of the arena from replicas[0] to every other replica (NVLink / NVSwitch), all devices driven by
// this thread inside one group call.
static int broadcast_arena(std::vector<ndl_pattern*>& reps) {
  const NcclApi& nccl = NcclApi::get();
  if (!nccl.ok) return fail(NDL_ENCCL, nccl.why);
  const int n = static_cast<int>(reps.size());
  std::vector<int> devs(n);
  for (int i = 0; i < n; i++) devs[i] = reps[i]->device;
  std::vector<NcclApi::comm_t> comms(n, nullptr);
  int rc = nccl.CommInitAll(comms.data(), n, devs.data());
  if (rc != 0) return fail(NDL_ENCCL, std::string("ncclCommInitAll failed: ") + nccl.GetErrorString(rc));
  int result = NDL_OK;
  std::string msg;
  rc = nccl.GroupStart();
  for (int i = 0; i < n && rc == 0; i++) {
    DeviceGuard guard(devs[i]);
    rc = nccl.Broadcast(reps[0]->arena, reps[i]->arena, reps[0]->arena_bytes, /*ncclUint8*/ 1, 0, comms[i], nullptr);
  }
  const int rc_end = nccl.GroupEnd();
  if (rc == 0) rc = rc_end;
  if (rc != 0) {
    result = NDL_ENCCL;
    msg = std::string("ncclBroadcast of the table arena failed: ") + nccl.GetErrorString(rc);
  }
  for (int i = 0; i < n; i++) {
    DeviceGuard guard(devs[i]);
    if (result == NDL_OK && cudaStreamSynchronize(nullptr) != cudaSuccess) {
      result = NDL_ECUDA;
      msg = std::string("waiting for the broadcast failed: ") + cudaGetErrorString(cudaGetLastError());
    }
  }
  for (int i = 0; i < n; i++) nccl.CommDestroy(comms[i]);
  return result == NDL_OK ? NDL_OK : fail(result, msg);
}

}  // namespace ndl

extern "C" {

int ndl_pattern_create(const uint8_t* blob, size_t blob_len, int device, ndl_pattern** out) {
  if (!out) return fail(NDL_EINVAL, "out must not be NULL");
  *out = nullptr;
  HostPattern hp;
  try {
    hp.cp = deserialize_pattern(blob, blob_len);
  } catch (const std::exception& e) {
    return fail(NDL_EBLOB, e.what());
  }
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
    cudaGetLastError();
    return fail(NDL_ECUDA, "no CUDA device available (needle_b200 has no CPU fallback)");
  }
  if (device < -1 || device >= count) return fail(NDL_EINVAL, "device ordinal out of range (-1 = every visible device)");
  build_host_pattern(hp);
  const int forced = device == -1 ? g_debug_replicas.load() : 0;  // tests: several replicas on GPU 0 (no NCCL involved)
  if (device >= 0 || (count == 1 && forced < 2)) {
    const int dev = device >= 0 ? device : 0;
    ndl_pattern* p = nullptr;
    int rc = instantiate_pattern(hp, dev, &p);
    if (rc != NDL_OK) return rc;
    DeviceGuard guard(dev);
    if (cudaMemcpy(p->arena, hp.arena.host.data(), p->arena_bytes, cudaMemcpyHostToDevice) != cudaSuccess) {
      const std::string why = cudaGetErrorString(cudaGetLastError());
      free_pattern(p);
      return fail(NDL_ECUDA, "uploading the table arena failed: " + why);
    }
    *out = p;
    return NDL_OK;
  }
  // device == -1: one replica per visible GPU.  The arena is uploaded to the first and NCCL-broadcast to the others;
  // batches are then sharded across the replicas by ndl_match_batch / ndl_match_lines (host buffers).
  ndl_pattern* root = new ndl_pattern();
  root->cp = hp.cp;
  root->device = -1;
  for (int k = 0; k < 4; k++) root->tables[k].host = hp.tables[k];
  const int n_replicas = forced >= 2 ? forced : count;
  for (int d = 0; d < n_replicas; d++) {
    ndl_pattern* r = nullptr;
    int rc = instantiate_pattern(hp, forced >= 2 ? 0 : d, &r);
    if (rc != NDL_OK) {
      free_pattern(root);
      return rc;
    }
    root->replicas.push_back(r);
  }
  if (forced >= 2) {
    DeviceGuard guard(0);
    for (ndl_pattern* r : root->replicas)
      if (cudaMemcpy(r->arena, hp.arena.host.data(), hp.arena.host.size(), cudaMemcpyHostToDevice) != cudaSuccess) {
        const std::string why = cudaGetErrorString(cudaGetLastError());
        free_pattern(root);
        return fail(NDL_ECUDA, "uploading the table arena failed: " + why);
      }
    *out = root;
    return NDL_OK;
  }
  {
    DeviceGuard guard(0);
    if (cudaMemcpy(root->replicas[0]->arena, hp.arena.host.data(), hp.arena.host.size(), cudaMemcpyHostToDevice) != cudaSuccess) {
      const std::string why = cudaGetErrorString(cudaGetLastError());
      free_pattern(root);
      return fail(NDL_ECUDA, "uploading the table arena failed: " + why);
    }
  }
  int rc = broadcast_arena(root->replicas);
  if (rc != NDL_OK) {
    free_pattern(root);
    return rc;
  }
  *out = root;
  return NDL_OK;
}

void ndl_pattern_destroy(ndl_pattern* p) { free_pattern(p); }

}  // extern "C"

// The pipelined NDL_MEM_HOST path of ndl_match_batch / ndl_match_lines on one device; ws_mutex is held.
static int match_host(ndl_pattern* p, BatchParams bp, int mode, const void* data, const uint64_t* offsets, uint64_t line_chars, uint64_t n,
                      int char_width, const int32_t* from, uint8_t* matched, int32_t* start, int32_t* end, cudaStream_t stream) {
  if (offsets && offsets[0] > offsets[n]) return fail(NDL_EINVAL, "offsets must be non-decreasing");
  const uint64_t base = offsets ? offsets[0] : 0;
  const uint64_t total_chars = offsets ? offsets[n] - base : n * line_chars;
  const size_t data_bytes = static_cast<size_t>(total_chars) * char_width;
  Workspace& ws = p->ws;
  int rc = ensure_workspace(ws, data_bytes + 64, n, from != nullptr, mode == NDL_MODE_FIND);
  if (rc != NDL_OK) return rc;
  if (!p->s_h2d) {
    NDL_CUDA(cudaStreamCreateWithFlags(&p->s_h2d, cudaStreamNonBlocking));
    NDL_CUDA(cudaStreamCreateWithFlags(&p->s_d2h, cudaStreamNonBlocking));
    NDL_CUDA(cudaEventCreateWithFlags(&p->ev_start, cudaEventDisableTiming));
    for (int k = 0; k < ndl_pattern::kMaxChunks; k++) {
      NDL_CUDA(cudaEventCreateWithFlags(&p->ev_h2d[k], cudaEventDisableTiming));
      NDL_CUDA(cudaEventCreateWithFlags(&p->ev_kernel[k], cudaEventDisableTiming));
    }
  }
  // Pageable buffers (a JVM heap array, malloc, numpy) cannot be the end of an asynchronous DMA: their input goes through
  // the pinned ring, their results through pinned mirrors - the pipeline below stays asynchronous either way.
  const bool in_pinned = data_bytes == 0 || host_is_pinned(data);
  const bool off_pinned = !offsets || host_is_pinned(offsets);
  const bool out_pinned = host_is_pinned(matched) && (mode != NDL_MODE_FIND || (host_is_pinned(start) && host_is_pinned(end)));
  if (!in_pinned || !off_pinned || (from && !host_is_pinned(from))) {
    if ((rc = ensure_ring(ws)) != NDL_OK) return rc;
  }
  uint8_t* out_m = matched;
  int32_t *out_s = start, *out_e = end;
  if (!out_pinned) {
    if ((rc = ensure_result_mirrors(ws, n, mode == NDL_MODE_FIND)) != NDL_OK) return rc;
    out_m = ws.h_matched;
    out_s = ws.h_start;
    out_e = ws.h_end;
  }
  const bool from_pinned = !from || host_is_pinned(from);
  // the staged copy starts at offsets[0]; bias the data pointer instead of rewriting the offsets
  bp.data = static_cast<const uint8_t*>(ws.data) - base * char_width;
  // chunks of about 64 MB of haystack (at most kMaxChunks), cut at line boundaries (host_chunks.h)
  const int n_chunks = host_chunk_count(data_bytes, n, 64u << 20, ndl_pattern::kMaxChunks);
  auto off_at = [&](uint64_t i) { return offsets ? offsets[i] : i * line_chars; };
  auto pipeline = [&]() -> int {
    NDL_CUDA(cudaEventRecord(p->ev_start, stream));
    NDL_CUDA(cudaStreamWaitEvent(p->s_h2d, p->ev_start, 0));
    NDL_CUDA(cudaStreamWaitEvent(p->s_d2h, p->ev_start, 0));
    uint64_t i0 = 0;
    for (int k = 0; k < n_chunks && i0 < n; k++) {
      const uint64_t i1 = host_chunk_end(offsets, n, i0, k, n_chunks);
      const uint64_t cnt = i1 - i0;
      const uint64_t o0 = off_at(i0), o1 = off_at(i1);
      if (o1 < o0 || o1 - base > total_chars) return fail(NDL_EINVAL, "offsets must be non-decreasing");
      const size_t c0 = static_cast<size_t>(o0 - base) * char_width, c1 = static_cast<size_t>(o1 - base) * char_width;
      // data first, so that the link is busy while the host looks at this chunk's offsets
      int r = h2d_copy(ws, static_cast<uint8_t*>(ws.data) + c0, static_cast<const uint8_t*>(data) + base * char_width + c0, c1 - c0, in_pinned,
                       p->s_h2d);
      if (r != NDL_OK) return r;
      // Equally spaced offsets need not cross the link (8 bytes per line: 11 % of the traffic of 64-byte lines): the host
      // checks every offset of the chunk - exact, one vectorised pass hidden behind the copy above - and the kernel computes
      // them instead.  The same pass rejects offsets that decrease.
      OffsetScan sc{true, false, 0};
      if (offsets) {
        sc = scan_offsets(offsets + i0, cnt);
        if (!sc.monotonic) return fail(NDL_EINVAL, "offsets must be non-decreasing");
        if (!sc.uniform && (r = h2d_copy(ws, ws.offsets + i0, offsets + i0, (cnt + 1) * sizeof(uint64_t), off_pinned, p->s_h2d)) != NDL_OK) return r;
      }
      if (from && (r = h2d_copy(ws, ws.from + i0, from + i0, cnt * sizeof(int32_t), from_pinned, p->s_h2d)) != NDL_OK) return r;
      NDL_CUDA(cudaEventRecord(p->ev_h2d[k], p->s_h2d));
      NDL_CUDA(cudaStreamWaitEvent(stream, p->ev_h2d[k], 0));
      BatchParams cb = bp;
      cb.n = cnt;
      if (offsets && !sc.uniform) {
        cb.offsets = ws.offsets + i0;
      } else {  // line 0 of the chunk is line i0 of the batch
        cb.offsets = nullptr;
        cb.line_chars = offsets ? sc.stride : line_chars;
        cb.data = static_cast<const uint8_t*>(ws.data) + c0;
      }
      cb.from = from ? ws.from + i0 : nullptr;
      cb.matched = ws.matched + i0;
      cb.start = ws.start + i0;
      cb.end = ws.end + i0;
      if ((r = launch_batch(p, cb, char_width, 0, stream)) != NDL_OK) return r;
      NDL_CUDA(cudaEventRecord(p->ev_kernel[k], stream));
      NDL_CUDA(cudaStreamWaitEvent(p->s_d2h, p->ev_kernel[k], 0));
      NDL_CUDA(cudaMemcpyAsync(out_m + i0, ws.matched + i0, cnt, cudaMemcpyDeviceToHost, p->s_d2h));
      if (mode == NDL_MODE_FIND) {
        NDL_CUDA(cudaMemcpyAsync(out_s + i0, ws.start + i0, cnt * sizeof(int32_t), cudaMemcpyDeviceToHost, p->s_d2h));
        NDL_CUDA(cudaMemcpyAsync(out_e + i0, ws.end + i0, cnt * sizeof(int32_t), cudaMemcpyDeviceToHost, p->s_d2h));
      }
      i0 = i1;
    }
    return NDL_OK;
  };
  rc = pipeline();
  // whatever happened, nothing of this call may still be in flight when it returns (the caller owns the buffers)
  const cudaError_t e1 = cudaStreamSynchronize(p->s_h2d), e2 = cudaStreamSynchronize(stream), e3 = cudaStreamSynchronize(p->s_d2h);
  if (rc != NDL_OK) {
    cudaGetLastError();
    return rc;
  }
  if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess)
    return fail(NDL_ECUDA, std::string("host-buffer pipeline failed: ") + cudaGetErrorString(e1 != cudaSuccess ? e1 : e2 != cudaSuccess ? e2 : e3));
  if (!out_pinned) {
    CopyPool& pool = CopyPool::instance();
    pool.copy(matched, ws.h_matched, n);
    if (mode == NDL_MODE_FIND) {
      pool.copy(start, ws.h_start, n * sizeof(int32_t));
      pool.copy(end, ws.h_end, n * sizeof(int32_t));
    }
  }
  return NDL_OK;
}

// A multi-device pattern (ndl_pattern_create with device = -1): the batch is cut into one contiguous, byte-balanced range
// of lines per GPU and every replica runs its own pipelined host path, concurrently.
static int match_impl(ndl_pattern* p, int mode, const void* data, const uint64_t* offsets, uint64_t line_chars, uint64_t n, int char_width,
                      const int32_t* from, uint8_t* matched, int32_t* start, int32_t* end, int mem_kind, void* stream_);

static int match_multi(ndl_pattern* p, int mode, const void* data, const uint64_t* offsets, uint64_t line_chars, uint64_t n, int char_width,
                       const int32_t* from, uint8_t* matched, int32_t* start, int32_t* end, int mem_kind, void* stream_) {
  if (mem_kind != NDL_MEM_HOST) return fail(NDL_EINVAL, "a multi-device pattern takes host buffers (device memory belongs to one GPU)");
  if (stream_) return fail(NDL_EINVAL, "a multi-device pattern takes no stream (a stream belongs to one GPU)");
  if (offsets && offsets[0] > offsets[n]) return fail(NDL_EINVAL, "offsets must be non-decreasing");
  std::lock_guard<std::mutex> multi_lock(p->multi_mutex);
  const int g = static_cast<int>(p->replicas.size());
  std::vector<uint64_t> bounds(g + 1, n);
  bounds[0] = 0;
  for (int k = 0; k + 1 < g; k++) {
    const uint64_t e = bounds[k] < n ? host_chunk_end(offsets, n, bounds[k], k, g) : n;
    bounds[k + 1] = e;
  }
  std::vector<int> rcs(g, NDL_OK);
  std::vector<std::string> msgs(g);
  std::vector<std::thread> threads;
  auto work = [&](int k) {
    const uint64_t i0 = bounds[k], cnt = bounds[k + 1] - i0;
    if (cnt == 0) return;
    const uint8_t* d = static_cast<const uint8_t*>(data) + (offsets ? 0 : i0 * line_chars * char_width);
    rcs[k] = match_impl(p->replicas[k], mode, d, offsets ? offsets + i0 : nullptr, line_chars, cnt, char_width, from ? from + i0 : nullptr,
                        matched + i0, start ? start + i0 : nullptr, end ? end + i0 : nullptr, NDL_MEM_HOST, nullptr);
    if (rcs[k] != NDL_OK) msgs[k] = ndl_last_error();
  };
  for (int k = 1; k < g; k++) threads.emplace_back(work, k);
  work(0);
  for (auto& t : threads) t.join();
  for (int k = 0; k < g; k++)
    if (rcs[k] != NDL_OK) return fail(rcs[k], "GPU " + std::to_string(p->replicas[k]->device) + ": " + msgs[k]);
  return NDL_OK;
}

static int find_long_impl(ndl_pattern* p, const void* data, uint64_t n_chars, int char_width, int64_t from, int32_t entry_state,
                          int64_t last_init, uint8_t* matched, int64_t* start, int64_t* end, int32_t* exit_state, int mem_kind,
                          void* stream_, bool host_results);

// ndl_match_batch (offsets != NULL) and ndl_match_lines (offsets == NULL: haystack i = data[i * line_chars ..)).
static int match_impl(ndl_pattern* p, int mode, const void* data, const uint64_t* offsets, uint64_t line_chars, uint64_t n, int char_width,
                      const int32_t* from, uint8_t* matched, int32_t* start, int32_t* end, int mem_kind, void* stream_) {
  if (!p) return fail(NDL_EINVAL, "pattern must not be NULL");
  if (mode < 0 || mode > 2) return fail(NDL_EINVAL, "mode must be NDL_MODE_MATCHES, _CONTAINEDIN or _FIND");
  if (char_width != 1 && char_width != 2) return fail(NDL_EINVAL, "char_width must be 1 or 2");
  if (mem_kind != NDL_MEM_HOST && mem_kind != NDL_MEM_DEVICE) return fail(NDL_EINVAL, "mem_kind must be NDL_MEM_HOST or NDL_MEM_DEVICE");
  if (n == 0) return NDL_OK;
  if (!matched) return fail(NDL_EINVAL, "matched must not be NULL");
  if (mode == NDL_MODE_FIND && (!start || !end)) return fail(NDL_EINVAL, "start and end are required for NDL_MODE_FIND");
  if (!offsets && line_chars >= (1ull << 31)) return fail(NDL_EINVAL, "line_chars must be below 2^31");
  if (p->device < 0) return match_multi(p, mode, data, offsets, line_chars, n, char_width, from, matched, start, end, mem_kind, stream_);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  DeviceGuard guard(p->device);
  NDL_DEVICE(guard);

  BatchParams bp;
  std::memset(&bp, 0, sizeof(bp));
  bp.n = n;
  bp.mode = mode;
  bp.line_chars = line_chars;
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
  // A few long haystacks through the batch call (Matcher.find() on a document, a handful of files): a batch kernel would walk
  // each of them with a single lane.  find() without a `from` is exactly ndl_find_long per haystack, which cuts it into segments
  // for the whole GPU.
  constexpr uint64_t kLongRouteChars = 1u << 16, kLongRouteMaxLines = 256;
  if (mode == NDL_MODE_FIND && !from && n <= kLongRouteMaxLines) {
    bool all_long = true;
    for (uint64_t i = 0; i < n && all_long; i++) {
      const uint64_t o0 = offsets ? offsets[i] : i * line_chars, o1 = offsets ? offsets[i + 1] : (i + 1) * line_chars;
      if (o1 < o0) return fail(NDL_EINVAL, "offsets must be non-decreasing");
      all_long = o1 - o0 >= kLongRouteChars && o1 - o0 < (1ull << 31);
    }
    if (all_long) {
      for (uint64_t i = 0; i < n; i++) {
        const uint64_t o0 = offsets ? offsets[i] : i * line_chars, o1 = offsets ? offsets[i + 1] : (i + 1) * line_chars;
        uint8_t m = 0;
        int64_t st = -1, en = -1;
        const int rc = find_long_impl(p, static_cast<const uint8_t*>(data) + o0 * static_cast<uint64_t>(char_width), o1 - o0, char_width, 0, 0, -1, &m,
                                      &st, &en, nullptr, NDL_MEM_HOST, stream_, false);
        if (rc != NDL_OK) return rc;
        matched[i] = m;
        start[i] = static_cast<int32_t>(st);
        end[i] = static_cast<int32_t>(en);
      }
      return NDL_OK;
    }
  }
  std::lock_guard<std::mutex> lock(p->ws_mutex);
  if (!stream) {  // the replicas of a multi-device pattern run concurrently: each on a stream of its own, not the legacy default stream
    if (!p->s_own) NDL_CUDA(cudaStreamCreateWithFlags(&p->s_own, cudaStreamNonBlocking));
    stream = p->s_own;
  }
  return match_host(p, bp, mode, data, offsets, line_chars, n, char_width, from, matched, start, end, stream);
}

extern "C" {

int ndl_match_batch(ndl_pattern* p, int mode, const void* data, const uint64_t* offsets, uint64_t n, int char_width,
                    const int32_t* from, uint8_t* matched, int32_t* start, int32_t* end, int mem_kind, void* stream) {
  if (n != 0 && !offsets) return fail(NDL_EINVAL, "offsets must not be NULL");
  return match_impl(p, mode, data, offsets, 0, n, char_width, from, matched, start, end, mem_kind, stream);
}

int ndl_match_lines(ndl_pattern* p, int mode, const void* data, uint64_t n, uint64_t line_chars, int char_width, uint8_t* matched,
                    int32_t* start, int32_t* end, int mem_kind, void* stream) {
  return match_impl(p, mode, data, nullptr, line_chars, n, char_width, nullptr, matched, start, end, mem_kind, stream);
}

// Page-locked host memory for the caller's batch buffers (a JNI shim backs a direct ByteBuffer with it): buffers from here
// are DMA'd directly, anything else goes through the library's bounce ring.
void* ndl_host_alloc(size_t bytes) {
  void* p = nullptr;
  if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable) != cudaSuccess) {
    fail(NDL_ENOMEM, std::string("cudaHostAlloc failed: ") + cudaGetErrorString(cudaGetLastError()));
    return nullptr;
  }
  return p;
}

void ndl_host_free(void* p) {
  if (p) cudaFreeHost(p);
}

int ndl_pattern_device_count(const ndl_pattern* p) { return !p ? 0 : p->device < 0 ? static_cast<int>(p->replicas.size()) : 1; }

int ndl_find_all_batch(ndl_pattern* p, const void* data, const uint64_t* offsets, uint64_t n, int char_width, uint32_t* counts,
                       const uint64_t* match_offsets, int32_t* starts, int32_t* ends, int mem_kind, void* stream_) {
  if (!p) return fail(NDL_EINVAL, "pattern must not be NULL");
  if (char_width != 1 && char_width != 2) return fail(NDL_EINVAL, "char_width must be 1 or 2");
  if (mem_kind != NDL_MEM_HOST && mem_kind != NDL_MEM_DEVICE) return fail(NDL_EINVAL, "mem_kind must be NDL_MEM_HOST or NDL_MEM_DEVICE");
  if (n == 0) return NDL_OK;
  if (!offsets || !counts) return fail(NDL_EINVAL, "offsets and counts must not be NULL");
  if (match_offsets && (!starts || !ends)) return fail(NDL_EINVAL, "starts and ends are required with match_offsets");
  if (p->device < 0) return fail(NDL_EINVAL, "ndl_find_all_batch needs a single-device pattern");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  DeviceGuard guard(p->device);
  NDL_DEVICE(guard);
  FindAllParams q;
  std::memset(&q, 0, sizeof(q));
  q.b.n = n;
  q.b.mode = NDL_MODE_FIND;
  q.b.min_length = p->cp.min_length;
  q.b.max_length = p->cp.max_length;
  q.b.reverse_mode = p->cp.reverse_mode;
  q.b.reverse_char = p->cp.reverse_char;
  q.b.fwd = p->tables[kForwards].view();
  q.b.bwd = p->tables[kBackwards].view();
  auto launch = [&]() -> int {
    // the tile kernels (staged lines, shared-memory tables) when the pattern has an image for find(); mode kModeFindAll
    const Lines8Blob& qimg = char_width == 1 ? p->q8[NDL_MODE_FIND] : p->q16[NDL_MODE_FIND];
    const Lines8Blob& limg = char_width == 1 ? p->l8[NDL_MODE_FIND] : p->l16[NDL_MODE_FIND];
    if ((qimg.ok || limg.ok) && n >= 2 && n < (1ull << 31)) {
      BatchParams bp = q.b;
      bp.mode = kModeFindAll;
      bp.counts = q.counts;
      bp.match_offsets = q.match_offsets;
      bp.start = q.starts;
      bp.end = q.ends;
      Lines8Params lp;
      fill_lines8_params(lp, bp, qimg.ok ? qimg : limg);
      const uint64_t want_ctas = (n + 671) / 672;
      const int blocks = static_cast<int>(want_ctas < static_cast<uint64_t>(p->sm_count) ? want_ctas : p->sm_count);
      if (qimg.ok) linesq_kernel_for(qimg.char_mode)<<<blocks, kQThreads, kL8DynSmem, stream>>>(lp);
      else lines8_kernel<<<blocks, kL8Threads, kL8DynSmem, stream>>>(lp);
      g_launches.fetch_add(1);
      NDL_CUDA(cudaGetLastError());
      return NDL_OK;
    }
    const int threads = 256;
    const uint64_t want = (n + threads - 1) / threads, max_blocks = static_cast<uint64_t>(p->sm_count) * 32;
    const int blocks = static_cast<int>(want < max_blocks ? want : max_blocks);
    if (char_width == 1)
      find_all_kernel<uint8_t><<<blocks, threads, 0, stream>>>(q);
    else
      find_all_kernel<uint16_t><<<blocks, threads, 0, stream>>>(q);
    g_launches.fetch_add(1);
    NDL_CUDA(cudaGetLastError());
    return NDL_OK;
  };
  if (mem_kind == NDL_MEM_DEVICE) {
    q.b.data = data;
    q.b.offsets = offsets;
    q.counts = counts;
    q.match_offsets = match_offsets;
    q.starts = starts;
    q.ends = ends;
    return launch();
  }
  // host buffers: stage in, launch, stage out on `stream`, then wait
  if (offsets[0] > offsets[n] || !scan_offsets(offsets, n).monotonic) return fail(NDL_EINVAL, "offsets must be non-decreasing");
  if (match_offsets && (match_offsets[0] > match_offsets[n] || !scan_offsets(match_offsets, n).monotonic))
    return fail(NDL_EINVAL, "match_offsets must be non-decreasing");
  const uint64_t base = offsets[0];
  const size_t data_bytes = static_cast<size_t>(offsets[n] - base) * char_width;
  const uint64_t total = match_offsets ? match_offsets[n] - match_offsets[0] : 0;
  std::lock_guard<std::mutex> lock(p->ws_mutex);
  Workspace& ws = p->ws;
  int rc = ensure_workspace(ws, data_bytes + 64, n, false, true);
  if (rc != NDL_OK) return rc;
  // grow-only staging of the counts / CSR arrays (no allocation on the steady-state path)
  if (n + 1 > ws.counts_cap) {
    cudaFree(ws.counts);
    cudaFree(ws.match_offsets);
    ws.counts = nullptr;
    ws.match_offsets = nullptr;
    ws.counts_cap = 0;
    const size_t cap = n + n / 8 + 16;
    NDL_CUDA(cudaMalloc(&ws.counts, cap * sizeof(uint32_t)));
    NDL_CUDA(cudaMalloc(&ws.match_offsets, (cap + 1) * sizeof(uint64_t)));
    ws.counts_cap = cap;
  }
  if (match_offsets && total + 1 > ws.all_cap) {
    cudaFree(ws.all_starts);
    cudaFree(ws.all_ends);
    ws.all_starts = ws.all_ends = nullptr;
    ws.all_cap = 0;
    const size_t cap = total + total / 8 + 16;
    NDL_CUDA(cudaMalloc(&ws.all_starts, cap * sizeof(int32_t)));
    NDL_CUDA(cudaMalloc(&ws.all_ends, cap * sizeof(int32_t)));
    ws.all_cap = cap;
  }
  uint32_t* d_counts = ws.counts;
  uint64_t* d_moff = match_offsets ? ws.match_offsets : nullptr;
  int32_t *d_starts = match_offsets ? ws.all_starts : nullptr, *d_ends = match_offsets ? ws.all_ends : nullptr;
  if (match_offsets) NDL_CUDA(cudaMemcpyAsync(d_moff, match_offsets, (n + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, stream));
  NDL_CUDA(cudaMemcpyAsync(ws.offsets, offsets, (n + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, stream));
  if (data_bytes)
    NDL_CUDA(cudaMemcpyAsync(ws.data, static_cast<const uint8_t*>(data) + base * char_width, data_bytes, cudaMemcpyHostToDevice, stream));
  q.b.data = static_cast<const uint8_t*>(ws.data) - base * char_width;
  q.b.offsets = ws.offsets;
  q.counts = d_counts;
  q.match_offsets = d_moff;
  // the device arrays start at match 0 of this call
  q.starts = d_starts ? d_starts - match_offsets[0] : nullptr;
  q.ends = d_ends ? d_ends - match_offsets[0] : nullptr;
  if ((rc = launch()) != NDL_OK) return rc;
  NDL_CUDA(cudaMemcpyAsync(counts, d_counts, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
  if (match_offsets && total) {
    NDL_CUDA(cudaMemcpyAsync(starts + match_offsets[0], d_starts, total * sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
    NDL_CUDA(cudaMemcpyAsync(ends + match_offsets[0], d_ends, total * sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
  }
  NDL_CUDA(cudaStreamSynchronize(stream));
  return NDL_OK;
}
}  // extern "C"

// device scratch of the ndl_find_long family (kept with the pattern; calls are serialised by ws_mutex)
struct LongScratch { SeqResult r; unsigned long long first_seg, first_bad; int64_t back; Long8Epilogue epi; int64_t back2[2]; };

static int find_long_impl(ndl_pattern* p, const void* data, uint64_t n_chars, int char_width, int64_t from, int32_t entry_state,
                          int64_t last_init, uint8_t* matched, int64_t* start, int64_t* end, int32_t* exit_state, int mem_kind,
                          void* stream_, bool host_results) {  // host_results: device haystack, results to host pointers
  if (!p) return fail(NDL_EINVAL, "pattern must not be NULL");
  if (char_width != 1 && char_width != 2) return fail(NDL_EINVAL, "char_width must be 1 or 2");
  if (mem_kind == NDL_MEM_DEVICE_DATA) {  // device haystack, results to host pointers
    mem_kind = NDL_MEM_DEVICE;
    host_results = true;
  }
  if (mem_kind != NDL_MEM_HOST && mem_kind != NDL_MEM_DEVICE) return fail(NDL_EINVAL, "mem_kind must be NDL_MEM_HOST, NDL_MEM_DEVICE or NDL_MEM_DEVICE_DATA");
  if (!matched || !end || (!start && !exit_state)) return fail(NDL_EINVAL, "matched, start and end must not be NULL");
  if (from < 0) return fail(NDL_EINVAL, "from must be >= 0");
  if (p->device < 0) return fail(NDL_EINVAL, "ndl_find_long needs a single-device pattern");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  DeviceGuard guard(p->device);
  NDL_DEVICE(guard);
  const int64_t n = static_cast<int64_t>(n_chars);
  const int64_t kIntMax = INT64_MAX;

  std::lock_guard<std::mutex> lock(p->ws_mutex);
  // haystack on the device
  const uint8_t* d_data = static_cast<const uint8_t*>(data);
  if (mem_kind == NDL_MEM_HOST) {
    int rc = ensure_workspace(p->ws, static_cast<size_t>(n_chars) * char_width + 64, 16, false, true);
    if (rc != NDL_OK) return rc;
    if (n_chars) {
      const bool pinned = host_is_pinned(data);
      if (!pinned && (rc = ensure_ring(p->ws)) != NDL_OK) return rc;
      if ((rc = h2d_copy(p->ws, p->ws.data, data, static_cast<size_t>(n_chars) * char_width, pinned, stream)) != NDL_OK) return rc;
    }
    d_data = static_cast<const uint8_t*>(p->ws.data);
  }
  // scratch: SeqResult + 2 atomics + back result (kept with the pattern; calls are serialised by ws_mutex)
  if (!p->ws.long_scratch) NDL_CUDA(cudaMalloc(&p->ws.long_scratch, sizeof(LongScratch)));
  typedef LongScratch Scratch;
  Scratch* d_sc = static_cast<Scratch*>(p->ws.long_scratch);
  Scratch h;
  const DevTable fwd = p->tables[kForwards].view();
  const int dead = fwd.n_states;

  // run one sequential walk and fetch its result
  auto seq = [&](int64_t p0, int64_t p1, int32_t state0, int64_t count_from, int64_t last_init, SeqResult& out) -> int {
    if (char_width == 1)
      seq_walk_kernel<uint8_t><<<1, 32, 0, stream>>>(fwd, d_data, p0, p1, state0, count_from, last_init, &d_sc->r);
    else
      seq_walk_kernel<uint16_t><<<1, 32, 0, stream>>>(fwd, reinterpret_cast<const uint16_t*>(d_data), p0, p1, state0, count_from, last_init, &d_sc->r);
    g_launches.fetch_add(1);
    NDL_CUDA(cudaGetLastError());
    NDL_CUDA(cudaMemcpyAsync(&out, &d_sc->r, sizeof(SeqResult), cudaMemcpyDeviceToHost, stream));
    NDL_CUDA(cudaStreamSynchronize(stream));
    return NDL_OK;
  };

  int64_t last = -1;
  int rc = NDL_OK;
  SeqResult r;
  r.state = entry_state;
  r.last = last_init;
  r.pos = from;
  if (entry_state < 0 || entry_state > dead) return fail(NDL_EINVAL, "entry_state out of range");
  const bool root_acc = p->tables[kForwards].host.root_accepting;
  const Lines8Blob& qimg = char_width == 1 ? p->q8[NDL_MODE_FIND] : p->q16[NDL_MODE_FIND];
  const Lines8Blob& img = qimg.ok && long8_kernel_for(qimg.char_mode) ? qimg : char_width == 1 ? p->l8[NDL_MODE_FIND] : p->l16[NDL_MODE_FIND];
  const int64_t cw = char_width;
  // the chunk-parallel path is for the search phase (no match seen yet) of a pattern with a non-accepting root
  const bool fast = !root_acc && img.ok && long8_kernel_for(img.char_mode) != nullptr && from < n && last_init == -1 && entry_state != dead &&
                    (reinterpret_cast<uintptr_t>(d_data) & static_cast<uintptr_t>(cw - 1)) == 0;

  if (!fast) {
    g_long_passes.store(0);  // (test hook: 0 = the chunk-parallel path was not taken)
    // plain sequential walk (accepting root, no shared-memory image for this char width, or continuing a match that is
    // already under way): DFAClassBuilder.java:335-471
    last = last_init;
    if (root_acc && entry_state == 0 && last_init == -1) last = from < n ? from : 0;
    if (from < n && entry_state != dead) {
      if ((rc = seq(from, n, entry_state, from, last, r)) != NDL_OK) return rc;
      last = r.last;
    }
  } else {
    // canonical state encoding of the image (long8.cuh): state id <-> table entry without flags / copy offset
    const bool swar = cm_is_swar(img.char_mode);
    const bool s1 = img.char_mode == kCmBytes1;
    const uint32_t row_bytes = img.char_mode == kCmBytesH ? img.row_bytes / 2 : img.row_bytes;  // what an entry holds per row
    const uint32_t eb = swar && cm_u16(img.char_mode) ? 2u : 4u;
    const uint32_t w_rows = swar ? 128u / eb / static_cast<uint32_t>(img.replicated) : 1u;
    auto enc = [&](int32_t state) -> uint32_t {
      const uint32_t st = static_cast<uint32_t>(state);
      if (swar) return kQAbsTrans + (st / w_rows) * 128u + (st % w_rows) * eb;
      return s1 ? st : st * row_bytes;
    };
    // head: exact walk up to the first 16-byte boundary
    const uintptr_t a0 = reinterpret_cast<uintptr_t>(d_data) + static_cast<uintptr_t>(from * cw);
    int64_t head_end = from + static_cast<int64_t>(((a0 + 15) & ~static_cast<uintptr_t>(15)) - a0) / cw;
    if (head_end > n) head_end = n;
    bool done = false;
    if (head_end > from) {
      if ((rc = seq(from, head_end, entry_state, from, -1, r)) != NDL_OK) return rc;
      if (r.state == dead) {
        last = r.last;
        done = true;
      } else if (r.last != -1) {  // a match began in the head and may run on: follow it to its end
        SeqResult r2;
        if ((rc = seq(head_end, n, r.state, head_end, r.last, r2)) != NDL_OK) return rc;
        last = r2.last;
        r = r2;
        done = true;
      }
    }
    if (!done) {
      const uint64_t n_segs = static_cast<uint64_t>(n - head_end) * static_cast<uint64_t>(cw) / kLongSeg;  // segments are 256 BYTES
      const uint64_t n_tiles = (n_segs + 31) / 32;
      int32_t state = r.state;
      int64_t pos = head_end;
      if (n_segs > 0) {
        Workspace& ws = p->ws;
        if (n_tiles > ws.seam_cap) {
          cudaFree(ws.seam_guess);
          cudaFree(ws.seam_exit);
          cudaFree(ws.seam_acc);
          ws.seam_guess = ws.seam_exit = ws.seam_acc = nullptr;
          ws.seam_cap = 0;
          const size_t cap = n_tiles + n_tiles / 8 + 16;
          NDL_CUDA(cudaMalloc(&ws.seam_guess, cap * sizeof(uint32_t)));
          NDL_CUDA(cudaMalloc(&ws.seam_exit, cap * sizeof(uint32_t)));
          NDL_CUDA(cudaMalloc(&ws.seam_acc, cap * sizeof(uint32_t)));
          ws.seam_cap = cap;
        }
        Long8Params lp;
        lp.data = d_data + head_end * cw;
        lp.n_segs = n_segs;
        lp.image = img.dev;
        lp.trans_bytes = img.trans_bytes;
        lp.root_entry = img.root_entry;
        lp.row_bytes = row_bytes;
        lp.entry0 = enc(r.state);
        lp.q = img.q;
        lp.ua = img.ua;
        lp.ub = img.ub;
        lp.xa = img.xa;
        lp.xb = img.xb;
        lp.mixed_page = img.mixed_page;
        lp.replicated = img.replicated;
        lp.seam_guess = ws.seam_guess;
        lp.seam_exit = ws.seam_exit;
        lp.seam_acc = ws.seam_acc;
        lp.first_seg = &d_sc->first_seg;
        lp.first_bad = &d_sc->first_bad;
        Long8Kernel kern = long8_kernel_for(img.char_mode);
        bool& ready = char_width == 1 ? p->long8_ready : p->long16_ready;
        if (!ready) {  // once per pattern and char width (calls are serialised by ws_mutex)
          NDL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kL8DynSmem));
          ready = true;
        }
        const uint32_t block_warps = swar ? kQWarps : kL8Warps;
        uint64_t want = (n_tiles + block_warps - 1) / block_warps;
        int blocks = static_cast<int>(want < static_cast<uint64_t>(p->sm_count) ? want : p->sm_count);
        Long8Decode dk;
        dk.kind = swar ? 2 : s1 ? 1 : 0;
        dk.row_bytes = row_bytes ? row_bytes : 1;
        dk.w_rows = w_rows;
        dk.entry_bytes = eb;
        // One pass over the segments + the seam check + the epilogue (one host round trip).  entry_in / exit_out: refinement.
        auto run_pass = [&](const uint32_t* entry_in, uint32_t* exit_out) -> int {
          NDL_CUDA(cudaMemsetAsync(&d_sc->first_seg, 0xff, 2 * sizeof(unsigned long long), stream));
          lp.entry_in = entry_in;
          lp.exit_out = exit_out;
          kern<<<blocks, block_warps * 32, kL8DynSmem, stream>>>(lp);
          g_launches.fetch_add(1);
          NDL_CUDA(cudaGetLastError());
          long8_seam_kernel<<<p->sm_count * 4, 256, 0, stream>>>(ws.seam_guess, ws.seam_exit, n_tiles, &d_sc->first_bad);
          g_launches.fetch_add(1);
          NDL_CUDA(cudaGetLastError());
          if (char_width == 1)
            long8_epilogue_kernel<uint8_t><<<1, 32, 0, stream>>>(fwd, d_data, head_end, n, n_segs, r.state, dk, ws.seam_exit, ws.seam_acc,
                                                                 &d_sc->first_seg, &d_sc->first_bad, &d_sc->epi);
          else
            long8_epilogue_kernel<uint16_t><<<1, 32, 0, stream>>>(fwd, reinterpret_cast<const uint16_t*>(d_data), head_end, n, n_segs, r.state, dk,
                                                                  ws.seam_exit, ws.seam_acc, &d_sc->first_seg, &d_sc->first_bad, &d_sc->epi);
          g_launches.fetch_add(1);
          NDL_CUDA(cudaGetLastError());
          NDL_CUDA(cudaMemcpyAsync(&h.epi, &d_sc->epi, sizeof(Long8Epilogue), cudaMemcpyDeviceToHost, stream));
          NDL_CUDA(cudaStreamSynchronize(stream));
          return NDL_OK;
        };
        auto ensure_exits = [&]() -> int {
          if (n_segs > ws.exit_cap) {
            cudaFree(ws.exit_a);
            cudaFree(ws.exit_b);
            ws.exit_a = ws.exit_b = nullptr;
            ws.exit_cap = 0;
            const size_t cap = n_segs + n_segs / 8 + 1024;
            NDL_CUDA(cudaMalloc(&ws.exit_a, cap * sizeof(uint32_t)));
            NDL_CUDA(cudaMalloc(&ws.exit_b, cap * sizeof(uint32_t)));
            ws.exit_cap = cap;
          }
          return NDL_OK;
        };
        int passes = 1;
        const bool recorded = p->long_refines;  // this pattern needed refinement last time: record the exits right away
        if (recorded && (rc = ensure_exits()) != NDL_OK) return rc;
        if ((rc = run_pass(nullptr, recorded ? ws.exit_a : nullptr)) != NDL_OK) return rc;
        p->long_refines = h.epi.status == 1;
        if (h.epi.status == 1) {
          // A guess was wrong before any match: the pattern remembers further back than the 16-byte warm-up (`q[a-z ]*7`).
          // Refine: repeat the pass with every segment entering in the exit state its predecessor had in the previous pass -
          // exact as soon as the automaton's memory fits the segments covered so far - a bounded number of times.
          if ((rc = ensure_exits()) != NDL_OK) return rc;
          uint32_t *prev = ws.exit_a, *next = ws.exit_b;
          if (!recorded) {
            if ((rc = run_pass(nullptr, prev)) != NDL_OK) return rc;  // the same guesses, exits recorded
            passes++;
          }
          constexpr int kMaxRefinePasses = 8;
          for (int pass = 0; pass < kMaxRefinePasses && h.epi.status == 1; pass++) {
            passes++;
            if ((rc = run_pass(prev, next)) != NDL_OK) return rc;
            uint32_t* tmp = prev;
            prev = next;
            next = tmp;
          }
        }
        g_long_passes.store(h.epi.status == 1 ? -1 : passes);
        if (h.epi.status == 1) {
          // still unverified (an automaton that remembers arbitrarily far back, `a.*c`): the exact sequential walk from
          // the end of the head
          if ((rc = seq(head_end, n, r.state, head_end, -1, r)) != NDL_OK) return rc;
          last = r.last;
        } else {
          r = h.epi.r;
          last = r.last;
        }
        done = true;
      }
      if (!done) {
        if (pos < n) {
          if ((rc = seq(pos, n, state, pos, -1, r)) != NDL_OK) return rc;
          last = r.last;
        } else {
          r.state = state;
          r.pos = pos;
          last = -1;
        }
      }
    }
  }

  // find(from, to) glue (DFAClassBuilder.java:625-659) with 64-bit indices
  int64_t st = -1;
  if (last != -1) {
    if (p->cp.reverse_mode == kReverseFixedLength) {
      st = last - p->cp.min_length;
    } else if (start) {
      BatchParams bp;
      std::memset(&bp, 0, sizeof(bp));
      bp.reverse_mode = p->cp.reverse_mode;
      bp.reverse_char = p->cp.reverse_char;
      bp.bwd = p->tables[kBackwards].view();
      if (char_width == 1)
        seq_back_kernel<uint8_t><<<1, 32, 0, stream>>>(bp, d_data, last - 1, from, kIntMax, &d_sc->back);
      else
        seq_back_kernel<uint16_t><<<1, 32, 0, stream>>>(bp, reinterpret_cast<const uint16_t*>(d_data), last - 1, from, kIntMax, &d_sc->back);
      g_launches.fetch_add(1);
      NDL_CUDA(cudaGetLastError());
      NDL_CUDA(cudaMemcpyAsync(&st, &d_sc->back, sizeof(int64_t), cudaMemcpyDeviceToHost, stream));
      NDL_CUDA(cudaStreamSynchronize(stream));
    }
  }
  const uint8_t m = last != -1;
  if (exit_state) *exit_state = r.state;  // state after the last char that was read (DEAD when the walk died)
  if (mem_kind == NDL_MEM_HOST || host_results) {
    *matched = m;
    if (start) *start = st;
    *end = last;
  } else {
    NDL_CUDA(cudaMemcpyAsync(matched, &m, 1, cudaMemcpyHostToDevice, stream));
    if (start) NDL_CUDA(cudaMemcpyAsync(start, &st, sizeof(int64_t), cudaMemcpyHostToDevice, stream));
    NDL_CUDA(cudaMemcpyAsync(end, &last, sizeof(int64_t), cudaMemcpyHostToDevice, stream));
    NDL_CUDA(cudaStreamSynchronize(stream));
  }
  return NDL_OK;
}

// ndl_find_long on a multi-device pattern (host haystack): the protocol of needle_b200/sharding.py inside the library.  The haystack
// is cut into one contiguous chunk per GPU; every replica thread copies its chunk over its own link and scans it from a GUESSED entry
// state (the walk, from the root, of the 16 chars before the chunk - done on the host, the haystack is here); the records (entry,
// end, exit) are then resolved in order from the root: a record whose entry is not the exit of the verified record before it is
// scanned again from the now known state (the chunk is still on its GPU).  The end of the match is the last accepting index on the
// verified path (indexForwards, DFAClassBuilder.java:335-471); the reverse pass (indexBackwards, :529-614) starts on the replica that
// holds end - 1 and is handed down for as long as the BACKWARDS automaton is alive at a chunk boundary.
static int find_long_multi(ndl_pattern* root, const void* data, uint64_t n_chars, int char_width, int64_t from, uint8_t* matched,
                           int64_t* start, int64_t* end, int mem_kind, void* stream_) {
  if (char_width != 1 && char_width != 2) return fail(NDL_EINVAL, "char_width must be 1 or 2");
  if (mem_kind != NDL_MEM_HOST) return fail(NDL_EINVAL, "a multi-device pattern takes host buffers (device memory belongs to one GPU)");
  if (stream_) return fail(NDL_EINVAL, "a multi-device pattern takes no stream (a stream belongs to one GPU)");
  if (!matched || !start || !end) return fail(NDL_EINVAL, "matched, start and end must not be NULL");
  if (from < 0) return fail(NDL_EINVAL, "from must be >= 0");
  std::lock_guard<std::mutex> multi_lock(root->multi_mutex);
  const int64_t n = static_cast<int64_t>(n_chars);
  const int g = static_cast<int>(root->replicas.size());
  constexpr int64_t kMinChunk = 1 << 20;  // below a MiB per GPU the split costs more than it saves
  if (g < 2 || n - from < g * kMinChunk)
    return find_long_impl(root->replicas[0], data, n_chars, char_width, from, 0, -1, matched, start, end, nullptr, NDL_MEM_HOST, nullptr, false);

  const HostDeviceTable& fwd = root->tables[kForwards].host;
  const int32_t fwd_dead = fwd.n_states, bwd_dead = root->tables[kBackwards].host.n_states;
  const uint8_t* const bytes = static_cast<const uint8_t*>(data);
  auto char_at = [&](int64_t i) -> uint32_t {
    return char_width == 1 ? bytes[i] : reinterpret_cast<const uint16_t*>(bytes)[i];
  };
  std::vector<int64_t> b(g + 1);
  for (int k = 0; k <= g; k++) b[k] = k == g ? n : (from + (n - from) / g * k) & ~static_cast<int64_t>(255) ;
  b[0] = from;
  for (int k = 1; k <= g; k++) if (b[k] < b[k - 1]) b[k] = b[k - 1];

  struct Rec {
    int32_t entry = 0, exit = 0;
    int64_t end_local = -1;
    int rc = NDL_OK;
    std::string msg;
  };
  std::vector<Rec> rec(g);
  const bool pinned = host_is_pinned(data);
  auto scan = [&](int k, int32_t entry, bool stage) {
    ndl_pattern* rep = root->replicas[k];
    Rec& r = rec[k];
    r.entry = entry;
    const int64_t len = b[k + 1] - b[k];
    if (len == 0 && k > 0) {  // (replica 0 always scans: an accepting root matches the empty haystack, :356)
      r.end_local = -1;
      r.exit = entry;
      return;
    }
    auto body = [&]() -> int {
      DeviceGuard guard(rep->device);
      NDL_DEVICE(guard);
      if (!rep->s_own) NDL_CUDA(cudaStreamCreateWithFlags(&rep->s_own, cudaStreamNonBlocking));
      if (stage) {
        std::lock_guard<std::mutex> lock(rep->ws_mutex);
        int rc = ensure_workspace(rep->ws, static_cast<size_t>(len) * char_width + 64, 16, false, true);
        if (rc != NDL_OK) return rc;
        if (!pinned && (rc = ensure_ring(rep->ws)) != NDL_OK) return rc;
        rc = h2d_copy(rep->ws, rep->ws.data, bytes + static_cast<size_t>(b[k]) * char_width, static_cast<size_t>(len) * char_width, pinned,
                      rep->s_own);
        if (rc != NDL_OK) return rc;
      }
      uint8_t m = 0;
      return find_long_impl(rep, rep->ws.data, static_cast<uint64_t>(len), char_width, 0, entry, -1, &m, nullptr, &r.end_local, &r.exit,
                            NDL_MEM_DEVICE, rep->s_own, true);
    };
    r.rc = body();
    if (r.rc != NDL_OK) r.msg = ndl_last_error();
  };
  // round 0: every replica at once, entry states guessed from the 16-char halo
  {
    std::vector<std::thread> threads;
    auto guess = [&](int k) -> int32_t {
      int32_t s = 0;
      for (int64_t i = std::max(from, b[k] - 16); i < b[k]; i++) s = fwd.trans[static_cast<size_t>(s) * fwd.n_classes + fwd.cmap[char_at(i)]];
      return s;
    };
    for (int k = 1; k < g; k++) threads.emplace_back([&, k] { scan(k, guess(k), true); });
    scan(0, 0, true);
    for (auto& t : threads) t.join();
  }
  for (int k = 0; k < g; k++)
    if (rec[k].rc != NDL_OK) return fail(rec[k].rc, "GPU " + std::to_string(root->replicas[k]->device) + ": " + rec[k].msg);
  // resolve in order; a wrong guess is scanned again from the true state
  int32_t state = 0;
  int64_t last = -1;
  for (int k = 0; k < g && state != fwd_dead; k++) {
    if (rec[k].entry != state) {
      scan(k, state, false);
      if (rec[k].rc != NDL_OK) return fail(rec[k].rc, "GPU " + std::to_string(root->replicas[k]->device) + ": " + rec[k].msg);
    }
    if (rec[k].end_local != -1) last = b[k] + rec[k].end_local;
    state = rec[k].exit;
  }
  *matched = last != -1;
  *end = last;
  *start = -1;
  if (last == -1) return NDL_OK;
  if (root->cp.reverse_mode == kReverseFixedLength) {
    *start = last - root->cp.min_length;
    return NDL_OK;
  }
  // reverse pass, handed down replica by replica
  const int64_t kNoStart = INT64_MAX;
  int32_t bstate = 0;
  int64_t st = (root->tables[kBackwards].host.root_accepting && root->cp.reverse_mode == 0) ? from : kNoStart;
  int64_t index = last - 1;
  for (int k = g - 1; k >= 0 && index >= from; k--) {
    const int64_t len = b[k + 1] - b[k];
    if (len == 0 || index < b[k] || index >= b[k + 1]) continue;
    ndl_pattern* rep = root->replicas[k];
    int64_t s_local = kNoStart;
    const int rc = ndl_find_long_back(rep, rep->ws.data, static_cast<uint64_t>(len), char_width, index - b[k], 0, bstate,
                                      st == kNoStart ? kNoStart : st - b[k], &s_local, &bstate, NDL_MEM_DEVICE, rep->s_own);
    if (rc != NDL_OK) return rc;
    st = s_local == kNoStart ? kNoStart : s_local + b[k];
    index = b[k] - 1;
    if (bstate == bwd_dead) break;
  }
  *start = st;
  return NDL_OK;
}

extern "C" {

int ndl_find_long(ndl_pattern* p, const void* data, uint64_t n_chars, int char_width, int64_t from, uint8_t* matched,
                  int64_t* start, int64_t* end, int mem_kind, void* stream) {
  if (p && p->device < 0) return find_long_multi(p, data, n_chars, char_width, from, matched, start, end, mem_kind, stream);
  return find_long_impl(p, data, n_chars, char_width, from, 0, -1, matched, start, end, nullptr, mem_kind, stream, false);
}

int ndl_find_long_from(ndl_pattern* p, const void* data, uint64_t n_chars, int char_width, int64_t from, int32_t entry_state,
                       int64_t last_init, uint8_t* matched, int64_t* start, int64_t* end, int32_t* exit_state, int mem_kind,
                       void* stream) {
  return find_long_impl(p, data, n_chars, char_width, from, entry_state, last_init, matched, start, end, exit_state, mem_kind, stream, false);
}

int32_t ndl_forwards_walk_host(const ndl_pattern* p, const void* data, uint64_t n_chars, int char_width, int32_t entry_state) {
  if (!p || (char_width != 1 && char_width != 2) || (!data && n_chars)) return -1;
  const HostDeviceTable& t = p->tables[kForwards].host;
  if (entry_state < 0 || entry_state > t.n_states) return -1;
  int32_t s = entry_state;
  for (uint64_t i = 0; i < n_chars; i++) {
    const uint32_t c = char_width == 1 ? static_cast<const uint8_t*>(data)[i] : static_cast<const uint16_t*>(data)[i];
    s = t.trans[static_cast<size_t>(s) * t.n_classes + t.cmap[c]];
  }
  return s;
}

int ndl_forwards_state_count(const ndl_pattern* p) { return p ? p->tables[kForwards].host.n_states : -1; }
int ndl_backwards_state_count(const ndl_pattern* p) { return p ? p->tables[kBackwards].host.n_states : -1; }
int ndl_backwards_root_accepting(const ndl_pattern* p) { return p && p->tables[kBackwards].host.root_accepting ? 1 : 0; }
int ndl_reverse_mode(const ndl_pattern* p) { return p ? p->cp.reverse_mode : -1; }
int ndl_min_length(const ndl_pattern* p) { return p ? p->cp.min_length : -1; }

int ndl_find_long_back(ndl_pattern* p, const void* data, uint64_t n_chars, int char_width, int64_t index, int64_t lower,
                       int32_t entry_state, int64_t last_init, int64_t* start, int32_t* exit_state, int mem_kind, void* stream_) {
  if (!p) return fail(NDL_EINVAL, "pattern must not be NULL");
  if (char_width != 1 && char_width != 2) return fail(NDL_EINVAL, "char_width must be 1 or 2");
  if (mem_kind != NDL_MEM_HOST && mem_kind != NDL_MEM_DEVICE) return fail(NDL_EINVAL, "mem_kind must be NDL_MEM_HOST or NDL_MEM_DEVICE");
  if (!start || !exit_state) return fail(NDL_EINVAL, "start and exit_state must not be NULL");
  if (p->cp.reverse_mode == kReverseFixedLength) return fail(NDL_EINVAL, "fixed-length pattern: start = end - min_length, no reverse pass");
  const int dead = p->tables[kBackwards].host.n_states;
  if (entry_state < 0 || entry_state > dead) return fail(NDL_EINVAL, "entry_state out of range");
  if (lower < 0 || index >= static_cast<int64_t>(n_chars)) return fail(NDL_EINVAL, "index / lower out of range");
  if (p->device < 0) return fail(NDL_EINVAL, "ndl_find_long_back needs a single-device pattern");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  DeviceGuard guard(p->device);
  NDL_DEVICE(guard);
  std::lock_guard<std::mutex> lock(p->ws_mutex);
  const uint8_t* d_data = static_cast<const uint8_t*>(data);
  if (mem_kind == NDL_MEM_HOST && index >= lower) {
    // only s[lower, index] is read: stage that window
    const size_t w0 = static_cast<size_t>(lower) * char_width, w1 = static_cast<size_t>(index + 1) * char_width;
    int rc = ensure_workspace(p->ws, w1 - w0 + 64, 16, false, true);
    if (rc != NDL_OK) return rc;
    NDL_CUDA(cudaMemcpyAsync(p->ws.data, static_cast<const uint8_t*>(data) + w0, w1 - w0, cudaMemcpyHostToDevice, stream));
    d_data = static_cast<const uint8_t*>(p->ws.data) - w0;
  }
  if (!p->ws.long_scratch) NDL_CUDA(cudaMalloc(&p->ws.long_scratch, sizeof(LongScratch)));
  int64_t* d_out = static_cast<LongScratch*>(p->ws.long_scratch)->back2;
  BatchParams bp;
  std::memset(&bp, 0, sizeof(bp));
  bp.reverse_mode = p->cp.reverse_mode;
  bp.reverse_char = p->cp.reverse_char;
  bp.bwd = p->tables[kBackwards].view();
  if (char_width == 1)
    seq_back_from_kernel<uint8_t><<<1, 32, 0, stream>>>(bp, d_data, index, lower, entry_state, last_init, d_out);
  else
    seq_back_from_kernel<uint16_t><<<1, 32, 0, stream>>>(bp, reinterpret_cast<const uint16_t*>(d_data), index, lower, entry_state, last_init, d_out);
  g_launches.fetch_add(1);
  NDL_CUDA(cudaGetLastError());
  int64_t h[2];
  NDL_CUDA(cudaMemcpyAsync(h, d_out, sizeof(h), cudaMemcpyDeviceToHost, stream));
  NDL_CUDA(cudaStreamSynchronize(stream));
  *start = h[0];
  *exit_state = static_cast<int32_t>(h[1]);
  return NDL_OK;
}

}  // extern "C"
