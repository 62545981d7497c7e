// Host-side helpers of the NDL_MEM_HOST paths (no CUDA in here): a small persistent pool of copy threads.
//
// Callers whose buffers are pageable (a Java heap array behind GetPrimitiveArrayCritical, a numpy array, malloc)
// cannot be the source of an asynchronous DMA; the library bounces such input through its own pinned ring
// (capi_device.cu) and these threads do the pageable -> pinned copies, several in parallel because one core
// copies at well under the PCIe Gen5 rate.
#pragma once
#include <algorithm>
#include <condition_variable>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace ndl {

class CopyPool {
 public:
  static CopyPool& instance() {
    static CopyPool pool;
    return pool;
  }
  int threads() const { return static_cast<int>(workers_.size()) + 1; }

  // memcpy(dst, src, bytes) split over the pool; returns when every piece is done.
  void copy(void* dst, const void* src, size_t bytes) {
    const size_t kMinPiece = 1u << 20;
    int pieces = static_cast<int>(std::min<size_t>(static_cast<size_t>(threads()), (bytes + kMinPiece - 1) / kMinPiece));
    if (pieces <= 1) {
      std::memcpy(dst, src, bytes);
      return;
    }
    // several callers (the replica threads of a multi-device pattern) may copy at once: every call counts its own pieces
    const size_t piece = ((bytes / static_cast<size_t>(pieces)) + 4095) & ~static_cast<size_t>(4095);
    int pending = 0;
    {
      std::lock_guard<std::mutex> lock(mutex_);
      for (int k = 1; k < pieces; k++) {
        const size_t lo = std::min(bytes, piece * static_cast<size_t>(k)), hi = std::min(bytes, lo + piece);
        if (hi > lo) {
          jobs_.push_back({static_cast<uint8_t*>(dst) + lo, static_cast<const uint8_t*>(src) + lo, hi - lo, &pending});
          pending++;
        }
      }
    }
    wake_.notify_all();
    std::memcpy(dst, src, std::min(bytes, piece));  // the calling thread takes the first piece
    std::unique_lock<std::mutex> lock(mutex_);
    done_.wait(lock, [&] { return pending == 0; });
  }

 private:
  struct Job {
    uint8_t* dst;
    const uint8_t* src;
    size_t bytes;
    int* pending;  // the call's count of unfinished pieces (guarded by mutex_)
  };
  CopyPool() {
    unsigned hw = std::thread::hardware_concurrency();
    int n = static_cast<int>(hw ? std::min(hw, 16u) : 4u) - 1;
    for (int i = 0; i < n; i++) workers_.emplace_back([this] { run(); });
  }
  ~CopyPool() {
    {
      std::lock_guard<std::mutex> lock(mutex_);
      stop_ = true;
    }
    wake_.notify_all();
    for (auto& t : workers_) t.join();
  }
  void run() {
    for (;;) {
      Job j;
      {
        std::unique_lock<std::mutex> lock(mutex_);
        wake_.wait(lock, [&] { return stop_ || !jobs_.empty(); });
        if (stop_ && jobs_.empty()) return;
        j = jobs_.back();
        jobs_.pop_back();
      }
      std::memcpy(j.dst, j.src, j.bytes);
      {
        std::lock_guard<std::mutex> lock(mutex_);
        if (--*j.pending == 0) done_.notify_all();
      }
    }
  }
  std::vector<std::thread> workers_;
  std::vector<Job> jobs_;
  std::mutex mutex_;
  std::condition_variable wake_, done_;
  bool stop_ = false;
};

}  // namespace ndl
