// Host-side helper of the sample buffer (srl_b200/buffer.py): a multi-threaded memcpy into the pinned staging block.
// One numpy copy into pinned memory per sample was what bound DeviceSlabBuffer.put (7 GB/s on one host thread,
// profiles/r1c_notes.md); an Atari sample is a 3.6 MB frame leaf, so a few threads with a persistent pool (no thread
// start per call) move it at several times that.  Pure host code: no CUDA call, usable without a GPU.
#include <string.h>

#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

#include "common.cuh"

namespace srl {
namespace {

class CopyPool {
 public:
  // copies [src, src + bytes) to dst with `parts` participants: parts - 1 pool threads + the caller
  void run(unsigned char* dst, const unsigned char* src, size_t bytes, int parts) {
    std::lock_guard<std::mutex> one_copy_at_a_time(call_mu_);
    grow(parts - 1);
    const size_t chunk = ((bytes + parts - 1) / parts + 63) & ~static_cast<size_t>(63);  // 64-byte aligned cuts
    {
      std::lock_guard<std::mutex> lk(mu_);
      dst_ = dst, src_ = src, bytes_ = bytes, chunk_ = chunk;
      active_ = parts - 1;  // workers 0 .. parts-2 take chunks 1 .. parts-1
      pending_ = parts - 1;
      ++generation_;
    }
    cv_.notify_all();
    copy_chunk(0);
    std::unique_lock<std::mutex> lk(mu_);
    done_cv_.wait(lk, [&] { return pending_ == 0; });
  }

  ~CopyPool() {
    {
      std::lock_guard<std::mutex> lk(mu_);
      stop_ = true;
    }
    cv_.notify_all();
    for (auto& t : threads_) t.join();
  }

 private:
  void copy_chunk(int part) {
    const size_t lo = static_cast<size_t>(part) * chunk_;
    if (lo >= bytes_) return;
    const size_t n = bytes_ - lo < chunk_ ? bytes_ - lo : chunk_;
    memcpy(dst_ + lo, src_ + lo, n);
  }

  void grow(int workers) {
    while (static_cast<int>(threads_.size()) < workers) {
      const int id = static_cast<int>(threads_.size());
      unsigned long long seen;
      {
        std::lock_guard<std::mutex> lk(mu_);
        seen = generation_;
      }
      threads_.emplace_back([this, id, seen]() mutable {
        for (;;) {
          std::unique_lock<std::mutex> lk(mu_);
          cv_.wait(lk, [&] { return stop_ || generation_ != seen; });
          if (stop_) return;
          seen = generation_;
          const bool mine = id < active_;
          lk.unlock();
          if (mine) {
            copy_chunk(id + 1);
            lk.lock();
            if (--pending_ == 0) done_cv_.notify_one();
          }
        }
      });
    }
  }

  std::mutex call_mu_, mu_;
  std::condition_variable cv_, done_cv_;
  std::vector<std::thread> threads_;
  unsigned char* dst_ = nullptr;
  const unsigned char* src_ = nullptr;
  size_t bytes_ = 0, chunk_ = 0;
  int active_ = 0, pending_ = 0;
  unsigned long long generation_ = 0;
  bool stop_ = false;
};

CopyPool& pool() {
  static CopyPool* p = new CopyPool();  // leaked on purpose: no join at process exit (threads may outlive statics)
  return *p;
}

}  // namespace
}  // namespace srl

extern "C" int srl_host_copy(void* dst, const void* src, size_t bytes, int threads) {
  using namespace srl;
  if (bytes == 0) return SRL_OK;
  SRL_REQUIRE(dst != nullptr && src != nullptr, SRL_ERR_INVALID_ARG, "srl_host_copy: null pointer");
  SRL_REQUIRE(threads >= 1 && threads <= 64, SRL_ERR_INVALID_ARG, "srl_host_copy: threads=%d outside [1, 64]", threads);
  // below ~256 KB per participant the hand-over costs more than it saves
  int parts = threads;
  const size_t min_part = 256 * 1024;
  if (static_cast<size_t>(parts) * min_part > bytes) parts = static_cast<int>(bytes / min_part);
  if (parts <= 1) {
    memcpy(dst, src, bytes);
    return SRL_OK;
  }
  pool().run(static_cast<unsigned char*>(dst), static_cast<const unsigned char*>(src), bytes, parts);
  return SRL_OK;
}
