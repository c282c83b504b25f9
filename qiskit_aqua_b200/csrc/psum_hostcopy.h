// psum_hostcopy.h -- host side of the pageable-input path: a persistent pool of copy threads that
// moves a caller's pageable state into the pinned staging ring ahead of the DMA.
//
// The reference API receives a pageable numpy array (legacy/weighted_pauli_operator.py:605
// evaluate_with_statevector(quantum_state)); the copy engines only reach full PCIe rate from pinned
// memory, so the state crosses a ring of pinned slots.  The host copy into a slot uses streaming
// (non-temporal) stores: the slot is written once and next read by the DMA engine, so pulling its
// lines into the cache first (read-for-ownership) would add a third of avoidable memory traffic
// while the DMA competes for the same memory controllers.
#pragma once
#include <emmintrin.h>
#include <stddef.h>
#include <string.h>

#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

namespace psum {

// dst 16-byte aligned; src arbitrary
inline void stream_copy(char* dst, const char* src, size_t n) {
  size_t i = 0;
  if (((uintptr_t)dst & 15u) == 0) {
    for (; i + 64 <= n; i += 64) {
      const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i));
      const __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i + 16));
      const __m128i c = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i + 32));
      const __m128i d = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i + 48));
      _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i), a);
      _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i + 16), b);
      _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i + 32), c);
      _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i + 48), d);
    }
    _mm_sfence();
  }
  if (i < n) memcpy(dst + i, src + i, n - i);
}

class CopyPool {
 public:
  explicit CopyPool(int nthreads) : n_(nthreads < 1 ? 1 : nthreads) {
    for (int t = 1; t < n_; ++t) workers_.emplace_back([this, t]() { run(t); });
  }
  ~CopyPool() {
    {
      std::lock_guard<std::mutex> lk(m_);
      stop_ = true;
      ++gen_;
    }
    cv_.notify_all();
    for (auto& w : workers_) w.join();
  }
  CopyPool(const CopyPool&) = delete;
  CopyPool& operator=(const CopyPool&) = delete;
  int threads() const { return n_; }

  // dst[0..bytes) <- src, split into page-aligned slices, one per thread; returns when all are done
  void copy(char* dst, const char* src, size_t bytes) {
    const size_t slice = ((bytes + n_ - 1) / n_ + 4095) & ~(size_t)4095;
    {
      std::lock_guard<std::mutex> lk(m_);
      dst_ = dst;
      src_ = src;
      bytes_ = bytes;
      slice_ = slice;
      pending_ = n_ - 1;
      ++gen_;
    }
    cv_.notify_all();
    do_slice(0, dst, src, bytes, slice);
    std::unique_lock<std::mutex> lk(m_);
    done_.wait(lk, [this]() { return pending_ == 0; });
  }

 private:
  static void do_slice(int t, char* dst, const char* src, size_t bytes, size_t slice) {
    const size_t o = (size_t)t * slice;
    if (o < bytes) stream_copy(dst + o, src + o, bytes - o < slice ? bytes - o : slice);
  }
  void run(int t) {
    unsigned long seen = 0;
    for (;;) {
      char* dst;
      const char* src;
      size_t bytes, slice;
      {
        std::unique_lock<std::mutex> lk(m_);
        cv_.wait(lk, [&]() { return gen_ != seen; });
        seen = gen_;
        if (stop_) return;
        dst = dst_;
        src = src_;
        bytes = bytes_;
        slice = slice_;
      }
      do_slice(t, dst, src, bytes, slice);
      {
        std::lock_guard<std::mutex> lk(m_);
        if (--pending_ == 0) done_.notify_one();
      }
    }
  }
  const int n_;
  std::vector<std::thread> workers_;
  std::mutex m_;
  std::condition_variable cv_, done_;
  unsigned long gen_ = 0;
  bool stop_ = false;
  char* dst_ = nullptr;
  const char* src_ = nullptr;
  size_t bytes_ = 0, slice_ = 0;
  int pending_ = 0;
};

}  // namespace psum
