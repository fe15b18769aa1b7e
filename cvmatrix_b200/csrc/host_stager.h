// Pageable host memory -> device at (close to) the PCIe rate.
//
// A drop-in user hands CVMatrix.fit ordinary numpy arrays.  cudaMemcpyAsync from pageable memory is staged by the
// driver through its own bounce buffer on the calling thread: ~10 GB/s measured on this pool's boxes, against 55 GB/s
// from page-locked memory - fit of BASELINE config 2 (4.1 GB) takes 390 ms instead of 75 ms.  HostStager owns a small
// ring of page-locked buffers and a few worker threads: every piece of the source is copied into a ring buffer by
// all threads at once (a multi-threaded memcpy runs at memory bandwidth), handed to the DMA engine with
// cudaMemcpyAsync, and the next piece is copied while that transfer is in flight.  Like the driver's own path, copy()
// returns once the source has been read completely; the device side stays asynchronous on `stream`.
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

namespace cvmx {

class HostStager {
 public:
  static constexpr size_t PIECE = (size_t)16 << 20;   // bytes per ring buffer
  static constexpr int NBUF = 4;

  HostStager() {
    unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    nthreads_ = (int)std::min(8u, std::max(1u, hw / 2));
    if (const char* e = std::getenv("CVMX_STAGE_THREADS")) nthreads_ = std::max(1, std::min(64, std::atoi(e)));
    for (int i = 0; i < NBUF; ++i) { pin_[i] = nullptr; done_[i] = nullptr; used_[i] = false; }
  }
  ~HostStager() {
    {
      std::lock_guard<std::mutex> lk(m_);
      stop_ = true; ++gen_;
    }
    cv_start_.notify_all();
    for (auto& t : workers_) t.join();
    for (int i = 0; i < NBUF; ++i) {
      if (done_[i]) { cudaEventSynchronize(done_[i]); cudaEventDestroy(done_[i]); }
      if (pin_[i]) cudaFreeHost(pin_[i]);
    }
  }
  HostStager(const HostStager&) = delete;
  HostStager& operator=(const HostStager&) = delete;

  // true when `p` is ordinary (not page-locked, not device / managed) host memory
  static bool pageable(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return true; }
    return a.type == cudaMemoryTypeUnregistered;
  }

  cudaError_t copy(void* dst_dev, const void* src_host, size_t bytes, cudaStream_t stream) {
    cudaError_t e = init();
    if (e != cudaSuccess) return e;
    const char* src = static_cast<const char*>(src_host);
    char* dst = static_cast<char*>(dst_dev);
    for (size_t off = 0; off < bytes; off += PIECE) {
      const size_t n = std::min(PIECE, bytes - off);
      const int b = next_++ % NBUF;
      if (used_[b] && (e = cudaEventSynchronize(done_[b])) != cudaSuccess) return e;   // its last transfer has left the buffer
      parallel_memcpy(static_cast<char*>(pin_[b]), src + off, n);
      if ((e = cudaMemcpyAsync(dst + off, pin_[b], n, cudaMemcpyHostToDevice, stream)) != cudaSuccess) return e;
      if ((e = cudaEventRecord(done_[b], stream)) != cudaSuccess) return e;
      used_[b] = true;
    }
    return cudaSuccess;
  }

 private:
  cudaError_t init() {
    if (ready_) return cudaSuccess;
    for (int i = 0; i < NBUF; ++i) {
      cudaError_t e = cudaHostAlloc(&pin_[i], PIECE, cudaHostAllocDefault);
      if (e != cudaSuccess) return e;
      e = cudaEventCreateWithFlags(&done_[i], cudaEventDisableTiming);
      if (e != cudaSuccess) return e;
    }
    try {
      for (int t = 1; t < nthreads_; ++t) workers_.emplace_back([this, t] { worker(t); });
    } catch (...) {
      nthreads_ = (int)workers_.size() + 1;   // fewer threads than asked for: slices are cut for the ones that exist
    }
    ready_ = true;
    return cudaSuccess;
  }

  void slice(int t, char* dst, const char* src, size_t n) const {
    const size_t per = ((n + nthreads_ - 1) / nthreads_ + 4095) / 4096 * 4096;
    const size_t a = std::min(n, per * (size_t)t), z = std::min(n, per * (size_t)(t + 1));
    if (z > a) std::memcpy(dst + a, src + a, z - a);
  }

  void parallel_memcpy(char* dst, const char* src, size_t n) {
    if (nthreads_ == 1 || n < ((size_t)1 << 20)) { std::memcpy(dst, src, n); return; }
    {
      std::lock_guard<std::mutex> lk(m_);
      job_dst_ = dst; job_src_ = src; job_n_ = n; pending_ = nthreads_ - 1; ++gen_;
    }
    cv_start_.notify_all();
    slice(0, dst, src, n);                       // the calling thread takes the first slice
    std::unique_lock<std::mutex> lk(m_);
    cv_done_.wait(lk, [this] { return pending_ == 0; });
  }

  void worker(int t) {
    unsigned long long seen = 0;
    for (;;) {
      char* dst; const char* src; size_t n;
      {
        std::unique_lock<std::mutex> lk(m_);
        cv_start_.wait(lk, [&] { return gen_ != seen; });
        seen = gen_;
        if (stop_) return;
        dst = job_dst_; src = job_src_; n = job_n_;
      }
      slice(t, dst, src, n);
      {
        std::lock_guard<std::mutex> lk(m_);
        if (--pending_ == 0) cv_done_.notify_one();
      }
    }
  }

  int nthreads_ = 1;
  bool ready_ = false, stop_ = false;
  void* pin_[NBUF];
  cudaEvent_t done_[NBUF];
  bool used_[NBUF];
  unsigned next_ = 0;
  std::vector<std::thread> workers_;
  std::mutex m_;
  std::condition_variable cv_start_, cv_done_;
  unsigned long long gen_ = 0;
  int pending_ = 0;
  char* job_dst_ = nullptr;
  const char* job_src_ = nullptr;
  size_t job_n_ = 0;
};

}  // namespace cvmx
