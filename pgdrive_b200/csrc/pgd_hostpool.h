// Host thread pool of the host-buffer step (pgd_hostpath.cu).  Pure C++: tools/sanitize_host_builds.sh compiles it with
// ThreadSanitizer / AddressSanitizer through oracle/hostpool_check.cpp.
#ifndef PGD_HOSTPOOL_H
#define PGD_HOSTPOOL_H
#include <stdint.h>

#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

// A pool of threads that spin while a step is in flight and sleep between steps.
// Every worker takes part in every job exactly once (it drains the item counter, then checks out) and run() returns only
// when all of them have checked out, so a job's descriptor is never read after run() has returned.
static inline void cpu_relax() {
#if defined(__x86_64__) || defined(__i386__)
  __builtin_ia32_pause();
#else
  std::this_thread::yield();
#endif
}

class HostPool {
 public:
  explicit HostPool(int workers) {
    for (int i = 0; i < workers; ++i) threads_.emplace_back([this] { worker(); });
  }
  ~HostPool() {
    {
      std::lock_guard<std::mutex> lk(m_);
      state_.store(2, std::memory_order_release);
    }
    cv_.notify_all();
    for (auto& t : threads_) t.join();
  }
  void begin() {  // wake the workers: they spin for jobs until end()
    {
      std::lock_guard<std::mutex> lk(m_);
      state_.store(1, std::memory_order_release);
    }
    cv_.notify_all();
  }
  void end() { state_.store(0, std::memory_order_release); }
  // fn(item) for item in [0, n): the caller takes part; returns when every item is done
  void run(int n, const std::function<void(int)>& fn) {
    fn_ = &fn;
    n_ = n;
    next_.store(0, std::memory_order_relaxed);
    checked_.store(0, std::memory_order_relaxed);
    job_.fetch_add(1, std::memory_order_release);
    drain();
    while (checked_.load(std::memory_order_acquire) < (int)threads_.size()) cpu_relax();
  }

 private:
  void drain() {
    for (;;) {
      const int i = next_.fetch_add(1, std::memory_order_acq_rel);
      if (i >= n_) return;
      (*fn_)(i);
    }
  }
  void worker() {
    uint64_t seen = 0;
    for (;;) {
      {
        std::unique_lock<std::mutex> lk(m_);
        cv_.wait(lk, [this] { return state_.load(std::memory_order_acquire) != 0; });
      }
      if (state_.load(std::memory_order_acquire) == 2) return;
      while (state_.load(std::memory_order_acquire) == 1) {
        const uint64_t j = job_.load(std::memory_order_acquire);
        if (j != seen) {
          seen = j;
          drain();
          checked_.fetch_add(1, std::memory_order_release);
        } else {
          cpu_relax();
        }
      }
    }
  }
  std::vector<std::thread> threads_;
  std::mutex m_;
  std::condition_variable cv_;
  std::atomic<int> state_{0};  // 0 asleep, 1 spinning for jobs, 2 stop
  const std::function<void(int)>* fn_ = nullptr;
  int n_ = 0;
  std::atomic<int> next_{0}, checked_{0};
  std::atomic<uint64_t> job_{0};
};

#endif
