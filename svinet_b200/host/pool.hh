// pool.hh -- a few persistent host threads for the start-up passes (thousands of short parallel phases:
// spawning 16 threads per phase cost more than the phases themselves).
#ifndef SVINET_B200_POOL_HH
#define SVINET_B200_POOL_HH

#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

class Pool {
 public:
  explicit Pool(unsigned nt) : nt_(nt ? nt : 1) {
    for (unsigned t = 1; t < nt_; ++t) th_.emplace_back([this, t] { loop(t); });
  }
  ~Pool() {
    {
      std::lock_guard<std::mutex> g(m_);
      stop_ = true;
      ++gen_;
    }
    cv_.notify_all();
    for (auto &x : th_) x.join();
  }
  unsigned size() const { return nt_; }
  // runs fn(t) for t = 0..size()-1 (t = 0 on the caller) and returns when all are done
  void run(const std::function<void(unsigned)> &fn) {
    {
      std::lock_guard<std::mutex> g(m_);
      fn_ = &fn;
      left_ = nt_ - 1;
      ++gen_;
    }
    cv_.notify_all();
    fn(0);
    std::unique_lock<std::mutex> g(m_);
    done_.wait(g, [this] { return left_ == 0; });
    fn_ = nullptr;
  }

 private:
  void loop(unsigned t) {
    unsigned long seen = 0;
    for (;;) {
      const std::function<void(unsigned)> *fn;
      {
        std::unique_lock<std::mutex> g(m_);
        cv_.wait(g, [&] { return gen_ != seen; });
        seen = gen_;
        if (stop_) return;
        fn = fn_;
      }
      (*fn)(t);
      {
        std::lock_guard<std::mutex> g(m_);
        if (--left_ == 0) done_.notify_one();
      }
    }
  }
  unsigned nt_;
  std::vector<std::thread> th_;
  std::mutex m_;
  std::condition_variable cv_, done_;
  const std::function<void(unsigned)> *fn_ = nullptr;
  unsigned left_ = 0;
  unsigned long gen_ = 0;
  bool stop_ = false;
};

#endif
