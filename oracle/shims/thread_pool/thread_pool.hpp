// TEST INFRASTRUCTURE.  Stand-in for rvaser/thread_pool 4.0.0 (a FetchContent dependency of the reference,
// CMakeLists.txt:61-76, whose source is not in the reference tree and cannot be fetched here).  Exposes exactly
// what src/polisher.cpp uses (:183 ctor, :376/:471/:499/:510 Submit, :501/:512 thread_map): a fixed pool of
// workers, FIFO task queue, futures.  Written from the published interface; nothing on the arithmetic of the path.
#ifndef ORACLE_SHIM_THREAD_POOL_HPP_
#define ORACLE_SHIM_THREAD_POOL_HPP_

#include <condition_variable>
#include <cstdint>
#include <functional>
#include <future>
#include <memory>
#include <mutex>
#include <queue>
#include <thread>
#include <type_traits>
#include <unordered_map>
#include <vector>

namespace thread_pool {

class ThreadPool {
 public:
  explicit ThreadPool(std::size_t num_threads = std::thread::hardware_concurrency()) : stop_(false) {
    if (num_threads == 0) num_threads = 1;
    for (std::size_t i = 0; i < num_threads; ++i) {
      workers_.emplace_back([this] { this->loop(); });
      ids_[workers_.back().get_id()] = static_cast<std::uint32_t>(i);
    }
  }
  ThreadPool(const ThreadPool&) = delete;
  ThreadPool& operator=(const ThreadPool&) = delete;
  ~ThreadPool() {
    {
      std::lock_guard<std::mutex> g(m_);
      stop_ = true;
    }
    cv_.notify_all();
    for (auto& t : workers_) t.join();
  }

  std::size_t num_threads() const { return workers_.size(); }
  const std::unordered_map<std::thread::id, std::uint32_t>& thread_map() const { return ids_; }

  template <typename F, typename... Args>
  auto Submit(F&& f, Args&&... args) -> std::future<typename std::result_of<F(Args...)>::type> {
    using R = typename std::result_of<F(Args...)>::type;
    auto task = std::make_shared<std::packaged_task<R()>>(std::bind(std::forward<F>(f), std::forward<Args>(args)...));
    std::future<R> fut = task->get_future();
    {
      std::lock_guard<std::mutex> g(m_);
      q_.emplace([task] { (*task)(); });
    }
    cv_.notify_one();
    return fut;
  }

 private:
  void loop() {
    for (;;) {
      std::function<void()> job;
      {
        std::unique_lock<std::mutex> g(m_);
        cv_.wait(g, [this] { return stop_ || !q_.empty(); });
        if (q_.empty()) return;
        job = std::move(q_.front());
        q_.pop();
      }
      job();
    }
  }
  std::vector<std::thread> workers_;
  std::unordered_map<std::thread::id, std::uint32_t> ids_;
  std::queue<std::function<void()>> q_;
  std::mutex m_;
  std::condition_variable cv_;
  bool stop_;
};

}  // namespace thread_pool
#endif
