// ORACLE SCAFFOLDING (test infrastructure): a tiny std::thread stand-in for the handful of oneTBB
// entry points the reference's placement / index translation units use (oneTBB is absent here).
// With max_allowed_parallelism == 1 (panmap's default `-t 1`) everything runs inline on the calling
// thread, which is the canonical order the parity tests pin.  With N > 1 parallel_for fans out over
// N std::threads (used only for the "--impl reference" CPU timing arm).
#pragma once
#include <algorithm>
#include <atomic>
#include <cstddef>
#include <deque>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>
namespace tbb {
namespace shim {
inline std::atomic<size_t>& parallelism() { static std::atomic<size_t> p{1}; return p; }
inline int& thread_index() { thread_local int idx = 0; return idx; }
inline bool& in_parallel() { thread_local bool f = false; return f; }
template <class ChunkFn>
inline void run_chunks(size_t nChunks, ChunkFn&& fn) {
    const size_t n = std::min(parallelism().load(), nChunks);
    if (n <= 1 || in_parallel()) { for (size_t c = 0; c < nChunks; ++c) fn(c); return; }
    std::atomic<size_t> next{0};
    auto worker = [&](int id) {
        thread_index() = id; in_parallel() = true;
        for (size_t c; (c = next.fetch_add(1)) < nChunks;) fn(c);
        in_parallel() = false; thread_index() = 0;
    };
    std::vector<std::thread> ts;
    for (size_t i = 1; i < n; ++i) ts.emplace_back(worker, static_cast<int>(i));
    worker(0);
    for (auto& t : ts) t.join();
}
}  // namespace shim

class global_control {
   public:
    enum parameter { max_allowed_parallelism, thread_stack_size };
    global_control(parameter p, size_t v) : p_(p) {
        if (p == max_allowed_parallelism) { prev_ = shim::parallelism(); shim::parallelism() = std::max<size_t>(1, v); }
    }
    ~global_control() { if (p_ == max_allowed_parallelism) shim::parallelism() = prev_; }
    static size_t active_value(parameter p) { return p == max_allowed_parallelism ? shim::parallelism().load() : 0; }
   private:
    parameter p_; size_t prev_ = 1;
};

template <class T>
class blocked_range {
    T b_, e_; size_t g_;
   public:
    blocked_range(T b, T e, size_t g = 1) : b_(b), e_(e), g_(g ? g : 1) {}
    T begin() const { return b_; }
    T end() const { return e_; }
    size_t size() const { return static_cast<size_t>(e_ - b_); }
    size_t grainsize() const { return g_; }
    bool empty() const { return !(b_ < e_); }
};

template <class T, class Body>
inline void parallel_for(const blocked_range<T>& r, const Body& body) {
    if (r.empty()) return;
    const size_t n = shim::parallelism();
    if (n <= 1 || shim::in_parallel()) { body(r); return; }
    const size_t total = r.size();
    size_t chunk = std::max(r.grainsize(), (total + n * 8 - 1) / (n * 8));
    const size_t nChunks = (total + chunk - 1) / chunk;
    shim::run_chunks(nChunks, [&](size_t c) {
        T b = r.begin() + static_cast<T>(c * chunk);
        T e = r.begin() + static_cast<T>(std::min(total, (c + 1) * chunk));
        body(blocked_range<T>(b, e, r.grainsize()));
    });
}
template <class I, class F>
inline void parallel_for(I first, I last, const F& f) {
    if (!(first < last)) return;
    shim::run_chunks(static_cast<size_t>(last - first), [&](size_t c) { f(static_cast<I>(first + c)); });
}
template <class It, class Cmp>
inline void parallel_sort(It b, It e, Cmp c) { std::sort(b, e, c); }
template <class It>
inline void parallel_sort(It b, It e) { std::sort(b, e); }

template <class T>
class enumerable_thread_specific {
    mutable std::mutex m_;
    std::vector<std::pair<int, std::unique_ptr<T>>> slots_;  // in creation order
    struct iter {
        typename std::vector<std::pair<int, std::unique_ptr<T>>>::iterator it;
        T& operator*() const { return *it->second; }
        T* operator->() const { return it->second.get(); }
        iter& operator++() { ++it; return *this; }
        bool operator!=(const iter& o) const { return it != o.it; }
        bool operator==(const iter& o) const { return it == o.it; }
    };
   public:
    T& local() {
        const int id = shim::thread_index();
        std::lock_guard<std::mutex> g(m_);
        for (auto& s : slots_) if (s.first == id) return *s.second;
        slots_.emplace_back(id, std::make_unique<T>());
        return *slots_.back().second;
    }
    size_t size() const { std::lock_guard<std::mutex> g(m_); return slots_.size(); }
    iter begin() { return iter{slots_.begin()}; }
    iter end() { return iter{slots_.end()}; }
};

namespace this_task_arena {
inline int current_thread_index() { return shim::thread_index(); }
inline int max_concurrency() { return static_cast<int>(shim::parallelism().load()); }
}  // namespace this_task_arena

class task_arena {
   public:
    explicit task_arena(int = 0) {}
    template <class F> void execute(F&& f) { f(); }
};
class task_group {
   public:
    template <class F> void run(F&& f) { f(); }
    void wait() {}
};
class spin_mutex {
    std::mutex m_;
   public:
    void lock() { m_.lock(); }
    void unlock() { m_.unlock(); }
    class scoped_lock {
        spin_mutex* m_ = nullptr;
       public:
        scoped_lock() = default;
        explicit scoped_lock(spin_mutex& m) : m_(&m) { m_->lock(); }
        ~scoped_lock() { if (m_) m_->unlock(); }
    };
};
template <class T>
class concurrent_vector : public std::deque<T> {
   public:
    using std::deque<T>::deque;
};
template <class K, class V> class concurrent_hash_map {};
}  // namespace tbb
