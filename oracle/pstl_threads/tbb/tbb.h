/* oracle/pstl_threads/tbb/tbb.h -- TEST INFRASTRUCTURE (never part of the product).
 *
 * The reference parallelises its step with std::for_each(std::execution::par, ...) (physicsWorld.cc:42,63,72,81,306,
 * 473,486).  libstdc++ runs that in parallel only when Intel TBB is installed (c++config.h: _GLIBCXX_USE_TBB_PAR_BACKEND
 * = __has_include(<tbb/tbb.h>)); without it every `par` loop runs serially, which is why oracle/_ref is a one-core
 * baseline.  TBB is not in this image.  This header is found by that __has_include and gives libstdc++'s backend
 * (pstl/parallel_backend_tbb.h) just enough of TBB's classic interface to compile -- with a parallel_for that really
 * runs on all host threads (OpenMP) and serial stand-ins for everything the reference never calls -- so the UNMODIFIED
 * reference can be timed "with all the host threads it can use" (oracle/_ref/libsph_ref_par.so, bench.py --impl
 * reference).  Results of that build are used for timing and for race-free stages only: the reference's in-place
 * viscosity update is a data race under a parallel backend.
 *
 * Written against the interface libstdc++ 13 expects; none of this is TBB code. */
#ifndef SPH_ORACLE_PSTL_THREADS_TBB_H
#define SPH_ORACLE_PSTL_THREADS_TBB_H

#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <new>
#include <utility>
#include <omp.h>

#define TBB_INTERFACE_VERSION 11000   /* the classic tbb::task branch of parallel_backend_tbb.h */

namespace tbb {

struct split {};

template <typename Value>
class blocked_range {
public:
    using const_iterator = Value;
    using size_type = std::size_t;
    blocked_range(Value b, Value e, size_type grain = 1) : b_(b), e_(e), grain_(grain) {}
    blocked_range(blocked_range& r, split) : b_(r.b_), e_(r.e_), grain_(r.grain_)
    {
        Value mid = r.b_ + (r.e_ - r.b_) / 2;
        b_ = mid; r.e_ = mid;
    }
    Value begin() const { return b_; }
    Value end() const { return e_; }
    size_type size() const { return (size_type)(e_ - b_); }
    size_type grainsize() const { return grain_; }
    bool empty() const { return !(b_ < e_); }
    bool is_divisible() const { return grain_ < size(); }
private:
    Value b_, e_;
    size_type grain_;
};

/* the one entry point the reference's loops reach: chunks of the range on every OpenMP thread */
template <typename Range, typename Body>
void parallel_for(const Range& range, const Body& body)
{
    const auto b = range.begin();
    const std::size_t n = range.size();
    if (n == 0) return;
    const std::size_t threads = (std::size_t)omp_get_max_threads();
    std::size_t chunk = n / (threads * 8) + 1;                 /* ~8 chunks per thread: dynamic balance, little overhead */
    if (chunk < 256) chunk = 256;
    const std::size_t chunks = (n + chunk - 1) / chunk;
    if (chunks <= 1 || threads <= 1 || omp_in_parallel()) { body(range); return; }
    #pragma omp parallel for schedule(dynamic, 1)
    for (std::size_t c = 0; c < chunks; c++) {
        const std::size_t lo = c * chunk, hi = lo + chunk < n ? lo + chunk : n;
        body(Range(b + lo, b + hi, range.grainsize()));
    }
}

/* serial stand-ins (not reached by the reference; present so the backend header compiles and stays correct) */
template <typename Range, typename Body>
void parallel_reduce(const Range& range, Body& body) { body(range); }

struct pre_scan_tag { static bool is_final_scan() { return false; } operator bool() const { return false; } };
struct final_scan_tag { static bool is_final_scan() { return true; } operator bool() const { return true; } };
template <typename Range, typename Body>
void parallel_scan(const Range& range, Body& body) { body(range, final_scan_tag()); }

template <typename F0, typename F1>
void parallel_invoke(const F0& f0, const F1& f1) { f0(); f1(); }

namespace this_task_arena {
template <typename F>
auto isolate(const F& f) -> decltype(f()) { return f(); }
inline int max_concurrency() { return omp_get_max_threads(); }
}

template <typename T>
class tbb_allocator : public std::allocator<T> {
public:
    template <typename U> struct rebind { using other = tbb_allocator<U>; };
    tbb_allocator() = default;
    template <typename U> tbb_allocator(const tbb_allocator<U>&) {}
};
template <typename T> using allocator = tbb_allocator<T>;

struct task_group_context {
    bool cancel_group_execution() { return false; }
};

/* The classic task interface: declarations only, so the backend header compiles.  Only libstdc++'s parallel stable
 * sort / merge use it, the reference calls neither (its one sort is a plain std::sort, physicsWorld.cc:484); reaching
 * it aborts loudly instead of running a task graph this header does not implement. */
namespace internal {
struct allocate_proxy {};
}
class task {
public:
    virtual ~task() = default;
    virtual task* execute() = 0;
    static task& self() { static struct idle_task : task { task* execute() override { return nullptr; } } t; return t; }
    task_group_context* group() { static task_group_context ctx; return &ctx; }
    static internal::allocate_proxy allocate_root() { return {}; }
    internal::allocate_proxy allocate_continuation() { return {}; }
    internal::allocate_proxy allocate_child() { return {}; }
    static internal::allocate_proxy allocate_additional_child_of(task&) { return {}; }
    task* parent() const { return nullptr; }
    void recycle_as_continuation() {}
    void recycle_as_child_of(task&) {}
    void set_ref_count(int) {}
    static void unsupported()
    {
        std::fputs("oracle/pstl_threads: parallel sort / merge (tbb::task) is not implemented by this stand-in\n", stderr);
        std::abort();
    }
    static void spawn(task&) { unsupported(); }
    static void spawn_root_and_wait(task&) { unsupported(); }
};

}  // namespace tbb

inline void* operator new(std::size_t bytes, const tbb::internal::allocate_proxy&) { return ::operator new(bytes); }
inline void operator delete(void* p, const tbb::internal::allocate_proxy&) { ::operator delete(p); }

#endif
