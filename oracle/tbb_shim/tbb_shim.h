// tbb_shim.h -- TEST INFRASTRUCTURE ONLY (part of oracle/, never linked into the product).
//
// A tiny stand-in for the subset of oneTBB that OpenVDB 13 names, so that the
// UNMODIFIED reference sources under /root/reference compile in an image that
// ships no TBB.  Everything runs on the calling thread except the 1-D
// parallel_for(blocked_range, body), which fans out over std::thread workers
// when tbb_shim::set_num_threads(n>1) was called: each worker copy-constructs
// the body (as TBB tasks do) and claims chunks of the range dynamically.  That
// is all the timed CPU baseline (LevelSetRayTracer::render(true),
// VolumeRender::render(true)) needs.
#pragma once
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstddef>
#include <deque>
#include <functional>
#include <iterator>
#include <list>
#include <map>
#include <memory>
#include <mutex>
#include <thread>
#include <type_traits>
#include <unordered_map>
#include <utility>
#include <vector>

#define TBB_VERSION_MAJOR 2021
#define TBB_VERSION_MINOR 5
#define TBB_INTERFACE_VERSION 12050

namespace tbb_shim {
inline std::atomic<int>& num_threads_ref() { static std::atomic<int> n{1}; return n; }
inline void set_num_threads(int n) { num_threads_ref() = n < 1 ? 1 : n; }
inline int  num_threads() { return num_threads_ref().load(); }
// true while a fan-out is running: nested parallel constructs stay serial
inline std::atomic<bool>& in_parallel() { static std::atomic<bool> f{false}; return f; }
}

namespace tbb {

struct split {};
struct auto_partitioner {}; struct simple_partitioner {}; struct static_partitioner {};
struct affinity_partitioner {};

template<typename T>
class blocked_range {
public:
    using const_iterator = T; using size_type = std::size_t;
    blocked_range() = default;
    blocked_range(T b, T e, size_type g = 1) : mB(b), mE(e), mG(g) {}
    blocked_range(blocked_range& r, split) : mB(r.mB), mE(r.mE), mG(r.mG) { r.mB = r.mE; }
    T begin() const { return mB; }
    T end() const { return mE; }
    size_type grainsize() const { return mG; }
    bool empty() const { return mB == mE || !(mB != mE); }
    size_type size() const { return size_type(dist(mB, mE, 0)); }
    bool is_divisible() const { return false; }
private:
    template<typename U> static auto dist(U a, U b, int) -> decltype(std::size_t(b - a)) { return std::size_t(b - a); }
    template<typename U> static std::size_t dist(U a, U b, long) { return std::size_t(std::distance(a, b)); }
    T mB{}, mE{}; size_type mG = 1;
};

template<typename RowT, typename ColT = RowT>
class blocked_range2d {
public:
    using row_range_type = blocked_range<RowT>; using col_range_type = blocked_range<ColT>;
    blocked_range2d(RowT rb, RowT re, std::size_t rg, ColT cb, ColT ce, std::size_t cg) : mR(rb, re, rg), mC(cb, ce, cg) {}
    blocked_range2d(RowT rb, RowT re, ColT cb, ColT ce) : mR(rb, re), mC(cb, ce) {}
    blocked_range2d(blocked_range2d& r, split) : mR(r.mR), mC(r.mC) {}
    bool empty() const { return mR.empty() || mC.empty(); }
    bool is_divisible() const { return false; }
    const row_range_type& rows() const { return mR; }
    const col_range_type& cols() const { return mC; }
private: row_range_type mR; col_range_type mC;
};

template<typename PageT, typename RowT = PageT, typename ColT = RowT>
class blocked_range3d {
public:
    using page_range_type = blocked_range<PageT>; using row_range_type = blocked_range<RowT>; using col_range_type = blocked_range<ColT>;
    blocked_range3d(PageT pb, PageT pe, RowT rb, RowT re, ColT cb, ColT ce) : mP(pb, pe), mR(rb, re), mC(cb, ce) {}
    blocked_range3d(PageT pb, PageT pe, std::size_t pg, RowT rb, RowT re, std::size_t rg, ColT cb, ColT ce, std::size_t cg)
        : mP(pb, pe, pg), mR(rb, re, rg), mC(cb, ce, cg) {}
    blocked_range3d(blocked_range3d& r, split) : mP(r.mP), mR(r.mR), mC(r.mC) {}
    bool empty() const { return mP.empty() || mR.empty() || mC.empty(); }
    bool is_divisible() const { return false; }
    const page_range_type& pages() const { return mP; }
    const row_range_type& rows() const { return mR; }
    const col_range_type& cols() const { return mC; }
private: page_range_type mP; row_range_type mR; col_range_type mC;
};

namespace shim_detail {
template<typename RangeT, typename BodyT>
inline void run_serial(const RangeT& range, const BodyT& body) {
    RangeT r(range);            // some bodies take Range& : hand them a mutable copy
    if (!r.empty()) body(r);
}
template<typename T, typename BodyT>
inline auto run_threaded(const blocked_range<T>& range, const BodyT& body, int)
    -> typename std::enable_if<std::is_integral<T>::value && std::is_copy_constructible<BodyT>::value>::type
{
    const int nt = tbb_shim::num_threads();
    const std::size_t n = range.size();
    bool expected = false;
    if (nt <= 1 || n < 2 || !tbb_shim::in_parallel().compare_exchange_strong(expected, true)) {
        run_serial(range, body); return;
    }
    const std::size_t chunk = std::max<std::size_t>(range.grainsize(), std::max<std::size_t>(1, n / (std::size_t(nt) * 16)));
    std::atomic<std::size_t> next{0};
    const T b0 = range.begin();
    auto worker = [&]() {
        BodyT local(body);      // TBB copy-constructs the body per task
        for (;;) {
            const std::size_t s = next.fetch_add(chunk);
            if (s >= n) break;
            const std::size_t e = std::min(n, s + chunk);
            blocked_range<T> sub(T(b0 + T(s)), T(b0 + T(e)), range.grainsize());
            local(sub);
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < nt; ++t) pool.emplace_back(worker);
    worker();
    for (auto& th : pool) th.join();
    tbb_shim::in_parallel() = false;
}
template<typename RangeT, typename BodyT>
inline void run_threaded(const RangeT& range, const BodyT& body, long) { run_serial(range, body); }
} // namespace shim_detail

template<typename RangeT, typename BodyT>
inline void parallel_for(const RangeT& range, const BodyT& body) { shim_detail::run_threaded(range, body, 0); }
template<typename RangeT, typename BodyT, typename PartT,
         typename = decltype(std::declval<RangeT>().empty())>
inline void parallel_for(const RangeT& range, const BodyT& body, const PartT&) { shim_detail::run_threaded(range, body, 0); }
template<typename RangeT, typename BodyT>
inline void parallel_for(const RangeT& range, const BodyT& body, affinity_partitioner&) { shim_detail::run_threaded(range, body, 0); }
template<typename IndexT, typename FuncT,
         typename = typename std::enable_if<std::is_integral<IndexT>::value>::type>
inline void parallel_for(IndexT first, IndexT last, const FuncT& f) { for (IndexT i = first; i < last; ++i) f(i); }

// imperative form: body(range) accumulates into itself; join() never needed serially
template<typename RangeT, typename BodyT>
inline void parallel_reduce(const RangeT& range, BodyT& body) { RangeT r(range); if (!r.empty()) body(r); }
template<typename RangeT, typename BodyT, typename PartT,
         typename = decltype(std::declval<BodyT&>().join(std::declval<BodyT&>()))>
inline void parallel_reduce(const RangeT& range, BodyT& body, const PartT&) { RangeT r(range); if (!r.empty()) body(r); }
// functional form
template<typename RangeT, typename ValueT, typename FuncT, typename ReduceT>
inline ValueT parallel_reduce(const RangeT& range, const ValueT& identity, const FuncT& func, const ReduceT&)
{ RangeT r(range); if (r.empty()) return identity; return func(r, identity); }
template<typename RangeT, typename ValueT, typename FuncT, typename ReduceT, typename PartT>
inline ValueT parallel_reduce(const RangeT& range, const ValueT& identity, const FuncT& func, const ReduceT&, const PartT&)
{ RangeT r(range); if (r.empty()) return identity; return func(r, identity); }

template<typename It> inline void parallel_sort(It b, It e) { std::sort(b, e); }
template<typename It, typename Cmp> inline void parallel_sort(It b, It e, const Cmp& c) { std::sort(b, e, c); }
template<typename C> inline void parallel_sort(C& c) { std::sort(c.begin(), c.end()); }

template<typename... Fs> inline void parallel_invoke(Fs&&... fs) { (void)std::initializer_list<int>{ (std::forward<Fs>(fs)(), 0)... }; }

class spin_mutex {
public:
    spin_mutex() = default;
    spin_mutex(const spin_mutex&) = delete;
    void lock() { while (mF.test_and_set(std::memory_order_acquire)) std::this_thread::yield(); }
    void unlock() { mF.clear(std::memory_order_release); }
    bool try_lock() { return !mF.test_and_set(std::memory_order_acquire); }
    class scoped_lock {
    public:
        scoped_lock() = default;
        explicit scoped_lock(spin_mutex& m) : mM(&m) { m.lock(); }
        ~scoped_lock() { if (mM) mM->unlock(); }
        void acquire(spin_mutex& m) { m.lock(); mM = &m; }
        void release() { if (mM) { mM->unlock(); mM = nullptr; } }
    private: spin_mutex* mM = nullptr;
    };
private: std::atomic_flag mF = ATOMIC_FLAG_INIT;
};
using mutex = spin_mutex;
using null_mutex = spin_mutex;

enum ets_key_usage_type { ets_key_per_instance, ets_no_key };

template<typename T, typename Alloc = std::allocator<T>, ets_key_usage_type K = ets_no_key>
class enumerable_thread_specific {
    using ListT = std::list<T>;
public:
    using iterator = typename ListT::iterator; using const_iterator = typename ListT::const_iterator;
    using reference = T&; using value_type = T; using size_type = std::size_t;
    using range_type = blocked_range<iterator>; using const_range_type = blocked_range<const_iterator>;
    enumerable_thread_specific() : mMake([]() { return std::unique_ptr<T>(new T()); }) {}
    template<typename U = T, typename = typename std::enable_if<std::is_copy_constructible<U>::value>::type>
    explicit enumerable_thread_specific(const T& exemplar)
        : mMake([exemplar]() { return std::unique_ptr<T>(new T(exemplar)); }) {}
    template<typename F, typename = decltype(std::declval<F>()()),
             typename = typename std::enable_if<!std::is_same<typename std::decay<F>::type, T>::value>::type>
    explicit enumerable_thread_specific(F f) : mMake([f]() { return std::unique_ptr<T>(new T(f())); }) {}
    T& local() { bool e; return local(e); }
    T& local(bool& exists) {
        std::lock_guard<std::mutex> g(mMx);
        auto id = std::this_thread::get_id();
        auto it = mIdx.find(id);
        exists = it != mIdx.end();
        if (exists) return *it->second;
        auto p = mMake();
        mItems.emplace_back(std::move(*p));
        T* q = &mItems.back();
        mIdx[id] = q;
        return *q;
    }
    iterator begin() { return mItems.begin(); } iterator end() { return mItems.end(); }
    const_iterator begin() const { return mItems.begin(); } const_iterator end() const { return mItems.end(); }
    size_type size() const { return mItems.size(); }
    bool empty() const { return mItems.empty(); }
    void clear() { mItems.clear(); mIdx.clear(); }
    range_type range(std::size_t g = 1) { return range_type(begin(), end(), g); }
    const_range_type range(std::size_t g = 1) const { return const_range_type(begin(), end(), g); }
    template<typename F> T combine(F f) {
        if (mItems.empty()) return *mMake();
        auto it = mItems.begin(); T r(*it);
        for (++it; it != mItems.end(); ++it) r = f(r, *it);
        return r;
    }
    template<typename F> void combine_each(F f) { for (auto& v : mItems) f(v); }
private:
    std::function<std::unique_ptr<T>()> mMake;
    ListT mItems; std::map<std::thread::id, T*> mIdx; std::mutex mMx;
};

template<typename T>
class combinable {
public:
    combinable() : mMake([]() { return T(); }) {}
    template<typename F> explicit combinable(F f) : mMake(f) {}
    T& local() { bool e; return local(e); }
    T& local(bool& exists) {
        std::lock_guard<std::mutex> g(mMx);
        auto id = std::this_thread::get_id(); auto it = mIdx.find(id);
        exists = it != mIdx.end();
        if (exists) return *it->second;
        mItems.emplace_back(mMake()); mIdx[id] = &mItems.back(); return mItems.back();
    }
    void clear() { mItems.clear(); mIdx.clear(); }
    template<typename F> T combine(F f) {
        if (mItems.empty()) return mMake();
        auto it = mItems.begin(); T r(*it);
        for (++it; it != mItems.end(); ++it) r = f(r, *it);
        return r;
    }
    template<typename F> void combine_each(F f) { for (auto& v : mItems) f(v); }
private:
    std::function<T()> mMake; std::list<T> mItems; std::map<std::thread::id, T*> mIdx; std::mutex mMx;
};

template<typename T, typename A = std::allocator<T>>
class concurrent_vector : public std::deque<T> {   // deque: push_back keeps references valid
public:
    using std::deque<T>::deque;
    typename std::deque<T>::iterator push_back(const T& v) { std::deque<T>::push_back(v); return std::prev(this->end()); }
    typename std::deque<T>::iterator push_back(T&& v) { std::deque<T>::push_back(std::move(v)); return std::prev(this->end()); }
    typename std::deque<T>::iterator grow_by(std::size_t n) { auto s = this->size(); this->resize(s + n); return this->begin() + s; }
};

template<typename K> struct tbb_hash_compare {
    static std::size_t hash(const K& k) { return std::hash<K>()(k); }
    static bool equal(const K& a, const K& b) { return a == b; }
};

template<typename K, typename V, typename HC = tbb_hash_compare<K>>
class concurrent_hash_map {
    struct H { std::size_t operator()(const K& k) const { return HC().hash(k); } };
    struct E { bool operator()(const K& a, const K& b) const { return HC().equal(a, b); } };
    using MapT = std::unordered_map<K, V, H, E>;
public:
    using value_type = typename MapT::value_type; using iterator = typename MapT::iterator;
    using const_iterator = typename MapT::const_iterator; using key_type = K; using mapped_type = V;
    class const_accessor {
    public:
        const value_type& operator*() const { return *mP; } const value_type* operator->() const { return mP; }
        bool empty() const { return !mP; } void release() { mP = nullptr; }
    protected: friend class concurrent_hash_map; value_type* mP = nullptr;
    };
    class accessor : public const_accessor {
    public:
        value_type& operator*() const { return *this->mP; } value_type* operator->() const { return this->mP; }
    };
    bool find(const_accessor& a, const K& k) const { auto it = const_cast<MapT&>(mM).find(k); if (it == mM.end()) { a.mP = nullptr; return false; } a.mP = &*it; return true; }
    bool find(accessor& a, const K& k) { auto it = mM.find(k); if (it == mM.end()) { a.mP = nullptr; return false; } a.mP = &*it; return true; }
    bool insert(const_accessor& a, const K& k) { auto r = mM.emplace(k, V()); a.mP = &*r.first; return r.second; }
    bool insert(accessor& a, const K& k) { auto r = mM.emplace(k, V()); a.mP = &*r.first; return r.second; }
    bool insert(const value_type& v) { return mM.insert(v).second; }
    bool insert(accessor& a, const value_type& v) { auto r = mM.insert(v); a.mP = &*r.first; return r.second; }
    bool erase(const K& k) { return mM.erase(k) > 0; }
    bool erase(const_accessor& a) { if (!a.mP) return false; K k = a.mP->first; a.mP = nullptr; return mM.erase(k) > 0; }
    bool erase(accessor& a) { if (!a.mP) return false; K k = a.mP->first; a.mP = nullptr; return mM.erase(k) > 0; }
    std::size_t size() const { return mM.size(); } bool empty() const { return mM.empty(); } void clear() { mM.clear(); }
    std::size_t count(const K& k) const { return mM.count(k); }
    iterator begin() { return mM.begin(); } iterator end() { return mM.end(); }
    const_iterator begin() const { return mM.begin(); } const_iterator end() const { return mM.end(); }
private: MapT mM;
};

class task_group_context {
public:
    bool cancel_group_execution() { bool w = mC.exchange(true); return !w; }
    bool is_group_execution_cancelled() const { return mC.load(); }
    void reset() { mC = false; }
private: std::atomic<bool> mC{false};
};
namespace task {
inline task_group_context* current_context() { static thread_local task_group_context ctx; ctx.reset(); return &ctx; }
}
enum task_group_status { not_complete, complete, canceled };
class task_group {
public:
    template<typename F> void run(const F& f) { f(); }
    template<typename F> task_group_status run_and_wait(const F& f) { f(); return complete; }
    task_group_status wait() { return complete; }
    void cancel() {}
    bool is_canceling() { return false; }
};
class task_arena {
public:
    struct attach {};
    static const int automatic = -1;
    task_arena(int = automatic, unsigned = 1) {}
    explicit task_arena(attach) {}
    template<typename F> auto execute(F&& f) -> decltype(f()) { return f(); }
    void initialize() {} void initialize(int, unsigned = 1) {}
    int max_concurrency() const { return tbb_shim::num_threads(); }
};
namespace this_task_arena {
inline int max_concurrency() { return tbb_shim::num_threads(); }
template<typename F> auto isolate(F&& f) -> decltype(f()) { return f(); }
}
class global_control {
public:
    enum parameter { max_allowed_parallelism, thread_stack_size };
    global_control(parameter p, std::size_t v) : mP(p), mOld(tbb_shim::num_threads()) { if (p == max_allowed_parallelism) tbb_shim::set_num_threads(int(v)); }
    ~global_control() { if (mP == max_allowed_parallelism) tbb_shim::set_num_threads(mOld); }
    static std::size_t active_value(parameter) { return std::size_t(tbb_shim::num_threads()); }
private: parameter mP; int mOld;
};
class tick_count {
    using clk = std::chrono::steady_clock;
public:
    class interval_t {
    public:
        interval_t() : mD(0) {} explicit interval_t(double s) : mD(s) {}
        double seconds() const { return mD; }
        interval_t operator+(const interval_t& o) const { return interval_t(mD + o.mD); }
        interval_t& operator+=(const interval_t& o) { mD += o.mD; return *this; }
    private: double mD;
    };
    static tick_count now() { tick_count t; t.mT = clk::now(); return t; }
    friend interval_t operator-(const tick_count& a, const tick_count& b) { return interval_t(std::chrono::duration<double>(a.mT - b.mT).count()); }
private: clk::time_point mT;
};
} // namespace tbb
