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
#include <condition_variable>
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

// concurrent_vector: segmented storage (segment k holds 16 << k elements, elements never move), so that push_back /
// grow_by from several threads and element access by index or iterator can run at the same time, as with TBB's own.
// Growth is serialised by a mutex; the segment table is a fixed array of atomic pointers, reads take no lock.
template<typename T, typename A = std::allocator<T>>
class concurrent_vector {
    static constexpr int kSegs = 48, kLog0 = 4;
    static int segOf(std::size_t i) { int k = 0; std::size_t n = (i >> kLog0) + 1; while (n >>= 1) ++k; return k; }
    static std::size_t segBase(int k) { return ((std::size_t(1) << k) - 1) << kLog0; }
    static std::size_t segSize(int k) { return std::size_t(1) << (k + kLog0); }
    T* slot(std::size_t i) const { const int k = segOf(i); return mSeg[k].load(std::memory_order_acquire) + (i - segBase(k)); }
    template<typename VecT, typename RefT>
    class iter {
    public:
        using iterator_category = std::random_access_iterator_tag; using value_type = T; using difference_type = std::ptrdiff_t;
        using pointer = typename std::remove_reference<RefT>::type*; using reference = RefT;
        iter() = default; iter(VecT* v, std::size_t i) : mV(v), mI(i) {}
        template<typename V2, typename R2> iter(const iter<V2, R2>& o) : mV(o.mV), mI(o.mI) {}
        reference operator*() const { return *mV->slot(mI); } pointer operator->() const { return mV->slot(mI); }
        reference operator[](difference_type d) const { return *mV->slot(mI + d); }
        iter& operator++() { ++mI; return *this; } iter operator++(int) { iter t(*this); ++mI; return t; }
        iter& operator--() { --mI; return *this; } iter operator--(int) { iter t(*this); --mI; return t; }
        iter& operator+=(difference_type d) { mI += d; return *this; } iter& operator-=(difference_type d) { mI -= d; return *this; }
        friend iter operator+(iter a, difference_type d) { a.mI += d; return a; } friend iter operator+(difference_type d, iter a) { a.mI += d; return a; }
        friend iter operator-(iter a, difference_type d) { a.mI -= d; return a; }
        friend difference_type operator-(const iter& a, const iter& b) { return difference_type(a.mI) - difference_type(b.mI); }
        friend bool operator==(const iter& a, const iter& b) { return a.mI == b.mI; } friend bool operator!=(const iter& a, const iter& b) { return a.mI != b.mI; }
        friend bool operator<(const iter& a, const iter& b) { return a.mI < b.mI; } friend bool operator>(const iter& a, const iter& b) { return a.mI > b.mI; }
        friend bool operator<=(const iter& a, const iter& b) { return a.mI <= b.mI; } friend bool operator>=(const iter& a, const iter& b) { return a.mI >= b.mI; }
        VecT* mV = nullptr; std::size_t mI = 0;
    };
public:
    using value_type = T; using size_type = std::size_t; using reference = T&; using const_reference = const T&;
    using iterator = iter<concurrent_vector, T&>; using const_iterator = iter<const concurrent_vector, const T&>;
    concurrent_vector() { for (auto& s : mSeg) s.store(nullptr, std::memory_order_relaxed); }
    explicit concurrent_vector(size_type n, const T& v = T()) : concurrent_vector() { grow_to(n, &v); }
    concurrent_vector(const concurrent_vector& o) : concurrent_vector() { for (size_type i = 0; i < o.size(); ++i) push_back(o[i]); }
    concurrent_vector& operator=(const concurrent_vector& o) { if (this != &o) { clear(); for (size_type i = 0; i < o.size(); ++i) push_back(o[i]); } return *this; }
    ~concurrent_vector() { clear(); for (auto& s : mSeg) { ::operator delete(static_cast<void*>(s.load())); } }
    iterator push_back(const T& v) { std::lock_guard<std::mutex> g(mMx); const size_type i = mSize.load(std::memory_order_relaxed); reserve_locked(i + 1); new (slot(i)) T(v); mSize.store(i + 1, std::memory_order_release); return iterator(this, i); }
    iterator push_back(T&& v) { std::lock_guard<std::mutex> g(mMx); const size_type i = mSize.load(std::memory_order_relaxed); reserve_locked(i + 1); new (slot(i)) T(std::move(v)); mSize.store(i + 1, std::memory_order_release); return iterator(this, i); }
    template<typename... Args> iterator emplace_back(Args&&... a) { std::lock_guard<std::mutex> g(mMx); const size_type i = mSize.load(std::memory_order_relaxed); reserve_locked(i + 1); new (slot(i)) T(std::forward<Args>(a)...); mSize.store(i + 1, std::memory_order_release); return iterator(this, i); }
    iterator grow_by(size_type n) { std::lock_guard<std::mutex> g(mMx); const size_type s = mSize.load(std::memory_order_relaxed); reserve_locked(s + n); for (size_type i = s; i < s + n; ++i) new (slot(i)) T(); mSize.store(s + n, std::memory_order_release); return iterator(this, s); }
    iterator grow_by(size_type n, const T& v) { std::lock_guard<std::mutex> g(mMx); const size_type s = mSize.load(std::memory_order_relaxed); reserve_locked(s + n); for (size_type i = s; i < s + n; ++i) new (slot(i)) T(v); mSize.store(s + n, std::memory_order_release); return iterator(this, s); }
    void resize(size_type n) { grow_to(n, nullptr); }
    void reserve(size_type n) { std::lock_guard<std::mutex> g(mMx); reserve_locked(n); }
    void clear() { std::lock_guard<std::mutex> g(mMx); const size_type n = mSize.load(); for (size_type i = 0; i < n; ++i) slot(i)->~T(); mSize.store(0); }
    size_type size() const { return mSize.load(std::memory_order_acquire); }
    bool empty() const { return size() == 0; }
    reference operator[](size_type i) { return *slot(i); } const_reference operator[](size_type i) const { return *slot(i); }
    reference at(size_type i) { return *slot(i); } const_reference at(size_type i) const { return *slot(i); }
    reference front() { return *slot(0); } reference back() { return *slot(size() - 1); }
    const_reference front() const { return *slot(0); } const_reference back() const { return *slot(size() - 1); }
    iterator begin() { return iterator(this, 0); } iterator end() { return iterator(this, size()); }
    const_iterator begin() const { return const_iterator(this, 0); } const_iterator end() const { return const_iterator(this, size()); }
    const_iterator cbegin() const { return begin(); } const_iterator cend() const { return end(); }
private:
    void reserve_locked(size_type n) {
        if (n == 0) return;
        for (int k = 0, last = segOf(n - 1); k <= last; ++k)
            if (!mSeg[k].load(std::memory_order_relaxed)) mSeg[k].store(static_cast<T*>(::operator new(segSize(k) * sizeof(T))), std::memory_order_release);
    }
    void grow_to(size_type n, const T* v) {
        std::lock_guard<std::mutex> g(mMx);
        const size_type s = mSize.load(std::memory_order_relaxed);
        if (n <= s) { for (size_type i = n; i < s; ++i) slot(i)->~T(); mSize.store(n, std::memory_order_release); return; }
        reserve_locked(n);
        for (size_type i = s; i < n; ++i) { if (v) new (slot(i)) T(*v); else new (slot(i)) T(); }
        mSize.store(n, std::memory_order_release);
    }
    std::atomic<T*> mSeg[kSegs]; std::atomic<size_type> mSize{0}; mutable std::mutex mMx;
};

template<typename K> struct tbb_hash_compare {
    static std::size_t hash(const K& k) { return std::hash<K>()(k); }
    static bool equal(const K& a, const K& b) { return a == b; }
};

// concurrent_hash_map with TBB's locking contract: find / insert / erase may be called from any number of threads; an
// `accessor` holds a WRITE lock on its element and a `const_accessor` a READ lock until it is released or destroyed;
// erase(key) waits for the element's lock.  The table itself is guarded by one mutex (the registry of a tree's value
// accessors -- tree/Tree.h:1081-1082,1423-1450 -- sees one insert and one erase per task, so a single lock is enough).
// Iteration is not safe against concurrent modification, exactly as in TBB.
template<typename K, typename V, typename HC = tbb_hash_compare<K>>
class concurrent_hash_map {
    struct H { std::size_t operator()(const K& k) const { return HC().hash(k); } };
    struct E { bool operator()(const K& a, const K& b) const { return HC().equal(a, b); } };
public:
    using value_type = std::pair<const K, V>; using key_type = K; using mapped_type = V;
private:
    struct Node {
        value_type kv; std::mutex mx; std::condition_variable cv; int readers = 0; bool writer = false;
        explicit Node(const K& k) : kv(k, V()) {} explicit Node(const value_type& v) : kv(v) {}
        void lock(bool write) { std::unique_lock<std::mutex> l(mx); if (write) { cv.wait(l, [&] { return !writer && readers == 0; }); writer = true; } else { cv.wait(l, [&] { return !writer; }); ++readers; } }
        void unlock(bool write) { { std::lock_guard<std::mutex> l(mx); if (write) writer = false; else --readers; } cv.notify_all(); }
    };
    using NodeP = std::shared_ptr<Node>;
    using MapT = std::unordered_map<K, NodeP, H, E>;
    template<typename MapIt, typename RefT>
    class iter {
    public:
        using iterator_category = std::forward_iterator_tag; using value_type = typename concurrent_hash_map::value_type;
        using difference_type = std::ptrdiff_t; using pointer = typename std::remove_reference<RefT>::type*; using reference = RefT;
        iter() = default; explicit iter(MapIt it) : mIt(it) {}
        reference operator*() const { return mIt->second->kv; } pointer operator->() const { return &mIt->second->kv; }
        iter& operator++() { ++mIt; return *this; } iter operator++(int) { iter t(*this); ++mIt; return t; }
        friend bool operator==(const iter& a, const iter& b) { return a.mIt == b.mIt; } friend bool operator!=(const iter& a, const iter& b) { return a.mIt != b.mIt; }
        MapIt mIt;
    };
public:
    using iterator = iter<typename MapT::iterator, value_type&>; using const_iterator = iter<typename MapT::const_iterator, const value_type&>;
    class const_accessor {
    public:
        const_accessor() = default; const_accessor(const const_accessor&) = delete; const_accessor& operator=(const const_accessor&) = delete;
        ~const_accessor() { release(); }
        const value_type& operator*() const { return mN->kv; } const value_type* operator->() const { return &mN->kv; }
        bool empty() const { return !mN; }
        void release() { if (mN) { mN->unlock(mWrite); mN.reset(); } }
    protected:
        friend class concurrent_hash_map;
        void hold(const NodeP& n, bool write) { release(); n->lock(write); mN = n; mWrite = write; }
        NodeP mN; bool mWrite = false;
    };
    class accessor : public const_accessor {
    public:
        value_type& operator*() const { return this->mN->kv; } value_type* operator->() const { return &this->mN->kv; }
    };
    concurrent_hash_map() = default;
    bool find(const_accessor& a, const K& k) const { return lookup(a, k, false); }
    bool find(accessor& a, const K& k) { return lookup(a, k, true); }
    bool insert(const_accessor& a, const K& k) { return emplace(&a, false, k, nullptr); }
    bool insert(accessor& a, const K& k) { return emplace(&a, true, k, nullptr); }
    bool insert(const value_type& v) { return emplace(nullptr, false, v.first, &v); }
    bool insert(const_accessor& a, const value_type& v) { return emplace(&a, false, v.first, &v); }
    bool insert(accessor& a, const value_type& v) { return emplace(&a, true, v.first, &v); }
    bool erase(const K& k) {
        NodeP n;
        { std::lock_guard<std::mutex> g(mMx); auto it = mM.find(k); if (it == mM.end()) return false; n = it->second; mM.erase(it); }
        n->lock(true); n->unlock(true);       // wait for the holders of the element, as TBB does
        return true;
    }
    bool erase(const_accessor& a) { return eraseHeld(a); }
    bool erase(accessor& a) { return eraseHeld(a); }
    std::size_t size() const { std::lock_guard<std::mutex> g(mMx); return mM.size(); }
    bool empty() const { return size() == 0; }
    void clear() { std::lock_guard<std::mutex> g(mMx); mM.clear(); }
    std::size_t count(const K& k) const { std::lock_guard<std::mutex> g(mMx); return mM.count(k); }
    iterator begin() { return iterator(mM.begin()); } iterator end() { return iterator(mM.end()); }
    const_iterator begin() const { return const_iterator(mM.begin()); } const_iterator end() const { return const_iterator(mM.end()); }
private:
    bool lookup(const_accessor& a, const K& k, bool write) const {
        a.release();
        NodeP n;
        { std::lock_guard<std::mutex> g(mMx); auto it = mM.find(k); if (it == mM.end()) return false; n = it->second; }
        a.hold(n, write);
        return true;
    }
    bool emplace(const_accessor* a, bool write, const K& k, const value_type* v) {
        if (a) a->release();
        NodeP n; bool fresh;
        { std::lock_guard<std::mutex> g(mMx);
          auto it = mM.find(k);
          fresh = it == mM.end();
          if (fresh) { n = v ? std::make_shared<Node>(*v) : std::make_shared<Node>(k); mM.emplace(k, n); } else n = it->second; }
        if (a) a->hold(n, write);
        return fresh;
    }
    bool eraseHeld(const_accessor& a) {
        if (a.empty()) return false;
        const K k = a.mN->kv.first; NodeP n = a.mN; bool gone = false;
        { std::lock_guard<std::mutex> g(mMx); auto it = mM.find(k); if (it != mM.end() && it->second == n) { mM.erase(it); gone = true; } }
        a.release();
        return gone;
    }
    MapT mM; mutable std::mutex mMx;
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
