// Minimal host-only SYCL 2020 subset -- TEST INFRASTRUCTURE, written for this repository (not reference code).
//
// Purpose: the reference (TimThuering/N-Body-Simulation) includes <sycl/sycl.hpp> in every hot-path translation
// unit and no SYCL compiler exists in this image.  This header implements exactly the part of the SYCL API the
// reference uses, on the host, so that the reference's OWN, UNMODIFIED sources can be compiled with g++ where they lie
// under /root/reference (recipe: oracle/Makefile, target _ref) and serve as the parity oracle and the CPU baseline.
//
// Execution model (the same a library-only SYCL CPU back end uses):
//   * queue::submit runs the command group synchronously; buffers alias the host memory they were built from
//     (SYCL write-back semantics degenerate to "already there"); accessors are plain pointers.
//   * parallel_for(range)    -> OpenMP loop over work-items.
//   * parallel_for(nd_range) -> OpenMP loop over work-groups; the work-items of a group run one after another on
//     the group's thread.  A kernel that owns local memory may call nd_item::barrier(): then each work-item of the
//     group runs on its own fiber and barrier() switches to the next one (round robin), which is a correct
//     implementation of a work-group barrier for well-formed kernels.
//   * atomic_ref maps to the GCC __atomic builtins, so kernels that synchronise through global memory between
//     work-GROUPS are safe.  Work-items of one group never run concurrently, so a kernel in which a work-item
//     spin-waits for ANOTHER work-item of its own group to make progress would hang -- the reference's lock-based
//     octree insertion (locks are taken and released inside one insertion) and its CPU centre-of-mass pass (children
//     have larger node ids than their parents and are visited first) do not.
//   * device::is_gpu() is false: the reference then picks its own CPU code paths (computeCenterOfMass_CPU).
//   * sycl::rsqrt(x) = 1.0 / std::sqrt(x), what the host back ends of AdaptiveCpp and DPC++ both compute.
#pragma once

// (the real <sycl/sycl.hpp> pulls in most of the standard library; the reference relies on that for <iostream>,
// <chrono> and <stdexcept>)
#include <algorithm>
#include <array>
#include <atomic>
#include <chrono>
#include <iostream>
#include <map>
#include <sstream>
#include <stdexcept>
#include <cstring>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <limits>
#include <memory>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

#if defined(_OPENMP)
#include <omp.h>
#endif

namespace sycl {

// ---------------------------------------------------------------------------------------------------------------------
// index space
template<int D = 1>
struct range {
    static_assert(D == 1, "host shim: 1-dimensional index spaces only");
    std::size_t v;
    range(std::size_t n = 0) : v(n) {}
    std::size_t get(int) const { return v; }
    std::size_t operator[](int) const { return v; }
    std::size_t size() const { return v; }
};

template<int D>
struct item;

template<int D = 1>
struct id {
    static_assert(D == 1, "host shim: 1-dimensional index spaces only");
    std::size_t v;
    id(std::size_t n = 0) : v(n) {}
    id(const item<D> &it);
    std::size_t get(int) const { return v; }
    std::size_t operator[](int) const { return v; }
    operator std::size_t() const { return v; }
};

template<int D = 1>
struct item {
    std::size_t v, n;
    std::size_t get_id(int = 0) const { return v; }
    std::size_t get_linear_id() const { return v; }
    std::size_t get(int) const { return v; }
    std::size_t operator[](int) const { return v; }
    range<D> get_range() const { return range<D>(n); }
    operator std::size_t() const { return v; }
};

template<int D>
inline id<D>::id(const item<D> &it) : v(it.v) {}

template<int D = 1>
struct nd_range {
    range<D> global, local;
    nd_range(range<D> g, range<D> l) : global(g), local(l) {}
    range<D> get_global_range() const { return global; }
    range<D> get_local_range() const { return local; }
};

namespace access {
enum class fence_space { local_space, global_space, global_and_local };
enum class address_space { global_space, local_space, private_space, generic_space };
}  // namespace access

enum class memory_order { relaxed, acquire, release, acq_rel, seq_cst };
enum class memory_scope { work_item, sub_group, work_group, device, system };

namespace detail {

inline int to_builtin(memory_order o, bool is_load, bool is_store) {
    switch (o) {
        case memory_order::relaxed: return __ATOMIC_RELAXED;
        case memory_order::acquire: return is_store ? __ATOMIC_SEQ_CST : __ATOMIC_ACQUIRE;
        case memory_order::release: return is_load ? __ATOMIC_SEQ_CST : __ATOMIC_RELEASE;
        case memory_order::acq_rel: return is_load ? __ATOMIC_ACQUIRE : (is_store ? __ATOMIC_RELEASE : __ATOMIC_ACQ_REL);
        default: return __ATOMIC_SEQ_CST;
    }
}

// ---- fibers: one per work-item of a group whose kernel may call barrier() -------------------------------------------
#if defined(__x86_64__)
extern "C" void nbshim_fiber_switch(void **save_sp, void *load_sp);
// callee-saved registers of the SysV x86-64 ABI; weak so that every translation unit may carry a copy
__asm__(
    ".text\n"
    ".weak nbshim_fiber_switch\n"
    ".type nbshim_fiber_switch,@function\n"
    "nbshim_fiber_switch:\n"
    "  pushq %rbp\n  pushq %rbx\n  pushq %r12\n  pushq %r13\n  pushq %r14\n  pushq %r15\n"
    "  movq %rsp, (%rdi)\n"
    "  movq %rsi, %rsp\n"
    "  popq %r15\n  popq %r14\n  popq %r13\n  popq %r12\n  popq %rbx\n  popq %rbp\n"
    "  ret\n"
    ".size nbshim_fiber_switch,.-nbshim_fiber_switch\n");
#else
#error "host SYCL shim: the fiber switch is written for x86-64"
#endif

// local memory arena of the work-group running on this thread (plain pointer: one %fs-relative load per access)
inline thread_local char *tls_local_base = nullptr;

struct group_exec {
    // state of the work-group currently running on this thread
    bool fibers = false;                 // work-items run on fibers (barrier allowed)
    std::size_t local_size = 1;
    std::size_t current = 0;             // local id of the running work-item
    void *sched_sp = nullptr;
    std::vector<void *> fiber_sp;
    std::vector<char> done;
    std::function<void(std::size_t)> *body = nullptr;
    std::vector<char *> stacks;
    static constexpr std::size_t stack_bytes = 64 * 1024;

    ~group_exec() {
        for (char *s: stacks) std::free(s);
    }
};

inline group_exec &tls_group() {
    static thread_local group_exec g;
    return g;
}

inline void fiber_entry() {
    group_exec &g = tls_group();
    std::size_t me = g.current;
    (*g.body)(me);
    g.done[me] = 1;
    nbshim_fiber_switch(&g.fiber_sp[me], g.sched_sp);
    std::abort();  // a finished fiber is never resumed
}

inline void run_group_on_fibers(group_exec &g, std::size_t local_size, std::function<void(std::size_t)> &body) {
    g.fibers = true;
    g.local_size = local_size;
    g.body = &body;
    g.fiber_sp.assign(local_size, nullptr);
    g.done.assign(local_size, 0);
    while (g.stacks.size() < local_size) {
        void *p = nullptr;
        if (posix_memalign(&p, 64, group_exec::stack_bytes) != 0) std::abort();
        g.stacks.push_back(static_cast<char *>(p));
    }
    for (std::size_t l = 0; l < local_size; ++l) {
        // initial frame: six callee-saved register slots, then the entry address the first switch "returns" to
        std::uintptr_t top = reinterpret_cast<std::uintptr_t>(g.stacks[l] + group_exec::stack_bytes);
        top &= ~static_cast<std::uintptr_t>(15);
        void **sp = reinterpret_cast<void **>(top - 16);  // 16-byte aligned slot for the entry address
        sp[0] = reinterpret_cast<void *>(&fiber_entry);
        sp[1] = nullptr;
        for (int k = 1; k <= 6; ++k) sp[-k] = nullptr;
        g.fiber_sp[l] = sp - 6;
    }
    std::size_t remaining = local_size;
    while (remaining) {
        for (std::size_t l = 0; l < local_size; ++l) {
            if (g.done[l]) continue;
            g.current = l;
            nbshim_fiber_switch(&g.sched_sp, g.fiber_sp[l]);
            if (g.done[l]) --remaining;
        }
    }
    g.fibers = false;
    g.body = nullptr;
}

inline void group_barrier() {
    group_exec &g = tls_group();
    if (g.fibers) {
        std::size_t me = g.current;
        nbshim_fiber_switch(&g.fiber_sp[me], g.sched_sp);
        return;
    }
    if (g.local_size > 1) {
        std::fprintf(stderr, "host SYCL shim: nd_item::barrier() in a kernel without local memory is not supported\n");
        std::abort();
    }
}

}  // namespace detail

// ---------------------------------------------------------------------------------------------------------------------
template<int D = 1>
struct nd_item {
    std::size_t global_id, local_id, group_id, global_range, local_range;
    std::size_t get_global_id(int = 0) const { return global_id; }
    std::size_t get_global_linear_id() const { return global_id; }
    std::size_t get_local_id(int = 0) const { return local_id; }
    std::size_t get_local_linear_id() const { return local_id; }
    std::size_t get_group(int = 0) const { return group_id; }
    std::size_t get_group_linear_id() const { return group_id; }
    std::size_t get_group_range(int = 0) const { return global_range / local_range; }
    std::size_t get_global_range(int = 0) const { return global_range; }
    std::size_t get_local_range(int = 0) const { return local_range; }
    void barrier(access::fence_space = access::fence_space::global_and_local) const {
        __atomic_thread_fence(__ATOMIC_SEQ_CST);
        detail::group_barrier();
    }
    void mem_fence(access::fence_space = access::fence_space::global_and_local) const {
        __atomic_thread_fence(__ATOMIC_SEQ_CST);
    }
};

inline void atomic_fence(memory_order, memory_scope) { __atomic_thread_fence(__ATOMIC_SEQ_CST); }

// ---------------------------------------------------------------------------------------------------------------------
template<class T, memory_order DefaultOrder = memory_order::seq_cst, memory_scope DefaultScope = memory_scope::device,
        access::address_space Space = access::address_space::generic_space>
class atomic_ref {
    T *p;
public:
    explicit atomic_ref(T &ref) : p(&ref) {}
    T load(memory_order o = DefaultOrder, memory_scope = DefaultScope) const {
        return __atomic_load_n(p, detail::to_builtin(o, true, false));
    }
    void store(T v, memory_order o = DefaultOrder, memory_scope = DefaultScope) const {
        __atomic_store_n(p, v, detail::to_builtin(o, false, true));
    }
    T exchange(T v, memory_order o = DefaultOrder, memory_scope = DefaultScope) const {
        return __atomic_exchange_n(p, v, detail::to_builtin(o, false, false));
    }
    bool compare_exchange_strong(T &expected, T desired, memory_order o = DefaultOrder, memory_scope = DefaultScope) const {
        return __atomic_compare_exchange_n(p, &expected, desired, false, detail::to_builtin(o, false, false),
                                           __ATOMIC_ACQUIRE);
    }
    bool compare_exchange_weak(T &expected, T desired, memory_order o = DefaultOrder, memory_scope s = DefaultScope) const {
        return compare_exchange_strong(expected, desired, o, s);
    }
    T fetch_add(T v, memory_order o = DefaultOrder, memory_scope = DefaultScope) const {
        return __atomic_fetch_add(p, v, detail::to_builtin(o, false, false));
    }
    T fetch_sub(T v, memory_order o = DefaultOrder, memory_scope = DefaultScope) const {
        return __atomic_fetch_sub(p, v, detail::to_builtin(o, false, false));
    }
    operator T() const { return load(); }
    T operator=(T v) const { store(v); return v; }
    T operator++(int) const { return fetch_add(1); }
    T operator+=(T v) const { return fetch_add(v) + v; }
};

// ---------------------------------------------------------------------------------------------------------------------
// devices and selectors
namespace info {
namespace device {
struct name { using return_type = std::string; };
struct max_compute_units { using return_type = unsigned; };
}  // namespace device
}  // namespace info

class device {
public:
    bool is_gpu() const { return false; }
    bool is_cpu() const { return true; }
    template<class Param>
    typename Param::return_type get_info() const {
        if constexpr (std::is_same_v<Param, info::device::name>) {
            return std::string("host CPU (g++/OpenMP through oracle/sycl_shim)");
        } else {
#if defined(_OPENMP)
            return static_cast<typename Param::return_type>(omp_get_max_threads());
#else
            return typename Param::return_type(1);
#endif
        }
    }
};

struct default_selector_t {};
struct gpu_selector_t {};
struct cpu_selector_t {};
inline constexpr default_selector_t default_selector_v{};
inline constexpr gpu_selector_t gpu_selector_v{};   // accepted; the only device is the host
inline constexpr cpu_selector_t cpu_selector_v{};

// ---------------------------------------------------------------------------------------------------------------------
// buffers and accessors
template<class T, int D = 1>
class buffer {
    static_assert(D == 1, "host shim: 1-dimensional buffers only");
    T *ptr;
    std::size_t count;
    std::shared_ptr<std::vector<T>> owned;
public:
    using value_type = T;
    buffer(T *host, range<D> r) : ptr(host), count(r.size()) {}
    buffer(const T *host, range<D> r) : ptr(const_cast<T *>(host)), count(r.size()) {}
    explicit buffer(range<D> r) : count(r.size()), owned(std::make_shared<std::vector<T>>(r.size())) { ptr = owned->data(); }
    // SYCL 2020 contiguous-container constructor: the buffer uses (and writes back to) the container's memory
    template<class Container, class = std::enable_if_t<std::is_same_v<typename Container::value_type, T> &&
                                                       std::is_pointer_v<decltype(std::declval<Container &>().data())>>>
    buffer(Container &c) : ptr(c.data()), count(c.size()) {}
    std::size_t size() const { return count; }
    range<D> get_range() const { return range<D>(count); }
    T *host_data() const { return ptr; }
};

class handler;

struct read_only_t {};
struct write_only_t {};
struct read_write_t {};
struct no_init_t {};
inline constexpr read_only_t read_only{};
inline constexpr write_only_t write_only{};
inline constexpr read_write_t read_write{};
inline constexpr no_init_t no_init{};

template<class T, int D = 1>
class accessor {
    T *ptr;
    std::size_t count;
public:
    using value_type = T;
    template<class... Tags>
    accessor(buffer<T, D> &b, handler &, Tags...) : ptr(b.host_data()), count(b.size()) {}
    T &operator[](std::size_t i) const { return ptr[i]; }
    std::size_t size() const { return count; }
    T *get_pointer() const { return ptr; }
};

template<class T, int D = 1>
class host_accessor {
    T *ptr;
    std::size_t count;
public:
    using value_type = T;
    template<class... Tags>
    host_accessor(buffer<T, D> &b, Tags...) : ptr(b.host_data()), count(b.size()) {}
    T &operator[](std::size_t i) const { return ptr[i]; }
    std::size_t size() const { return count; }
    T *get_pointer() const { return ptr; }
};

template<class T, int D, class... Tags>
accessor(buffer<T, D> &, handler &, Tags...) -> accessor<T, D>;
template<class T, int D, class... Tags>
host_accessor(buffer<T, D> &, Tags...) -> host_accessor<T, D>;

template<class T, int D = 1>
class local_accessor {
    std::size_t offset, count;
public:
    local_accessor(range<D> r, handler &h);
    T &operator[](std::size_t i) const { return reinterpret_cast<T *>(detail::tls_local_base + offset)[i]; }
    std::size_t size() const { return count; }
};

// ---------------------------------------------------------------------------------------------------------------------
class handler {
    std::size_t local_bytes = 0;
    template<class T, int D> friend class local_accessor;

    std::size_t reserve_local(std::size_t bytes) {
        std::size_t off = (local_bytes + 63) & ~std::size_t(63);
        local_bytes = off + bytes;
        return off;
    }

public:
    template<class K>
    void single_task(const K &kernel) { kernel(); }

    template<class K>
    void parallel_for(range<1> r, const K &kernel) {
        const long long n = static_cast<long long>(r.size());
#pragma omp parallel for schedule(dynamic, 64) if (n > 256)
        for (long long i = 0; i < n; ++i) {
            if constexpr (std::is_invocable_v<const K &, id<1>>) {
                kernel(id<1>(static_cast<std::size_t>(i)));
            } else {
                item<1> it{static_cast<std::size_t>(i), static_cast<std::size_t>(n)};
                kernel(it);
            }
        }
    }

    template<class K>
    void parallel_for(nd_range<1> r, const K &kernel) {
        const std::size_t global = r.global.size(), local = r.local.size();
        if (local == 0 || global % local != 0) {
            std::fprintf(stderr, "host SYCL shim: the global range must be a multiple of the local range\n");
            std::abort();
        }
        const long long groups = static_cast<long long>(global / local);
        const std::size_t lbytes = local_bytes;
        const bool use_fibers = lbytes > 0 && local > 1;   // only kernels with local memory synchronise with barriers
#pragma omp parallel if (groups > 1)
        {
            std::vector<char> arena(lbytes + 64);
            detail::group_exec &g = detail::tls_group();
#pragma omp for schedule(dynamic, 1)
            for (long long grp = 0; grp < groups; ++grp) {
                detail::tls_local_base = reinterpret_cast<char *>((reinterpret_cast<std::uintptr_t>(arena.data()) + 63) & ~std::uintptr_t(63));
                g.local_size = local;
                if (use_fibers) {
                    std::function<void(std::size_t)> body = [&](std::size_t l) {
                        nd_item<1> it{static_cast<std::size_t>(grp) * local + l, l, static_cast<std::size_t>(grp), global, local};
                        kernel(it);
                    };
                    detail::run_group_on_fibers(g, local, body);
                } else {
                    for (std::size_t l = 0; l < local; ++l) {
                        nd_item<1> it{static_cast<std::size_t>(grp) * local + l, l, static_cast<std::size_t>(grp), global, local};
                        kernel(it);
                    }
                }
            }
            detail::tls_local_base = nullptr;
            g.local_size = 1;
        }
    }
};

template<class T, int D>
inline local_accessor<T, D>::local_accessor(range<D> r, handler &h) : offset(h.reserve_local(r.size() * sizeof(T))), count(r.size()) {}

class event {
public:
    void wait() const {}
    void wait_and_throw() const {}
};

class queue {
    device dev;
public:
    queue() = default;
    template<class Selector>
    explicit queue(const Selector &) {}
    template<class F>
    event submit(F &&command_group) {
        handler h;
        command_group(h);
        return event();
    }
    void wait() const {}
    void wait_and_throw() const {}
    device get_device() const { return dev; }
};

// ---------------------------------------------------------------------------------------------------------------------
// math built-ins used by the reference (host definitions)
inline double rsqrt(double x) { return 1.0 / std::sqrt(x); }
inline float rsqrt(float x) { return 1.0f / std::sqrt(x); }
inline double sqrt(double x) { return std::sqrt(x); }
inline float sqrt(float x) { return std::sqrt(x); }
inline double ceil(double x) { return std::ceil(x); }
inline double floor(double x) { return std::floor(x); }
inline double fabs(double x) { return std::fabs(x); }
template<class T> inline T min(T a, T b) { return b < a ? b : a; }
template<class T> inline T max(T a, T b) { return a < b ? b : a; }

}  // namespace sycl
