// Just enough of the CUDA execution model to run libbsi's kernel template on the host (test
// infrastructure): one block of BSI_BLOCK threads = BSI_BLOCK / 32 warps of 32 host threads each (the
// warps run independently, as on the device, so races between warps can show); warp collectives are
// rendezvous on the warp's barrier, atomics are the compiler's, shared memory is static storage.
// Every collective in the template is issued with the full mask at a warp-uniform point, so "all 32
// threads meet at the barrier in the same order" is exactly the device's behaviour.
#pragma once
#include <pthread.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#define __device__
#define __global__
#define __host__
#define __forceinline__ inline
#define __noinline__ __attribute__((noinline))
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(x)

struct bsi_dim3 { unsigned x, y, z; };
static thread_local bsi_dim3 threadIdx = {0, 0, 0};
static bsi_dim3 blockIdx = {0, 0, 0}, blockDim = {32, 1, 1}, gridDim = {1, 1, 1};

static inline unsigned bsi_host_lop3(int lut, unsigned a, unsigned b, unsigned c) {
    unsigned r = 0;
    for (int i = 0; i < 8; ++i)
        if ((lut >> i) & 1) r |= ((i & 4) ? a : ~a) & ((i & 2) ? b : ~b) & ((i & 1) ? c : ~c);
    return r;
}

static inline int __ffs(unsigned x) { return __builtin_ffs((int)x); }
static inline int __popc(unsigned x) { return __builtin_popcount(x); }

// ---- the warp ------------------------------------------------------------------------------------
#define BSI_EMUL_MAX_WARPS 8
static pthread_barrier_t bsi_warp_barriers[BSI_EMUL_MAX_WARPS], bsi_block_barrier;
static unsigned long long bsi_warp_slots[BSI_EMUL_MAX_WARPS][32];
#define bsi_warp_slot bsi_warp_slots[threadIdx.x >> 5]
static inline void bsi_rendezvous() { pthread_barrier_wait(&bsi_warp_barriers[threadIdx.x >> 5]); }
static inline void __syncwarp(unsigned = 0xFFFFFFFFu) { bsi_rendezvous(); }
static inline void __syncthreads() { pthread_barrier_wait(&bsi_block_barrier); }

template <typename T>
static inline T bsi_exchange(T v, int src) {
    const unsigned lane = threadIdx.x & 31u;
    unsigned long long w = 0;
    memcpy(&w, &v, sizeof(T));
    bsi_warp_slot[lane] = w;
    bsi_rendezvous();
    const unsigned long long got = bsi_warp_slot[src & 31];
    bsi_rendezvous();
    T r;
    memcpy(&r, &got, sizeof(T));
    return r;
}
template <typename T> static inline T __shfl_sync(unsigned, T v, int src) { return bsi_exchange(v, src); }
template <typename T> static inline T __shfl_xor_sync(unsigned, T v, int d) {
    return bsi_exchange(v, (int)((threadIdx.x & 31u) ^ (unsigned)d));
}
template <typename T> static inline T __shfl_up_sync(unsigned, T v, int d) {
    const int lane = (int)(threadIdx.x & 31u);
    const T got = bsi_exchange(v, lane >= d ? lane - d : lane);      // (every thread takes part)
    return lane >= d ? got : v;
}
static inline unsigned __ballot_sync(unsigned, bool pred) {
    const unsigned lane = threadIdx.x & 31u;
    bsi_warp_slot[lane] = pred ? 1ull : 0ull;
    bsi_rendezvous();
    unsigned r = 0;
    for (int l = 0; l < 32; ++l) r |= (unsigned)bsi_warp_slot[l] << l;
    bsi_rendezvous();
    return r;
}
static inline bool __all_sync(unsigned m, bool pred) { return __ballot_sync(m, pred) == 0xFFFFFFFFu; }
static inline bool __any_sync(unsigned m, bool pred) { return __ballot_sync(m, pred) != 0u; }
static inline unsigned __reduce_add_sync(unsigned, unsigned v) {
    const unsigned lane = threadIdx.x & 31u;
    bsi_warp_slot[lane] = v;
    bsi_rendezvous();
    unsigned r = 0;
    for (int l = 0; l < 32; ++l) r += (unsigned)bsi_warp_slot[l];
    bsi_rendezvous();
    return r;
}
static inline unsigned __reduce_max_sync(unsigned, unsigned v) {
    const unsigned lane = threadIdx.x & 31u;
    bsi_warp_slot[lane] = v;
    bsi_rendezvous();
    unsigned r = 0;
    for (int l = 0; l < 32; ++l) r = r > (unsigned)bsi_warp_slot[l] ? r : (unsigned)bsi_warp_slot[l];
    bsi_rendezvous();
    return r;
}

// one block of n_threads threads (a multiple of 32, at most 32 * BSI_EMUL_MAX_WARPS) runs `kernel`
#include <functional>
static inline void bsi_emul_block(unsigned n_threads, const std::function<void()> &kernel) {
    const unsigned n_warps = n_threads / 32;
    if (n_threads % 32 || n_warps == 0 || n_warps > BSI_EMUL_MAX_WARPS) { fprintf(stderr, "bad block size\n"); exit(9); }
    blockDim.x = n_threads;
    pthread_barrier_init(&bsi_block_barrier, nullptr, n_threads);
    for (unsigned w = 0; w < n_warps; ++w) pthread_barrier_init(&bsi_warp_barriers[w], nullptr, 32);
    struct Arg { const std::function<void()> *k; unsigned t; };
    Arg *args = new Arg[n_threads];
    pthread_t *th = new pthread_t[n_threads];
    for (unsigned t = 0; t < n_threads; ++t) {
        args[t] = {&kernel, t};
        pthread_create(&th[t], nullptr, [](void *p) -> void * {
            Arg *a = (Arg *)p;
            threadIdx.x = a->t;
            (*a->k)();
            return nullptr;
        }, &args[t]);
    }
    for (unsigned t = 0; t < n_threads; ++t) pthread_join(th[t], nullptr);
    for (unsigned w = 0; w < n_warps; ++w) pthread_barrier_destroy(&bsi_warp_barriers[w]);
    pthread_barrier_destroy(&bsi_block_barrier);
    delete[] args;
    delete[] th;
}

// ---- atomics / fences ----------------------------------------------------------------------------
static inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) {
    return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST);
}
static inline unsigned atomicAdd(unsigned *p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned atomicCAS(unsigned *p, unsigned expected, unsigned desired) {
    __atomic_compare_exchange_n(p, &expected, desired, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST);
    return expected;
}
static inline unsigned atomicExch(unsigned *p, unsigned v) { return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned long long atomicExch(unsigned long long *p, unsigned long long v) {
    return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST);
}
static inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }

// ---- device math the template uses ------------------------------------------------------------------
static inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
static inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
static inline unsigned long long max(unsigned long long a, unsigned long long b) { return a > b ? a : b; }
static inline unsigned long long min(unsigned long long a, unsigned long long b) { return a < b ? a : b; }
