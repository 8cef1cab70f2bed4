// Just enough of the CUDA execution model to run libbsi's kernel template on the host (test
// infrastructure): one block of 32 threads = one warp = 32 host threads; warp collectives are
// rendezvous on a barrier, atomics are the compiler's, shared memory is static storage.
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
static pthread_barrier_t bsi_warp_barrier;
static unsigned long long bsi_warp_slot[32];
static inline void bsi_rendezvous() { pthread_barrier_wait(&bsi_warp_barrier); }
static inline void __syncwarp(unsigned = 0xFFFFFFFFu) { bsi_rendezvous(); }
static inline void __syncthreads() { bsi_rendezvous(); }

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
