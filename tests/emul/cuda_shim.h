// TEST INFRASTRUCTURE ONLY -- a minimal single-host-thread SIMT emulator.
//
// The build container has no GPU, so the kernel sources under
// gpu4pyscf_b200/csrc are also compiled with g++ -DCPU_EMUL against this shim:
// every CUDA thread of a block becomes a ucontext fiber, __syncthreads() yields
// to the next fiber, blocks run one after another, atomics are plain adds.
// It exists to debug index arithmetic and recurrences before spending GPU
// minutes; it is never loaded by the product path (gpu4pyscf_b200/lib/loader.py
// only opens it when GPU4PYSCF_B200_EMUL=1 is set by the CPU test-suite).
#pragma once
#include <ucontext.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <stdint.h>
#include <math.h>
#include <vector>
#include <functional>
#include <algorithm>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __restrict__
#define __launch_bounds__(...)
#define __constant__ static
#define EMUL_SHARED static

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
struct uint3_ { unsigned x, y, z; };
struct double2 { double x, y; };
struct float4 { float x, y, z, w; };

extern uint3_ threadIdx, blockIdx;
extern dim3 blockDim, gridDim;
extern char *emul_dynamic_smem;

typedef int cudaError_t;
typedef void *cudaStream_t;
#define cudaSuccess 0

void emul_yield();
void emul_run_block(std::function<void()> body, unsigned nthreads);

// barriers: id 0 = whole block, 1..15 = named (team) barriers, 64+w = warp w
void emul_barrier(int id, int nthreads);
extern double emul_xchg[64][32];
static inline int emul_warp_threads()
{
    int nt = blockDim.x, w = threadIdx.x >> 5;
    return nt - w * 32 < 32 ? nt - w * 32 : 32;
}
static inline void __syncthreads() { emul_barrier(0, blockDim.x); }
static inline void __syncwarp(unsigned = 0xffffffff) { emul_barrier(64 + (threadIdx.x >> 5), emul_warp_threads()); }
template <class T> static inline T emul_shfl(T v, int src)
{
    static_assert(sizeof(T) <= 8, "shuffle of at most 8 bytes");
    int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (src < 0 || src >= emul_warp_threads()) src = lane;
    memcpy(&emul_xchg[w][lane], &v, sizeof(T));
    __syncwarp();
    T r;
    memcpy(&r, &emul_xchg[w][src], sizeof(T));
    __syncwarp();
    return r;
}
template <class T> static inline T __shfl_sync(unsigned, T v, int src) { return emul_shfl(v, src); }
template <class T> static inline T __shfl_down_sync(unsigned, T v, int d) { return emul_shfl(v, (int)(threadIdx.x & 31) + d); }
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int m) { return emul_shfl(v, (int)(threadIdx.x & 31) ^ m); }
static inline unsigned __ballot_sync(unsigned, bool pred)
{
    int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    long long p = pred ? 1 : 0;
    memcpy(&emul_xchg[w][lane], &p, 8);
    __syncwarp();
    unsigned m = 0;
    for (int i = 0; i < emul_warp_threads(); ++i) {
        long long q;
        memcpy(&q, &emul_xchg[w][i], 8);
        if (q) m |= 1u << i;
    }
    __syncwarp();
    return m;
}
static inline bool __any_sync(unsigned mask, bool pred) { return __ballot_sync(mask, pred) != 0; }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int min(int a, int b) { return a < b ? a : b; }
static inline double atomicAdd(double *p, double v) { double o = *p; *p = o + v; return o; }
static inline int atomicAdd(int *p, int v) { int o = *p; *p = o + v; return o; }
static inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { unsigned long long o = *p; *p = o + v; return o; }
static inline unsigned atomicAdd(unsigned *p, unsigned v) { unsigned o = *p; *p = o + v; return o; }
static inline int atomicMax(int *p, int v) { int o = *p; if (v > o) *p = v; return o; }
static inline double __ldg(const double *p) { return *p; }
static inline float __ldg(const float *p) { return *p; }
static inline int __ldg(const int *p) { return *p; }
static inline uint32_t __ldg(const uint32_t *p) { return *p; }
static inline double rsqrt(double x) { return 1. / sqrt(x); }

template <class K, class... Args>
void emul_launch(K kernel, dim3 grid, dim3 block, size_t smem, Args... args)
{
    gridDim = grid;
    blockDim = block;
    std::vector<char> buf(smem + 64);
    for (unsigned b = 0; b < grid.x; ++b) {
        blockIdx.x = b; blockIdx.y = 0; blockIdx.z = 0;
        memset(buf.data(), 0xff, buf.size());   // poison: NaN doubles expose uninitialised reads
        emul_dynamic_smem = buf.data();
        emul_run_block([&]() { kernel(args...); }, block.x * block.y * block.z);
    }
}
