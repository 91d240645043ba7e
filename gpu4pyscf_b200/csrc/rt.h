// Thin runtime indirection so the same sources build for sm_100a (nvcc) and for
// the CPU SIMT emulator used by the CPU-only test-suite (tests/emul).
#pragma once
#include "common.cuh"

#ifdef CPU_EMUL
#define KLAUNCH(kernel, grid, block, smem, stream, ...) \
    emul_launch(kernel, dim3(grid), dim3(block), smem, __VA_ARGS__)
static inline int rt_last_error(const char *what) { return 0; }
static inline int rt_num_sm() { return 2; }
static inline int rt_smem_max() { return 227 * 1024; }
static inline int rt_device() { return 0; }
template <class K> static inline int rt_prepare(K k, int smem, int nthreads) { return 1; }
static inline void *rt_upload(const void *src, size_t n)
{
    void *p = malloc(n);
    memcpy(p, src, n);
    return p;
}
#else
#include <stdio.h>
#define KLAUNCH(kernel, grid, block, smem, stream, ...) \
    kernel<<<grid, block, smem, stream>>>(__VA_ARGS__)
static inline int rt_last_error(const char *what)
{
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) {
        fprintf(stderr, "CUDA error in %s: %s\n", what, cudaGetErrorString(err));
        return 1;
    }
    return 0;
}
static inline int rt_device() { int d = 0; cudaGetDevice(&d); return d; }
static inline int rt_num_sm()
{
    int n = 0;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, rt_device());
    return n;
}
static inline int rt_smem_max()
{
    int n = 0;
    cudaDeviceGetAttribute(&n, cudaDevAttrMaxSharedMemoryPerBlockOptin, rt_device());
    return n;
}
// opt in to the dynamic shared memory size; returns resident blocks per SM
template <class K> static inline int rt_prepare(K k, int smem, int nthreads)
{
    if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess)
        return 0;
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k, nthreads, smem) != cudaSuccess)
        return 0;
    return nb;
}
static inline void *rt_upload(const void *src, size_t n)
{
    void *p = nullptr;
    if (cudaMalloc(&p, n) != cudaSuccess) return nullptr;
    cudaMemcpy(p, src, n, cudaMemcpyHostToDevice);
    return p;
}
#endif
