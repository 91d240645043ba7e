// Rys quadrature nodes/weights from the mpmath-generated tables (rys_tables.h,
// tools/gen_rys_tables.py).  Plays the role of the reference's rys_roots()
// (gpu4pyscf/lib/gvhf-rys/rys_roots.cu:25-88) with a different decomposition:
// one call evaluates ONE of the 2*NROOTS values, so the 2*NROOTS Clenshaw
// recurrences of a primitive quartet run on 2*NROOTS different threads.
#pragma once
#include "common.cuh"

#define RYS_NINT_ 40
#define RYS_NCHEB_ 14

// idx <  NROOTS : node  u_idx = t_idx^2
// idx >= NROOTS : weight W_{idx-NROOTS}   (sum_i W_i = F_0(x))
template <int NROOTS>
__device__ __forceinline__
double rys_value(int idx, double x, const double *__restrict__ cheb, const double *__restrict__ large)
{
    constexpr double XL = (35 + 5 * NROOTS) < 100 ? (35. + 5 * NROOTS) : 100.;
    if (x >= XL) {
        double r = large[NROOTS * (NROOTS - 1) + idx];
        double xi = 1. / x;
        return idx < NROOTS ? r * xi : r * sqrt(xi);
    }
    int it = (int)(x * 0.4);
    double s = (x - it * 2.5) * 0.8 - 1.;
    const double *c = cheb + RYS_NINT_ * RYS_NCHEB_ * NROOTS * (NROOTS - 1)
                    + (it * 2 * NROOTS + idx) * RYS_NCHEB_;
    double s2 = s + s;
    double b1 = c[RYS_NCHEB_ - 1];
    double b2 = 0.;
#pragma unroll
    for (int k = RYS_NCHEB_ - 2; k > 0; --k) {
        double t = s2 * b1 - b2 + c[k];
        b2 = b1;
        b1 = t;
    }
    return s * b1 - b2 + c[0];
}

// ---- shared-memory copy of the NROOTS slice of the tables -----------------
// Random per-lane gathers of 14-coefficient rows are what the root evaluation
// costs; from global memory every distinct row is its own L1 wavefront (32 per
// warp-wide load), from shared memory a 112-byte row is seven 16-byte loads
// with few bank conflicts.  Only the intervals below the asymptotic switch are
// copied: rows [0, RysSmem<NROOTS>::NINT) of the interval-major table, followed
// by the 2*NROOTS large-x constants.
template <int NROOTS>
struct RysSmem {
    static constexpr int XL = (35 + 5 * NROOTS) < 100 ? (35 + 5 * NROOTS) : 100;
    static constexpr int NINT = (XL * 2 + 4) / 5;                 // ceil(XL / 2.5)
    static constexpr int CHEB_DOUBLES = NINT * 2 * NROOTS * RYS_NCHEB_;
    static constexpr int DOUBLES = (CHEB_DOUBLES + 2 * NROOTS + 1) & ~1;   // keep 16-byte alignment
};

template <int NROOTS>
__device__ __forceinline__
void rys_stage_tables(double *tab, const double *__restrict__ cheb, const double *__restrict__ large,
                      int tid, int nthreads)
{
    const double *src = cheb + RYS_NINT_ * RYS_NCHEB_ * NROOTS * (NROOTS - 1);
    for (int i = tid; i < RysSmem<NROOTS>::CHEB_DOUBLES; i += nthreads) tab[i] = src[i];
    for (int i = tid; i < 2 * NROOTS; i += nthreads)
        tab[RysSmem<NROOTS>::CHEB_DOUBLES + i] = large[NROOTS * (NROOTS - 1) + i];
}

// same contract as rys_value, tables in shared memory (16-byte aligned)
template <int NROOTS>
__device__ __forceinline__
double rys_value_s(int idx, double x, const double *tab)
{
    constexpr double XL = RysSmem<NROOTS>::XL;
    if (x >= XL) {
        double r = tab[RysSmem<NROOTS>::CHEB_DOUBLES + idx];
        double xi = 1. / x;
        return idx < NROOTS ? r * xi : r * sqrt(xi);
    }
    int it = (int)(x * 0.4);
    double s = (x - it * 2.5) * 0.8 - 1.;
    const double2 *c2 = (const double2 *)(tab + (it * 2 * NROOTS + idx) * RYS_NCHEB_);
    double c[RYS_NCHEB_];
#pragma unroll
    for (int k = 0; k < RYS_NCHEB_ / 2; ++k) {
        double2 v = c2[k];
        c[2 * k] = v.x;
        c[2 * k + 1] = v.y;
    }
    double s2 = s + s;
    double b1 = c[RYS_NCHEB_ - 1];
    double b2 = 0.;
#pragma unroll
    for (int k = RYS_NCHEB_ - 2; k > 0; --k) {
        double t = fma(s2, b1, c[k] - b2);     // c[k]-b2 is off the critical path
        b2 = b1;
        b1 = t;
    }
    return fma(s, b1, c[0] - b2);
}
