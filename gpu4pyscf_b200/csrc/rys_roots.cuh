// Rys quadrature nodes/weights from the mpmath-generated tables (rys_tables.h,
// tools/gen_rys_tables.py).  Plays the role of the reference's rys_roots()
// (gpu4pyscf/lib/gvhf-rys/rys_roots.cu:25-88) with a different decomposition:
// one call evaluates ONE of the 2*NROOTS values, so the 2*NROOTS Clenshaw
// recurrences of a primitive quartet run on 2*NROOTS different threads.
#pragma once
#include "common.cuh"

#define RYS_NINT_ 40
#define RYS_NCHEB_ 14

// The generated tables (rys_tables.h) are Chebyshev series on s in [-1,1].  They are
// converted ONCE, on the host at upload time, to the monomial basis: the series decay
// fast enough (c_13/c_0 ~ 1e-14) that Horner on the monomial coefficients agrees with
// Clenshaw to 8e-16 relative (checked over all 5280 rows), and it needs one FMA per
// coefficient instead of two dependent operations.
static inline void rys_monomial_rows(double *tab, int nrows)
{
    long double T[RYS_NCHEB_][RYS_NCHEB_] = {{0}};
    T[0][0] = 1;
    T[1][1] = 1;
    for (int n = 1; n + 1 < RYS_NCHEB_; ++n)
        for (int k = 0; k < RYS_NCHEB_; ++k)
            T[n + 1][k] = (k > 0 ? 2 * T[n][k - 1] : 0) - T[n - 1][k];
    for (int r = 0; r < nrows; ++r) {
        double *c = tab + (size_t)r * RYS_NCHEB_;
        long double a[RYS_NCHEB_];
        for (int k = 0; k < RYS_NCHEB_; ++k) {
            a[k] = 0;
            for (int n = k; n < RYS_NCHEB_; ++n) a[k] += (long double)c[n] * T[n][k];
        }
        for (int k = 0; k < RYS_NCHEB_; ++k) c[k] = (double)a[k];
    }
}

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
    // the device tables hold the series in the monomial basis (rys_monomial_rows): Horner
    // even / odd halves in s^2: two independent Horner chains of half the depth
    const double s2 = s * s;
    double ve = c[RYS_NCHEB_ - 2], vo = c[RYS_NCHEB_ - 1];
#pragma unroll
    for (int k = RYS_NCHEB_ - 4; k >= 0; k -= 2) {
        ve = fma(ve, s2, c[k]);
        vo = fma(vo, s2, c[k + 1]);
    }
    return fma(vo, s, ve);
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
    // monomial coefficients; even / odd halves in s^2 = two independent Horner chains
    const double s2 = s * s;
    double ve = c[RYS_NCHEB_ - 2], vo = c[RYS_NCHEB_ - 1];
#pragma unroll
    for (int k = RYS_NCHEB_ - 4; k >= 0; k -= 2) {
        ve = fma(ve, s2, c[k]);
        vo = fma(vo, s2, c[k + 1]);
    }
    return fma(vo, s, ve);
}
