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
