// Shared definitions for the sm_100a direct-SCF J/K kernels.
//
// Data model (libcint slot layout, mirrored by the reference at
// gpu4pyscf/lib/gvhf-rys/vhf.cuh:5-25 and consumed through the RysIntEnvVars
// struct, vhf.cuh:47-54 / gpu4pyscf/gto/mole.py:1054-1082).
#pragma once
#include <stdint.h>

#ifdef CPU_EMUL
#include "cuda_shim.h"
#else
#include <cuda_runtime.h>
#endif

#define ATOM_OF       0
#define ANG_OF        1
#define NPRIM_OF      2
#define NCTR_OF       3
#define PTR_EXP       5
#define PTR_COEFF     6
#define PTR_BAS_COORD 7
#define BAS_SLOTS     8

#define LMAX_B200 4           // highest angular momentum handled on the GPU
#define PI_FAC 34.98683665524972497   // 2*pi^2.5

struct RysIntEnvVars {
    int natm;
    int nbas;
    int *atm;       // device
    int *bas;       // device
    double *env;    // device
    int *ao_loc;    // device
};

// One class launch: quartets (pq|rs), X = (r,s) pairs are block-resident ("tile
// side": their Cartesian components live in one thread's registers), Y = (p,q)
// pairs are streamed and their components are spread over threads.
struct JKTask {
    double *vj;             // [n_dm, nao, nao] or null
    double *vk;             // [n_dm, nao, nao] or null
    const double *dm;       // [n_dm, nao, nao]
    int n_dm;
    int nao;
    int nbas;
    const int *bas;         // device copies
    const double *env;
    const int *ao_loc;
    const uint32_t *pair_x; // ish*nbas+jsh, sorted by q_x descending
    const uint32_t *pair_y;
    const float *q_x;
    const float *q_y;
    int npair_x;
    int npair_y;
    const float *dm_cond;   // [nbas, nbas] log max|D_block|
    float cutoff;           // log(direct_scf_tol)
    float dm_penalty;       // max(dm_cond)
    float q_margin;         // 0 for exactly q-sorted Y lists; 4 for the reference's binned order
    int dedupe;             // 0: none; 1: same list, keep iy >= ix (1/2 on ix==iy);
                            // 2: keep pair_id(ij) >= pair_id(kl) (reference semantics); 3: reverse of 2
    int x_is_bra;           // only for dedupe 2/3 bookkeeping
    int *head;              // device work counter (zeroed by the launcher)
    int rank;               // items item%world==rank are processed by this device
    int world;
    int nprim_p, nprim_q, nprim_r, nprim_s;
    const double *rys_cheb; // device tables (rys_tables.h)
    const double *rys_large;
    unsigned long long *counter;  // optional: number of quartets evaluated (measurement)
    int nq;                 // quartet slots per block
    int slot_doubles;       // shared-memory doubles per slot
    double scale;           // factor applied to every integral (lr/sr mixing of RYS_build_k); 1 by default
    double omega;           // 0: 1/r; > 0: long-range erf(omega r)/r (reference: rys_roots_rs, rys_roots.cu:91-119)
    // optional second operator term evaluated in the same pass (range-separated exchange in ONE build:
    // the reference concatenates the full-range and the negated long-range roots, rys_roots_for_k.cu:45-60):
    // integral = scale * ( O[omega] + ratio2 * O[omega2] ),  ratio2 == 0: single term
    double ratio2;
    double omega2;
    // short-range exchange filter (K-only launches of the pure short-range operator; reference:
    // _fill_sr_vk_tasks, create_tasks.cu:245-369): per pair of the X / Y list the estimator s
    // (fill_s_estimator) and (centre, exponent sum) of the most diffuse primitives; null = off
    const float *sr_s_x;
    const float *sr_s_y;
    const float4 *sr_g_x;
    const float4 *sr_g_y;
    float sr_omega2;        // omega^2 of the attenuation
    const double *rec_y;    // optional per-pair records of the Y list (pair_record_kernel), or null
    int rec_stride;         // doubles per record
    int ychunk;             // work item = (X pair, chunk of `ychunk` consecutive Y candidates)
    int nychunk;            // ceil(npair_y / ychunk)
    // ---- row-stream kernels (rys_jk_rowt.cuh, rys_jk_roww.cuh): the Y list is grouped in rows
    // (all partners q of one shell p are consecutive, pair_y = p*nbas+q in (p,q) = kernel order)
    const int *row_ptr;     // [nrows + 1] offsets into pair_y / q_y / rec_y
    const int *row_p;       // [nrows] the shell p of each row
    const float *row_qmax;  // [nrows] largest q_y of the row
    int nrows;
    int rows_per_item;      // work item = (X pair, block of rows_per_item consecutive rows)
    int nrowblk;            // ceil(nrows / rows_per_item)
    const double *dmT;      // [n_dm, nao, nao] transposed density matrices (== dm when D is symmetric)
    double *vkT;            // [n_dm, nao, nao] receives the q-row K blocks transposed (== vk for symmetric D)
};

// Density threshold of the short-range filter for the quartet of pairs x and y: the largest exchange
// density block must exceed it (and so must dm_penalty).  th_x = omega^2 a_x / (a_x + omega^2).
// Same arithmetic as the reference (create_tasks.cu:343-352), FP32.
__device__ __forceinline__ float sr_density_cutoff(const float4 gx, const float sx, const float th_x,
                                                   const float4 gy, const float sy, const float cutoff)
{
    const float theta = th_x * gy.w / (th_x + gy.w);
    const float dx = gx.x - gy.x, dy = gx.y - gy.y, dz = gx.z - gy.z;
    const float rr = dx * dx + dy * dy + dz * dz;
    return cutoff - sx - sy + logf(rr + 1.f) + theta * rr;
}

// The filter for candidate y of the Y list against X pair ix; everything is read here (L1/L2 hits, only
// in short-range builds) so that the kernels keep no filter state in registers on their common path.
// (The task fields are passed by value: taking the address of a kernel parameter would copy the whole
// struct to local memory.  Not inlined on purpose: the registers the call clobbers are saved inside the
// short-range branch only, so builds without the filter execute no extra spill code.)
static __device__ __noinline__ bool sr_keep_(const float4 *g_x, const float *s_x, const float4 *g_y, const float *s_y,
                                             const float omega2, const float cutoff, const float dm_penalty,
                                             const int ix, const int y, const float dmax)
{
    const float4 gx = g_x[ix];
    const float th_x = omega2 * gx.w / (gx.w + omega2);
    const float dc = sr_density_cutoff(gx, s_x[ix], th_x, g_y[y], s_y[y], cutoff);
    return dm_penalty > dc && dmax > dc;
}
#define sr_keep(t, ix, y, dmax) \
    sr_keep_((t).sr_g_x, (t).sr_s_x, (t).sr_g_y, (t).sr_s_y, (t).sr_omega2, (t).cutoff, (t).dm_penalty, ix, y, dmax)

// Per-pair record (precomputed once per _VHFOpt.build, in pair-list order):
//   [0..2] A-B, [3..5] A, [6] (int p0, int q0) AO offsets, [7] 0.5 if ish == jsh else 1,
//   then per primitive pair: zeta, Px, Py, Pz, ci*cj*exp(-ai*aj/zeta*|AB|^2)
#define PAIR_REC_HDR 8
__host__ __device__ inline int pair_rec_stride(int npq) { return (PAIR_REC_HDR + 5 * npq + 1) & ~1; }

// ---- Cartesian component tables (lexical order xx, xy, xz, yy, yz, zz ...;
// same convention as libcint / the reference's rys_constant.cu:21-112, but
// computed, not tabulated).
__host__ __device__ constexpr int ncart(int l) { return (l + 1) * (l + 2) / 2; }

__host__ __device__ constexpr int cart_x(int l, int n)
{
    // n-th lexical component: x runs l..0, y runs l-x..0
    int idx = 0;
    for (int x = l; x >= 0; --x) {
        int cnt = l - x + 1;
        if (n < idx + cnt) return x;
        idx += cnt;
    }
    return 0;
}
__host__ __device__ constexpr int cart_y(int l, int n)
{
    int idx = 0;
    for (int x = l; x >= 0; --x) {
        int cnt = l - x + 1;
        if (n < idx + cnt) return (l - x) - (n - idx);
        idx += cnt;
    }
    return 0;
}
__host__ __device__ constexpr int cart_z(int l, int n)
{
    return l - cart_x(l, n) - cart_y(l, n);
}

// N consecutive doubles of a density-matrix row: 16-byte loads when the caller has
// checked that the row start is 16-byte aligned (even AO index, even nao)
template <int N>
__device__ __forceinline__ void ld_row(double (&dst)[N], const double *__restrict__ p, const bool vec)
{
    if (vec) {
#pragma unroll
        for (int i = 0; i + 1 < N; i += 2) {
#ifdef CPU_EMUL
            dst[i] = p[i];
            dst[i + 1] = p[i + 1];
#else
            const double2 v = __ldg((const double2 *)(p + i));
            dst[i] = v.x;
            dst[i + 1] = v.y;
#endif
        }
        if (N & 1) dst[N - 1] = __ldg(p + N - 1);
    } else {
#pragma unroll
        for (int i = 0; i < N; ++i) dst[i] = __ldg(p + i);
    }
}
