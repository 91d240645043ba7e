// Launch-shape selection for one class instantiation (plays the role of the
// reference's threads_scheme_for_k, gpu4pyscf/lib/gvhf-rys/rys_contract_k.cu:542-597).
#pragma once
#include <stdlib.h>
#include "rt.h"
#include "rys_jk_kernel.cuh"
#include "rys_jk_thread.cuh"
#include "rys_jk_warp.cuh"
#include "rys_jk_rowt.cuh"


typedef int (*jk_launcher_t)(JKTask *t, int nsm, int smem_max, cudaStream_t stream);

// B200_LAUNCH_LOG=1: one stderr line per class launch with its shape and the resources the
// compiler gave the kernel (tools/launch_table.py joins these with the per-launch times)
template <class K>
static inline void jk_log_launch(const char *kind, K kernel, const JKTask *t, int lp, int lq, int lr, int ls,
                                 int ndm, int nt, int smem, int per_sm, int grid)
{
#ifndef CPU_EMUL
    static const char *on = getenv("B200_LAUNCH_LOG");
    if (on == nullptr || on[0] != '1') return;
    cudaFuncAttributes a;
    if (cudaFuncGetAttributes(&a, kernel) != cudaSuccess) return;
    fprintf(stderr, "JKLAUNCH %s %d%d%d%d prim %d%d%d%d ndm %d threads %d smem %d blocks_per_sm %d grid %d regs %d stack %d\n",
            kind, lp, lq, lr, ls, t->nprim_p, t->nprim_q, t->nprim_r, t->nprim_s, ndm, nt, smem, per_sm, grid,
            a.numRegs, (int)a.localSizeBytes);
#endif
}

// Split the Y lists into chunks so that there are enough (X pair, Y chunk) work
// items for `workers` independent warps/teams; returns the number of items of
// this rank.
static inline int jk_set_chunks(JKTask *t, int workers)
{
    int ch = 1024;
    static const char *force = getenv("B200_YCHUNK");     // tests: exercise the chunk logic on tiny inputs
    if (force != nullptr && atoi(force) > 0) ch = atoi(force);
    else while (ch > 32 && (long long)t->npair_x * ((t->npair_y + ch - 1) / ch) < 8LL * workers * t->world) ch >>= 1;
    // the flat kernels index work items with 32-bit integers (item = atomicAdd(head) * world + rank):
    // grow the chunks until the item count stays far below 2^31 (pair lists of millions of entries)
    while ((long long)t->npair_x * ((t->npair_y + ch - 1) / ch) > (1LL << 29)) ch <<= 1;
    t->ychunk = ch;
    t->nychunk = (t->npair_y + ch - 1) / ch;
    long long items = (long long)t->npair_x * t->nychunk;
    return (int)((items + t->world - 1) / t->world);
}

template <int LP, int LQ, int LR, int LS, int NCS, int NDM>
int jk_launch(JKTask *t, int nsm, int smem_max, cudaStream_t stream)
{
    typedef JKCfg<LP, LQ, LR, LS, NCS> C;
    static_assert(C::G <= 1024, "too many threads per quartet");
    int npq = t->nprim_p * t->nprim_q;
    int nrs = t->nprim_r * t->nprim_s;
    if (npq > JK_MAXPP || nrs > JK_MAXPP) {
        fprintf(stderr, "rys_jk: more than %d primitive pairs per shell pair\n", JK_MAXPP);
        return 1;
    }
    int slot = jk_slot_doubles(npq, C::NROOTS, C::G2SIZE, C::GSIZE);
    int nq = JK_TARGET_THREADS / C::G;
    if (nq < 1) nq = 1;
    while (nq > 1 && (jk_blk_doubles(NDM, C::NFRS, nq * C::G, nq) + nq * slot) * 8 > smem_max / 2) --nq;
    int smem = (jk_blk_doubles(NDM, C::NFRS, nq * C::G, nq) + nq * slot) * 8;
    if (smem > smem_max) {
        fprintf(stderr, "rys_jk: class (%d%d|%d%d) needs %d B of shared memory\n", LP, LQ, LR, LS, smem);
        return 1;
    }
    t->nq = nq;
    t->slot_doubles = slot;
    t->ychunk = t->npair_y;
    t->nychunk = 1;
    int nt = nq * C::G;
    int per_sm = rt_prepare(rys_jk_kernel<LP, LQ, LR, LS, NCS, NDM>, smem, nt);
    if (per_sm < 1) {
        fprintf(stderr, "rys_jk: class (%d%d|%d%d) cannot be made resident (smem %d, threads %d)\n",
                LP, LQ, LR, LS, smem, nt);
        return 1;
    }
    int nitems = (t->npair_x + t->world - 1) / t->world;
    int grid = nsm * per_sm;
    if (grid > nitems) grid = nitems;
    if (grid < 1) return 0;
    jk_log_launch("block", rys_jk_kernel<LP, LQ, LR, LS, NCS, NDM>, t, LP, LQ, LR, LS, NDM, nt, smem, per_sm, grid);
    KLAUNCH((rys_jk_kernel<LP, LQ, LR, LS, NCS, NDM>), grid, nt, smem, stream, *t);
    return rt_last_error("rys_jk_kernel");
}

// thread-per-quartet register kernel (small classes)
template <int LP, int LQ, int LR, int LS, int NDM>
int jk_launch_thread(JKTask *t, int nsm, int smem_max, cudaStream_t stream)
{
    typedef JKCfg<LP, LQ, LR, LS, 1> C;
    int npq = t->nprim_p * t->nprim_q;
    int nrs = t->nprim_r * t->nprim_s;
    if (npq > JK_MAXPP || nrs > JK_MAXPP) {
        fprintf(stderr, "rys_jk: more than %d primitive pairs per shell pair\n", JK_MAXPP);
        return 1;
    }
    int nt = JKT_THREADS;
    int smem = (RysSmem<C::NROOTS>::DOUBLES + (nt / 32) * jkt_warp_doubles(nrs, NDM, C::NFRS)) * 8;
    t->nq = nt;
    t->slot_doubles = 0;
    int per_sm = rt_prepare(rys_jk_thread_kernel<LP, LQ, LR, LS, NDM>, smem, nt);
    if (per_sm < 1) {
        fprintf(stderr, "rys_jk_thread: class (%d%d|%d%d) cannot be made resident\n", LP, LQ, LR, LS);
        return 1;
    }
    int grid = nsm * per_sm;
    int nitems = jk_set_chunks(t, grid * (nt / 32));
    if (grid * (nt / 32) > nitems) grid = (nitems + nt / 32 - 1) / (nt / 32);
    if (grid < 1) return 0;
    jk_log_launch("thread", rys_jk_thread_kernel<LP, LQ, LR, LS, NDM>, t, LP, LQ, LR, LS, NDM, nt, smem, per_sm, grid);
    KLAUNCH((rys_jk_thread_kernel<LP, LQ, LR, LS, NDM>), grid, nt, smem, stream, *t);
    return rt_last_error("rys_jk_thread_kernel");
}

// Work items of a row-stream launch: (X pair, block of consecutive rows).  Enough items for
// `workers` independent warps/teams, few enough that the per-item X setup and the J_rs flush
// stay amortised.
static inline long long jk_set_rowblocks(JKTask *t, int workers)
{
    int rpi = 16;
    static const char *force = getenv("B200_ROWS_PER_ITEM");   // tests: exercise the block logic on tiny inputs
    if (force != nullptr && atoi(force) > 0) rpi = atoi(force);
    else while (rpi > 1 && (long long)t->npair_x * ((t->nrows + rpi - 1) / rpi) < 16LL * workers * t->world) rpi >>= 1;
    t->rows_per_item = rpi;
    t->nrowblk = (t->nrows + rpi - 1) / rpi;
    long long items = (long long)t->npair_x * t->nrowblk;
    return (items + t->world - 1) / t->world;
}

// warp-team kernel (mid and high classes)
template <int LP, int LQ, int LR, int LS, int NCS, int NDM, int BL = 0, int ROW = 0>
int jk_launch_warp(JKTask *t, int nsm, int smem_max, cudaStream_t stream)
{
    typedef JKWCfg<LP, LQ, LR, LS, NCS, BL> C;
    static_assert(C::TW <= 8, "team larger than a 256-thread block");
    int npq = t->nprim_p * t->nprim_q;
    int nrs = t->nprim_r * t->nprim_s;
    if (npq > JK_MAXPP || nrs > JK_MAXPP) {
        fprintf(stderr, "rys_jk: more than %d primitive pairs per shell pair\n", JK_MAXPP);
        return 1;
    }
    int nteams = jkw_block_threads(C::TEAM, C::TILE * (1 + NDM)) / C::TEAM;
    const int nterm = t->ratio2 != 0. ? 2 : 1;       // operator terms evaluated in this pass
    const int nprimq = npq * nrs * nterm;
    if ((BL || ROW) && nprimq != 1) {
        fprintf(stderr, "rys_jk_warp: the b-loop and row-stream variants only handle uncontracted single-operator launches\n");
        return 1;
    }
    // g-sets per lane slot: fill (not overfill) the team in the root/recurrence phase,
    // whose work items are (g-set, root)
#ifndef JKW_K_BIAS
#define JKW_K_BIAS 0
#endif
    int K = C::TEAM / (C::NSW * C::NROOTS) + JKW_K_BIAS;      // lanes of the root/recurrence phase: (g-set, root)
    if (K > 4) K = 4;
    if (K < 1) K = 1;
    if (nprimq > 1 && K > nprimq) K = nprimq;
    if (ROW) K = C::KROW;                                     // compile-time shape of the row-stream kernels
    const int minb = jkw_minblocks(C::TILE, NDM, NCS == 1 ? C::SUBSIZE : LS + 1, C::TEAM);
    const int rowd = ROW ? 2 * NDM * C::NFP * (C::NFR + C::NFS) : 0;
    auto smem_of = [&](int k) {
        return (RysSmem<C::NROOTS>::DOUBLES
                + nteams * jkw_team_doubles(C::TEAM, C::NSW, k, npq, nrs, NDM, C::NFRS, C::GSET, rowd, C::redd(ROW), nterm)) * 8;
    };
    const int budget = (smem_max + 1024) / minb - 1024;
    while (!ROW && K > 1 && smem_of(K) > budget) --K;
    while (nteams > 1 && smem_of(K) > smem_max) --nteams;     // fewer teams per block for the fattest classes
    // mid-size tiles: one 384-thread block per SM when shared memory keeps the 128-thread blocks below the
    // 12 warps the registers allow (B200_BIG_BLOCKS=0 disables)
    {
        static const char *big = getenv("B200_BIG_BLOCKS");
        const int nt_big = jkw_bound_threads(C::TEAM, C::TILE * (1 + NDM), ROW);
        if ((big == nullptr || big[0] != '0') && nt_big > nteams * C::TEAM && smem_of(K) <= smem_max) {
            const int per_sm_small = rt_prepare(rys_jk_warp_kernel<LP, LQ, LR, LS, NCS, NDM, BL, ROW>, smem_of(K),
                                                nteams * C::TEAM);
            const int small_teams = nteams;
            nteams = nt_big / C::TEAM;
            if (smem_of(K) > smem_max || per_sm_small * small_teams >= nteams) nteams = small_teams;
        }
    }
    const int nt = nteams * C::TEAM;
    int smem = smem_of(K);
    if (smem > smem_max) {
        fprintf(stderr, "rys_jk_warp: class (%d%d|%d%d) needs %d B of shared memory\n", LP, LQ, LR, LS, smem);
        return 1;
    }
    t->nq = K;
    t->slot_doubles = jkw_team_doubles(C::TEAM, C::NSW, K, npq, nrs, NDM, C::NFRS, C::GSET, rowd, C::redd(ROW), nterm);
    int per_sm = rt_prepare(rys_jk_warp_kernel<LP, LQ, LR, LS, NCS, NDM, BL, ROW>, smem, nt);
    if (per_sm < 1) {
        fprintf(stderr, "rys_jk_warp: class (%d%d|%d%d) cannot be made resident (smem %d, threads %d)\n",
                LP, LQ, LR, LS, smem, nt);
        return 1;
    }
    int grid = nsm * per_sm;
    long long nitems = ROW ? jk_set_rowblocks(t, grid * nteams) : jk_set_chunks(t, grid * nteams);
    if ((long long)grid * nteams > nitems) grid = (int)((nitems + nteams - 1) / nteams);
    if (grid < 1) return 0;
    jk_log_launch(ROW ? "roww" : (BL ? "warpbl" : "warp"), rys_jk_warp_kernel<LP, LQ, LR, LS, NCS, NDM, BL, ROW>, t,
                  LP, LQ, LR, LS, NDM, nt, smem, per_sm, grid);
    KLAUNCH((rys_jk_warp_kernel<LP, LQ, LR, LS, NCS, NDM, BL, ROW>), grid, nt, smem, stream, *t);
    return rt_last_error("rys_jk_warp_kernel");
}

// thread-per-quartet row-stream kernel (light classes)
template <int LP, int LQ, int LR, int LS, int NDM>
int jk_launch_rowt(JKTask *t, int nsm, int smem_max, cudaStream_t stream)
{
    typedef JKCfg<LP, LQ, LR, LS, 1> C;
    int npq = t->nprim_p * t->nprim_q;
    int nrs = t->nprim_r * t->nprim_s;
    if (npq > JK_MAXPP || nrs > JK_MAXPP) {
        fprintf(stderr, "rys_jk: more than %d primitive pairs per shell pair\n", JK_MAXPP);
        return 1;
    }
    int nt = JKT_THREADS;
    int smem = (RysSmem<C::NROOTS>::DOUBLES
                + (nt / 32) * jkrt_warp_doubles(nrs, NDM, C::NFRS, C::NFP, C::NFR, C::NFS)) * 8;
    int per_sm = rt_prepare(rys_jk_rowt_kernel<LP, LQ, LR, LS, NDM>, smem, nt);
    if (per_sm < 1) {
        fprintf(stderr, "rys_jk_rowt: class (%d%d|%d%d) cannot be made resident\n", LP, LQ, LR, LS);
        return 1;
    }
    int grid = nsm * per_sm;
    long long nitems = jk_set_rowblocks(t, grid * (nt / 32));
    if ((long long)grid * (nt / 32) > nitems) grid = (int)((nitems + nt / 32 - 1) / (nt / 32));
    if (grid < 1) return 0;
    jk_log_launch("rowt", rys_jk_rowt_kernel<LP, LQ, LR, LS, NDM>, t, LP, LQ, LR, LS, NDM, nt, smem, per_sm, grid);
    KLAUNCH((rys_jk_rowt_kernel<LP, LQ, LR, LS, NDM>), grid, nt, smem, stream, *t);
    return rt_last_error("rys_jk_rowt_kernel");
}
