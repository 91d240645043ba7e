// Row-stream variant of the thread-per-quartet J/K kernel (sm_100a, FP64).
//
// The Y = (p,q) pair list is grouped in ROWS: all partners q of one shell p are
// consecutive (row_ptr / row_p).  A warp owns a work item (X = (r,s) pair, block
// of rows) and walks the rows one at a time, so THREE of the four shells of
// every quartet it evaluates are warp-uniform and only q differs between lanes:
//   * D_pr, D_ps, D_rs are staged once per row / item in shared memory (uniform
//     reads), K_pr, K_ps, J_rs accumulate in registers over the whole row / item
//     and leave the warp through one shared-memory tree reduction per row;
//   * the q-dependent blocks D_qr, D_qs, D_pq, K_qr, K_qs, J_pq are addressed
//     with q as the COLUMN index (dmT / vkT hold the transposed matrices), so the
//     lanes of a warp -- consecutive partners of p, i.e. neighbouring shells --
//     touch neighbouring addresses: per quartet nf_q*(nf_p+nf_r+nf_s) loads and
//     atomics instead of the (nf_p+nf_q)*(nf_r+nf_s)+nf_p*nf_q scattered ones of
//     the flat-list kernel (rys_jk_thread.cuh).
// Screening is the reference's predicate (create_tasks.cu:113-180) evaluated per
// candidate; a row is skipped as a whole when its largest Schwarz bound fails.
//
// Same arithmetic as rys_jk_thread_kernel; replaces, like it, the reference's
// unrolled rys_jk_XXXX kernels (lib/gvhf-rys/unrolled_rys_jk.cu) and the digestion
// of rys_contract_k.cuh:190-232 for the light angular classes.
#pragma once
#include "common.cuh"
#include "rys_roots.cuh"
#include "rys_jk_kernel.cuh"
#include "rys_jk_thread.cuh"

#define JKR_RED_ROWS 8          // values reduced per round of the shared-memory tree
#define JKR_RED_STRIDE 33
#define JKR_SCREEN 4            // candidates screened per lane and pass (independent loads in flight)
#define JKR_QCAP (32 + 32 * JKR_SCREEN)   // survivor queue entries of a warp

// shared-memory doubles of one warp: X primitive pairs, D_rs, D_p*, 64-entry queue, reduction scratch
__host__ __device__ inline int jkrt_warp_doubles(int nrs, int ndm, int nfrs, int nfp, int nfr, int nfs)
{
    return (nrs * 5 + ndm * nfrs + ndm * nfp * (nfr + nfs) + JKR_QCAP / 2 + JKR_RED_ROWS * JKR_RED_STRIDE + 1) & ~1;
}

__host__ __device__ constexpr int jkrt_minblocks(int nf_total, int nacc)
{
#ifdef JKRT_MINBLOCKS_FIXED
    return JKRT_MINBLOCKS_FIXED;
#endif
    // registers ~ 2 * (integrals + accumulators) + ~70 of addressing / recurrence state
    return nf_total + nacc <= 4 ? 5 : (nf_total + nacc <= 12 ? 4 : (nf_total + nacc <= 48 ? 3 : 2));
}

// Sum `v[0..N)` of every lane over the warp and hand the N totals to `sink(i, total)`
// (executed by one lane per value).  Shared-memory transpose in rounds of JKR_RED_ROWS
// values: 8 stores, 8 loads and two shuffle steps per round instead of five shuffle
// steps per value.
template <int N, class Sink>
__device__ __forceinline__ void jkr_warp_reduce(const double (&v)[N], double *red, const int lane, Sink sink)
{
    const int rv = lane >> 2, part = lane & 3;
#pragma unroll
    for (int v0 = 0; v0 < N; v0 += JKR_RED_ROWS) {
#pragma unroll
        for (int k = 0; k < JKR_RED_ROWS; ++k)
            if (v0 + k < N) red[k * JKR_RED_STRIDE + lane] = v[v0 + k];
        __syncwarp();
        double s = 0.;
        if (v0 + rv < N) {
#pragma unroll
            for (int i = 0; i < 8; ++i) s += red[rv * JKR_RED_STRIDE + part * 8 + i];
        }
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        if (part == 0 && v0 + rv < N && s != 0.) sink(v0 + rv, s);
        __syncwarp();
    }
}

template <int LP, int LQ, int LR, int LS, int NDM>
__global__ void
__launch_bounds__(JKT_THREADS, jkrt_minblocks(ncart(LP) * ncart(LQ) * ncart(LR) * ncart(LS),
                                              NDM * (ncart(LP) * (ncart(LR) + ncart(LS)) + ncart(LR) * ncart(LS))))
rys_jk_rowt_kernel(JKTask t)
{
    typedef JKCfg<LP, LQ, LR, LS, 1> C;
    constexpr int NFP = C::NFP, NFQ = C::NFQ, NFR = C::NFR, NFS = C::NFS;
    constexpr int NFPQ = C::NFPQ, NFRS = C::NFRS;
    constexpr int NROOTS = C::NROOTS;
    constexpr int NKP = NFP * (NFR + NFS);          // K_pr | K_ps values of one density matrix

#ifdef CPU_EMUL
    double *smem = (double *)emul_dynamic_smem;
#else
    extern __shared__ double smem[];
#endif
    const int tid = threadIdx.x;
    const int npq = t.nprim_p * t.nprim_q;
    const int nrs = t.nprim_r * t.nprim_s;
    const int nterm = t.ratio2 != 0. ? 2 : 1;
    const int nbas = t.nbas;
    const int nao = t.nao;
    const size_t nao2 = (size_t)nao * nao;
    const bool do_j = t.vj != nullptr;
    const bool do_k = t.vk != nullptr;

    const int lane = tid & 31;
    const int warp = tid >> 5;
    double *rys_tab = smem;
    double *wsm = smem + RysSmem<NROOTS>::DOUBLES + warp * jkrt_warp_doubles(nrs, NDM, NFRS, NFP, NFR, NFS);
    double *rsprim = wsm;                            // nrs * 5: eta, Qx, Qy, Qz, coef
    double *drs = rsprim + nrs * 5;                  // NDM * NFRS
    double *dpx = drs + NDM * NFRS;                  // NDM * NFP * (NFR + NFS): D_pr | D_ps
    uint32_t *queue = (uint32_t *)(dpx + NDM * NKP); // JKR_QCAP entries
    double *red = dpx + NDM * NKP + JKR_QCAP / 2;
    rys_stage_tables<NROOTS>(rys_tab, t.rys_cheb, t.rys_large, tid, blockDim.x);
    __syncthreads();

    const int *bas = t.bas;
    const double *env = t.env;
    const int *ao_loc = t.ao_loc;
    const long long nitems = (long long)t.npair_x * t.nrowblk;

    while (true) {
        __syncwarp();
        long long item = 0;
        if (lane == 0) item = (long long)atomicAdd((unsigned long long *)t.head, 1ull) * t.world + t.rank;
        item = __shfl_sync(0xffffffffu, item, 0);
        if (item >= nitems) break;
        // row-block major: concurrent warps share the Y rows (L1/L2), X pairs heaviest first
        const int rb = (int)(item / t.npair_x);
        const int ix = (int)(item - (long long)rb * t.npair_x);
        const float thr = t.cutoff - t.q_x[ix];
        const int row_begin = rb * t.rows_per_item;
        const int row_end = min(t.nrows, row_begin + t.rows_per_item);
        // whole item dead?  (row_qmax is not sorted: test every row below, but skip the X setup when none survives)
        bool any_row = false;
        for (int row = row_begin + lane; row < row_end; row += 32)
            any_row = any_row || (t.row_qmax[row] + t.dm_penalty >= thr);
        if (!__any_sync(0xffffffffu, any_row)) continue;

        const uint32_t pair_rs = t.pair_x[ix];
        const int rsh = pair_rs / nbas;
        const int ssh = pair_rs - rsh * nbas;
        const int r0 = ao_loc[rsh];
        const int s0 = ao_loc[ssh];
        const double *Rr = env + bas[rsh * BAS_SLOTS + PTR_BAS_COORD];
        const double *Rs = env + bas[ssh * BAS_SLOTS + PTR_BAS_COORD];
        const double Cx = Rr[0], Cy = Rr[1], Cz = Rr[2];
        const double CDx = Rr[0] - Rs[0], CDy = Rr[1] - Rs[1], CDz = Rr[2] - Rs[2];
        {
            const double *er = env + bas[rsh * BAS_SLOTS + PTR_EXP];
            const double *es = env + bas[ssh * BAS_SLOTS + PTR_EXP];
            const double *cr = env + bas[rsh * BAS_SLOTS + PTR_COEFF];
            const double *cs = env + bas[ssh * BAS_SLOTS + PTR_COEFF];
            const double rr2 = CDx * CDx + CDy * CDy + CDz * CDz;
            for (int i = lane; i < nrs; i += 32) {
                const int ir = i / t.nprim_s;
                const int is = i - ir * t.nprim_s;
                const double ar = er[ir], as = es[is];
                const double eta = ar + as;
                const double f = as / eta;
                double *o = rsprim + i * 5;
                o[0] = eta;
                o[1] = Cx - f * CDx;
                o[2] = Cy - f * CDy;
                o[3] = Cz - f * CDz;
                o[4] = cr[ir] * cs[is] * exp(-ar * f * rr2);
            }
            for (int i = lane; i < NDM * NFRS; i += 32) {
                const int idm = i / NFRS;
                const int cd = i - idm * NFRS;
                const int c = cd / NFS;
                const int d = cd - c * NFS;
                drs[i] = __ldg(t.dm + idm * nao2 + (size_t)(r0 + c) * nao + s0 + d);
            }
        }
        double jacc[NDM * NFRS];
#pragma unroll
        for (int n = 0; n < NDM * NFRS; ++n) jacc[n] = 0.;

        const float *dc_r = t.dm_cond + (size_t)rsh * nbas;
        const float *dc_s = t.dm_cond + (size_t)ssh * nbas;
        const float d_rs = dc_r[ssh];
        const double fac_rs = (rsh == ssh ? .5 : 1.) * PI_FAC * t.scale;

        for (int row = row_begin; row < row_end; ++row) {
            if (t.row_qmax[row] + t.dm_penalty < thr) continue;
            const int psh = t.row_p[row];
            const int p0 = ao_loc[psh];
            const float *dc_p = t.dm_cond + (size_t)psh * nbas;
            // warp-uniform part of the density bound
            float d_uni = -1e30f;
            if (do_j) d_uni = d_rs;
            if (do_k) d_uni = fmaxf(d_uni, fmaxf(dc_r[psh], dc_s[psh]));
            __syncwarp();
            if (do_k) {
                for (int i = lane; i < NDM * NKP; i += 32) {
                    const int idm = i / NKP;
                    const int ak = i - idm * NKP;
                    const int a = ak / (NFR + NFS);
                    const int k = ak - a * (NFR + NFS);
                    const int col = k < NFR ? r0 + k : s0 + k - NFR;
                    dpx[i] = __ldg(t.dm + idm * nao2 + (size_t)(p0 + a) * nao + col);
                }
            }
            double kacc[NDM * NKP];
#pragma unroll
            for (int n = 0; n < NDM * NKP; ++n) kacc[n] = 0.;
            int ypos = t.row_ptr[row];
            const int y_end = t.row_ptr[row + 1];
            int qn = 0, qh = 0;                      // survivors queue[qh, qh + qn)
            bool exhausted = false;
            __syncwarp();

            while (true) {
                // ---------------- screen partners q of p into the warp's queue ----------------
                // JKR_SCREEN candidates per lane and pass: the dependent loads (pair id -> density
                // bounds of q) of all of them are in flight together
                if (!exhausted && qn < 32 && qh > 0) {      // fewer than 32 left: move them to the front
                    uint32_t keepv = 0;
                    if (lane < qn) keepv = queue[qh + lane];
                    __syncwarp();
                    if (lane < qn) queue[lane] = keepv;
                    qh = 0;
                    __syncwarp();
                }
                while (!exhausted && qn < 32) {
                    float qy[JKR_SCREEN];
                    int qsh[JKR_SCREEN];
#pragma unroll
                    for (int k = 0; k < JKR_SCREEN; ++k) {
                        const int y = ypos + k * 32 + lane;
                        const bool in = y < y_end;
                        qy[k] = in ? t.q_y[y] : -1e30f;
                        qsh[k] = in ? (int)(t.pair_y[y] - (uint32_t)psh * nbas) : psh;
                    }
                    float dq[JKR_SCREEN];
#pragma unroll
                    for (int k = 0; k < JKR_SCREEN; ++k) {
                        float dmax = d_uni;
                        if (do_j) dmax = fmaxf(dmax, dc_p[qsh[k]]);
                        if (do_k) dmax = fmaxf(dmax, fmaxf(dc_r[qsh[k]], dc_s[qsh[k]]));
                        dq[k] = dmax;
                    }
#pragma unroll
                    for (int k = 0; k < JKR_SCREEN; ++k) {
                        bool keep = qy[k] + t.dm_penalty >= thr && dq[k] + qy[k] > thr;
                        if (t.sr_s_x != nullptr && keep) keep = sr_keep(t, ix, ypos + k * 32 + lane, dq[k]);
                        if (t.dedupe) {
                            const uint32_t canon = psh >= qsh[k] ? (uint32_t)psh * nbas + qsh[k]
                                                                 : (uint32_t)qsh[k] * nbas + psh;
                            keep = keep && canon >= pair_rs;
                        }
                        const unsigned m = __ballot_sync(0xffffffffu, keep);
                        if (keep) queue[qn + __popc(m & ((1u << lane) - 1u))] = (uint32_t)(ypos + k * 32 + lane);
                        qn += __popc(m);
                    }
                    ypos += 32 * JKR_SCREEN;
                    exhausted = ypos >= y_end;
                    __syncwarp();
                }
                if (qn == 0) break;
                const int nb = qn < 32 ? qn : 32;

                if (lane < nb) {
                    const int iy = queue[qh + lane];
                    const double *rec = t.rec_y + (size_t)iy * t.rec_stride;
                    const double2 h0 = *(const double2 *)(rec + 0);
                    const double2 h1 = *(const double2 *)(rec + 2);
                    const double2 h2 = *(const double2 *)(rec + 4);
                    const double2 h3 = *(const double2 *)(rec + 6);
                    const double ABx = h0.x, ABy = h0.y, ABz = h1.x;
                    const double Ax = h1.y, Ay = h2.x, Az = h2.y;
                    const int q0 = ((const int *)&h3.x)[1];
                    double fac = fac_rs * h3.y;                  // 1/2 when psh == qsh
                    if (t.dedupe) {
                        const int qsh = t.pair_y[iy] - (uint32_t)psh * nbas;
                        const uint32_t canon = psh >= qsh ? (uint32_t)psh * nbas + qsh : (uint32_t)qsh * nbas + psh;
                        if (canon == pair_rs) fac *= .5;
                    }

                    double gout[NFPQ * NFRS];
#pragma unroll
                    for (int n = 0; n < NFPQ * NFRS; ++n) gout[n] = 0.;
                    for (int ipq = 0; ipq < npq; ++ipq) {
                        const double *pr = rec + PAIR_REC_HDR + ipq * 5;
                        const double zeta = pr[0], Px = pr[1], Py = pr[2], Pz = pr[3];
                        const double cpq = fac * pr[4];
                        for (int it = 0; it < nrs * nterm; ++it) {       // operator terms: one call site
                            const int term = it / nrs;
                            const int irs = it - term * nrs;
                            jkt_prim_quartet<LP, LQ, LR, LS>(gout, zeta, Px, Py, Pz, term ? cpq * t.ratio2 : cpq,
                                                             Ax, Ay, Az, ABx, ABy, ABz,
                                                             rsprim + irs * 5, Cx, Cy, Cz, CDx, CDy, CDz,
                                                             term ? t.omega2 : t.omega, rys_tab);
                        }
                    }

                    // ---------------- digestion: q is the column index of every global access ----------------
#pragma unroll
                    for (int idm = 0; idm < NDM; ++idm) {
                        if (do_j) {
                            const double *dmr = t.dm + idm * nao2 + (size_t)p0 * nao + q0;
                            double *vjr = t.vj + idm * nao2 + (size_t)p0 * nao + q0;
#pragma unroll
                            for (int a = 0; a < NFP; ++a)
#pragma unroll
                            for (int b = 0; b < NFQ; ++b) {
                                const double dpq = __ldg(dmr + (size_t)a * nao + b);
                                double s = 0.;
#pragma unroll
                                for (int cd = 0; cd < NFRS; ++cd) {
                                    const double gv = gout[(a * NFQ + b) * NFRS + cd];
                                    s += gv * drs[idm * NFRS + cd];
                                    jacc[idm * NFRS + cd] += gv * dpq;
                                }
                                JK_ATOMIC_ADD(vjr + (size_t)a * nao + b, s);
                            }
                        }
                        if (do_k) {
                            const double *dT = t.dmT + idm * nao2 + q0;
                            double *kT = t.vkT + idm * nao2 + q0;
                            const double *dp = dpx + idm * NKP;
#pragma unroll
                            for (int b = 0; b < NFQ; ++b) {
                                double dqr[NFR], dqs[NFS], kqr[NFR], kqs[NFS];
#pragma unroll
                                for (int c = 0; c < NFR; ++c) {
                                    dqr[c] = __ldg(dT + (size_t)(r0 + c) * nao + b);
                                    kqr[c] = 0.;
                                }
#pragma unroll
                                for (int d = 0; d < NFS; ++d) {
                                    dqs[d] = __ldg(dT + (size_t)(s0 + d) * nao + b);
                                    kqs[d] = 0.;
                                }
#pragma unroll
                                for (int a = 0; a < NFP; ++a)
#pragma unroll
                                for (int c = 0; c < NFR; ++c)
#pragma unroll
                                for (int d = 0; d < NFS; ++d) {
                                    const double gv = gout[((a * NFQ + b) * NFR + c) * NFS + d];
                                    kacc[idm * NKP + a * (NFR + NFS) + c] += gv * dqs[d];
                                    kacc[idm * NKP + a * (NFR + NFS) + NFR + d] += gv * dqr[c];
                                    kqr[c] += gv * dp[a * (NFR + NFS) + NFR + d];
                                    kqs[d] += gv * dp[a * (NFR + NFS) + c];
                                }
#pragma unroll
                                for (int c = 0; c < NFR; ++c) JK_ATOMIC_ADD(kT + (size_t)(r0 + c) * nao + b, kqr[c]);
#pragma unroll
                                for (int d = 0; d < NFS; ++d) JK_ATOMIC_ADD(kT + (size_t)(s0 + d) * nao + b, kqs[d]);
                            }
                        }
                    }
                }   // lane < nb

                __syncwarp();
                if (lane == 0 && t.counter != nullptr) atomicAdd(t.counter, (unsigned long long)nb);
                qh += nb;
                qn -= nb;
                if (exhausted && qn == 0) break;
            }

            // ---------------- row end: K_pr, K_ps leave the warp ----------------
            if (do_k) {
                jkr_warp_reduce<NDM * NKP>(kacc, red, lane, [&](const int v, const double s) {
                    const int idm = v / NKP;
                    const int ak = v - idm * NKP;
                    const int a = ak / (NFR + NFS);
                    const int k = ak - a * (NFR + NFS);
                    const int col = k < NFR ? r0 + k : s0 + k - NFR;
                    JK_ATOMIC_ADD(t.vk + idm * nao2 + (size_t)(p0 + a) * nao + col, s);
                });
            }
        }

        // ---------------- flush J_rs ----------------
        if (do_j) {
            jkr_warp_reduce<NDM * NFRS>(jacc, red, lane, [&](const int v, const double s) {
                const int idm = v / NFRS;
                const int cd = v - idm * NFRS;
                JK_ATOMIC_ADD(t.vj + idm * nao2 + (size_t)(r0 + cd / NFS) * nao + s0 + cd % NFS, s);
            });
        }
    }
}
