// Direct-SCF J/K kernel for one angular-momentum class (pq|rs), sm_100a, FP64.
//
// Replaces, with a different decomposition, the reference's rys_jk_kernel /
// rys_k_* family (gpu4pyscf/lib/gvhf-rys/rys_contract_jk.cu:33-333,
// unrolled_rys_jk.cu) including its in-kernel screening (_fill_vjk_tasks,
// create_tasks.cu:113-180), Rys roots, 2D recurrences (build_rys_gxyz.cuh) and
// J/K digestion (rys_contract_k.cuh:190-232).
//
// Decomposition
//   * X = (r,s) shell pair: block-resident.  The nf_r*nf_s Cartesian components
//     of the X pair are the per-thread register tile (split NCS ways along r
//     for f/g shells).
//   * Y = (p,q) shell pair: streamed from a q-sorted list, screened in-kernel
//     (Schwarz + density) into a shared-memory queue.  NQ quartets are in
//     flight per block ("slots"); the nf_p*nf_q components of the Y pair are
//     spread over the G = nf_p*nf_q*NCS threads of a slot.
//   * per primitive quartet: all NROOTS roots at once --
//       phase 1  2*NROOTS threads/slot: one Clenshaw series each (root or weight)
//       phase 2  3*NROOTS threads/slot: VRR+TRR for one (root, axis) in registers
//       phase 3  3*NROOTS*(lp+1)(lq+1) roles/slot: bra+ket HRR in registers
//       phase 4  every thread: gout[c][d] += gx*gy*gz for its (a,b) component
//   * digestion: J_rs accumulates in registers over the whole Y stream (one
//     flush per X pair); J_pq / K_pr / K_ps / K_qr / K_qs partial sums over the
//     thread tile, then global FP64 atomics.
#pragma once
#include "common.cuh"
#include "rys_roots.cuh"

template <int LP, int LQ, int LR, int LS, int NCS>
struct JKCfg {
    static constexpr int NFP = ncart(LP), NFQ = ncart(LQ), NFR = ncart(LR), NFS = ncart(LS);
    static constexpr int NFPQ = NFP * NFQ, NFRS = NFR * NFS;
    static constexpr int LPQ = LP + LQ, LRS = LR + LS;
    static constexpr int NROOTS = (LPQ + LRS) / 2 + 1;
    static constexpr int G2SIZE = (LPQ + 1) * (LRS + 1);
    static constexpr int GSIZE = (LP + 1) * (LQ + 1) * (LR + 1) * (LS + 1);
    static constexpr int NC = (NFR + NCS - 1) / NCS;
    static constexpr int G = NFPQ * NCS;
    static constexpr int SUBSIZE = (LR + 1) * (LS + 1);
};

// shared-memory layout, in doubles (used by the kernel and by the launcher)
#define JK_MAXPP 36                      // max primitive pairs of a shell pair
#define BLK_RSPRIM 0                     // nrs * 5 : eta, Qx, Qy, Qz, coef
#define BLK_RSGEOM (JK_MAXPP * 5)        // CD[3], C[3]
#define BLK_INTS   (BLK_RSGEOM + 6)      // 4 doubles = 8 ints
#define BLK_DRS    (BLK_INTS + 4)        // NDM * NFRS
#define SLOT_AB 0
#define SLOT_A 3
#define SLOT_FAC 6
#define SLOT_PQPRIM 8                    // npq * 5 : zeta, Px, Py, Pz, coef

__host__ __device__ inline int jk_blk_doubles(int ndm, int nfrs, int nthreads, int nq)
{
    int qcap = nthreads + nq;            // uint32 queue
    return BLK_DRS + ndm * nfrs + (qcap + 1) / 2 + 1;
}
__host__ __device__ inline int jk_slot_doubles(int npq, int nroots, int g2size, int gsize)
{
    int n = SLOT_PQPRIM + npq * 5 + 2 * nroots + nroots * 3 * (g2size + gsize);
    return n | 1;                        // odd stride: slots start in different banks
}

#ifndef JK_TARGET_THREADS
#define JK_TARGET_THREADS 256
#endif
#ifndef JK_MINBLOCKS
#define JK_MINBLOCKS 1
#endif
#ifdef JK_NO_ATOMICS       // measurement-only build: keeps the arithmetic, drops the global atomics
#define JK_ATOMIC_ADD(p, v) do { if ((v) == 1.2345e-300) atomicAdd(p, v); } while (0)
#else
#define JK_ATOMIC_ADD(p, v) atomicAdd(p, v)
#endif

template <int LP, int LQ, int LR, int LS, int NCS, int NDM>
__global__ void
__launch_bounds__((JKCfg<LP, LQ, LR, LS, NCS>::G > JK_TARGET_THREADS ? JKCfg<LP, LQ, LR, LS, NCS>::G : JK_TARGET_THREADS),
                  (JKCfg<LP, LQ, LR, LS, NCS>::G > JK_TARGET_THREADS / 2 ? 1 : JK_MINBLOCKS))
rys_jk_kernel(JKTask t)
{
    typedef JKCfg<LP, LQ, LR, LS, NCS> C;
    constexpr int NFQ = C::NFQ, NFR = C::NFR, NFS = C::NFS, NFRS = C::NFRS;
    constexpr int LPQ = C::LPQ, LRS = C::LRS, NROOTS = C::NROOTS;
    constexpr int G2SIZE = C::G2SIZE, GSIZE = C::GSIZE, NC = C::NC, G = C::G;
    constexpr int SUB = C::SUBSIZE;

#ifdef CPU_EMUL
    double *smem = (double *)emul_dynamic_smem;
#else
    extern __shared__ double smem[];
#endif
    const int tid = threadIdx.x;
    const int NT = blockDim.x;
    const int nq = t.nq;
    const int slot = tid / G;
    const int tg = tid - slot * G;
    const int ab = tg / NCS;
    const int cpart = tg - ab * NCS;
    const int a = ab / NFQ;
    const int b = ab - a * NFQ;
    const int c0 = cpart * NC;
    const int gbx = (cart_x(LP, a) * (LQ + 1) + cart_x(LQ, b)) * SUB;
    const int gby = (cart_y(LP, a) * (LQ + 1) + cart_y(LQ, b)) * SUB + GSIZE;
    const int gbz = (cart_z(LP, a) * (LQ + 1) + cart_z(LQ, b)) * SUB + 2 * GSIZE;

    const int npq = t.nprim_p * t.nprim_q;
    const int nrs = t.nprim_r * t.nprim_s;
    const int nbas = t.nbas;
    const int nao = t.nao;
    const size_t nao2 = (size_t)nao * nao;
    const bool do_j = t.vj != nullptr;
    const bool do_k = t.vk != nullptr;

    double *blk = smem;
    int *blk_ints = (int *)(blk + BLK_INTS);       // [0] item, [1] qn, [2] exhausted
    double *drs = blk + BLK_DRS;
    uint32_t *queue = (uint32_t *)(blk + BLK_DRS + NDM * NFRS);
    const int blk_doubles = jk_blk_doubles(NDM, NFRS, NT, nq);
    double *sl = smem + blk_doubles + slot * t.slot_doubles;
    double *sl_pq = sl + SLOT_PQPRIM;
    double *sl_rw = sl_pq + npq * 5;
    double *sl_g2 = sl_rw + 2 * NROOTS;
    double *sl_g = sl_g2 + NROOTS * 3 * G2SIZE;

    const int *bas = t.bas;
    const double *env = t.env;
    const int *ao_loc = t.ao_loc;

    // per-thread offsets of the tile rows when the tile is split along r
    int ox[NC], oy[NC], oz[NC];
#pragma unroll
    for (int n = 0; n < NC; ++n) {
        int c = c0 + n;
        if (c >= NFR) c = NFR - 1;
        ox[n] = cart_x(LR, c) * (LS + 1);
        oy[n] = cart_y(LR, c) * (LS + 1);
        oz[n] = cart_z(LR, c) * (LS + 1);
    }

    while (true) {
        // ---------------- next X pair ----------------
        __syncthreads();
        if (tid == 0) {
            blk_ints[0] = atomicAdd(t.head, 1) * t.world + t.rank;
            blk_ints[1] = 0;
            blk_ints[2] = 0;
        }
        __syncthreads();
        const int ix = blk_ints[0];
        if (ix >= t.npair_x) break;
        const uint32_t pair_rs = t.pair_x[ix];
        const int rsh = pair_rs / nbas;
        const int ssh = pair_rs - rsh * nbas;
        const int r0 = ao_loc[rsh];
        const int s0 = ao_loc[ssh];
        const float qx = t.q_x[ix];
        const float thr = t.cutoff - qx;
        {   // stage the X pair
            const double *Rr = env + bas[rsh * BAS_SLOTS + PTR_BAS_COORD];
            const double *Rs = env + bas[ssh * BAS_SLOTS + PTR_BAS_COORD];
            const double *er = env + bas[rsh * BAS_SLOTS + PTR_EXP];
            const double *es = env + bas[ssh * BAS_SLOTS + PTR_EXP];
            const double *cr = env + bas[rsh * BAS_SLOTS + PTR_COEFF];
            const double *cs = env + bas[ssh * BAS_SLOTS + PTR_COEFF];
            double dx = Rr[0] - Rs[0], dy = Rr[1] - Rs[1], dz = Rr[2] - Rs[2];
            double rr2 = dx * dx + dy * dy + dz * dz;
            for (int i = tid; i < nrs; i += NT) {
                int ir = i / t.nprim_s;
                int is = i - ir * t.nprim_s;
                double ar = er[ir], as = es[is];
                double eta = ar + as;
                double f = as / eta;
                double *o = blk + BLK_RSPRIM + i * 5;
                o[0] = eta;
                o[1] = Rr[0] - f * dx;
                o[2] = Rr[1] - f * dy;
                o[3] = Rr[2] - f * dz;
                o[4] = cr[ir] * cs[is] * exp(-ar * f * rr2);
            }
            if (tid < 3) {
                blk[BLK_RSGEOM + tid] = Rr[tid] - Rs[tid];
                blk[BLK_RSGEOM + 3 + tid] = Rr[tid];
            }
            for (int i = tid; i < NDM * NFRS; i += NT) {
                int idm = i / NFRS;
                int cd = i - idm * NFRS;
                int c = cd / NFS;
                int d = cd - c * NFS;
                drs[i] = __ldg(t.dm + idm * nao2 + (size_t)(r0 + c) * nao + s0 + d);
            }
        }
        double jacc[NDM][NC][NFS];
#pragma unroll
        for (int idm = 0; idm < NDM; ++idm)
#pragma unroll
            for (int n = 0; n < NC; ++n)
#pragma unroll
                for (int d = 0; d < NFS; ++d) jacc[idm][n][d] = 0.;

        int ypos = (t.dedupe == 1) ? ix : 0;
        const float *dc_r = t.dm_cond + (size_t)rsh * nbas;
        const float *dc_s = t.dm_cond + (size_t)ssh * nbas;
        const float d_rs = dc_r[ssh];
        __syncthreads();

        while (true) {
            // ---------------- fill the queue ----------------
            while (true) {
                int qn = blk_ints[1];
                int exhausted = blk_ints[2];
                __syncthreads();
                if (exhausted || qn >= nq) break;
                int y = ypos + tid;
                bool keep = false;
                if (y >= t.npair_y) {
                    blk_ints[2] = 1;
                } else {
                    float qy = t.q_y[y];
                    if (qy + t.dm_penalty + t.q_margin < thr) {
                        blk_ints[2] = 1;      // sorted: nothing further survives
                    } else {
                        uint32_t pair_pq = t.pair_y[y];
                        int psh = pair_pq / nbas;
                        int qsh = pair_pq - psh * nbas;
                        keep = qy + t.dm_penalty >= thr;
                        if (t.dedupe == 2) keep = keep && pair_rs >= pair_pq;
                        if (t.dedupe == 3) keep = keep && pair_pq >= pair_rs;
                        if (keep) {
                            const float *dc_p = t.dm_cond + (size_t)psh * nbas;
                            float dmax = -1e30f;
                            if (do_j) dmax = fmaxf(d_rs, dc_p[qsh]);
                            if (do_k) {
                                dmax = fmaxf(dmax, fmaxf(dc_r[psh], dc_r[qsh]));
                                dmax = fmaxf(dmax, fmaxf(dc_s[psh], dc_s[qsh]));
                            }
                            keep = dmax + qy > thr;
                            if (t.sr_s_x != nullptr && keep) keep = sr_keep(t, ix, y, dmax);     // common.cuh
                        }
                    }
                }
                if (keep) {
                    int pos = atomicAdd(&blk_ints[1], 1);
                    queue[pos] = (uint32_t)y;
                }
                ypos += NT;
                __syncthreads();
            }
            const int qn = blk_ints[1];
            const int exhausted = blk_ints[2];
            const int nsteps = exhausted ? (qn + nq - 1) / nq : qn / nq;

            for (int step = 0; step < nsteps; ++step) {
                // ---------------- slot setup ----------------
                int qi = step * nq + slot;
                const bool active = qi < qn;
                if (!active) qi = step * nq;
                const int iy = queue[qi];
                const uint32_t pair_pq = t.pair_y[iy];
                const int psh = pair_pq / nbas;
                const int qsh = pair_pq - psh * nbas;
                const int p0 = ao_loc[psh];
                const int q0 = ao_loc[qsh];
                {
                    const double *Rp = env + bas[psh * BAS_SLOTS + PTR_BAS_COORD];
                    const double *Rq = env + bas[qsh * BAS_SLOTS + PTR_BAS_COORD];
                    const double *ep = env + bas[psh * BAS_SLOTS + PTR_EXP];
                    const double *eq = env + bas[qsh * BAS_SLOTS + PTR_EXP];
                    const double *cp = env + bas[psh * BAS_SLOTS + PTR_COEFF];
                    const double *cq = env + bas[qsh * BAS_SLOTS + PTR_COEFF];
                    double dx = Rp[0] - Rq[0], dy = Rp[1] - Rq[1], dz = Rp[2] - Rq[2];
                    double rr2 = dx * dx + dy * dy + dz * dz;
                    for (int i = tg; i < npq; i += G) {
                        int ip = i / t.nprim_q;
                        int iq = i - ip * t.nprim_q;
                        double ap = ep[ip], aq = eq[iq];
                        double zeta = ap + aq;
                        double f = aq / zeta;
                        double *o = sl_pq + i * 5;
                        o[0] = zeta;
                        o[1] = Rp[0] - f * dx;
                        o[2] = Rp[1] - f * dy;
                        o[3] = Rp[2] - f * dz;
                        o[4] = cp[ip] * cq[iq] * exp(-ap * f * rr2);
                    }
                    if (tg == 0) {
                        sl[SLOT_AB + 0] = dx; sl[SLOT_AB + 1] = dy; sl[SLOT_AB + 2] = dz;
                        sl[SLOT_A + 0] = Rp[0]; sl[SLOT_A + 1] = Rp[1]; sl[SLOT_A + 2] = Rp[2];
                        double fac = PI_FAC * t.scale;
                        if (psh == qsh) fac *= .5;
                        if (rsh == ssh) fac *= .5;
                        if (pair_pq == pair_rs) fac *= .5;
                        if (!active) fac = 0.;
                        sl[SLOT_FAC] = fac;
                    }
                }
                double gout[NC][NFS];
#pragma unroll
                for (int n = 0; n < NC; ++n)
#pragma unroll
                    for (int d = 0; d < NFS; ++d) gout[n][d] = 0.;
                __syncthreads();

                for (int ipq = 0; ipq < npq; ++ipq) {
                for (int irs = 0; irs < nrs; ++irs) {
                    const double *pp = sl_pq + ipq * 5;
                    const double *rr = blk + BLK_RSPRIM + irs * 5;
                    // ---------------- phase 1: roots and weights ----------------
                    for (int idx = tg; idx < 2 * NROOTS; idx += G) {
                        double zeta = pp[0], eta = rr[0];
                        double dx = pp[1] - rr[1], dy = pp[2] - rr[2], dz = pp[3] - rr[3];
                        double ze = zeta + eta;
                        double theta = zeta * eta / ze;
                        double x = theta * (dx * dx + dy * dy + dz * dz);
                        double lrf = 1.;                       // long-range attenuation: theta -> theta * lrf
                        if (t.omega != 0.) {
                            lrf = t.omega * t.omega / (t.omega * t.omega + theta);
                            x *= lrf;
                        }
                        double v = rys_value<NROOTS>(idx, x, t.rys_cheb, t.rys_large);
                        if (idx < NROOTS) {
                            sl_rw[idx] = v * lrf / ze;
                        } else {
                            sl_rw[idx] = v * sqrt(lrf) * (sl[SLOT_FAC] * pp[4] * rr[4] / (zeta * eta * sqrt(ze)));
                        }
                    }
                    __syncthreads();
                    // ---------------- phase 2: VRR + TRR ----------------
                    for (int ra = tg; ra < 3 * NROOTS; ra += G) {
                        int root = ra / 3;
                        int axis = ra - root * 3;
                        double zeta = pp[0], eta = rr[0];
                        double rt_aa = sl_rw[root];
                        double rt_aij = rt_aa * eta;
                        double rt_akl = rt_aa * zeta;
                        double b00 = .5 * rt_aa;
                        double b10 = .5 / zeta * (1. - rt_aij);
                        double b01 = .5 / eta * (1. - rt_akl);
                        double P = pp[1 + axis], Q = rr[1 + axis];
                        double PQ = P - Q;
                        double c0p = (P - sl[SLOT_A + axis]) - rt_aij * PQ;
                        double cpq = (Q - blk[BLK_RSGEOM + 3 + axis]) + rt_akl * PQ;
                        double g[LRS + 1][LPQ + 1];
                        g[0][0] = (axis == 2) ? sl_rw[NROOTS + root] : 1.;
                        if (LPQ > 0) {
                            g[0][1] = c0p * g[0][0];
#pragma unroll
                            for (int e = 1; e < LPQ; ++e)
                                g[0][e + 1] = c0p * g[0][e] + e * b10 * g[0][e - 1];
                        }
                        if (LRS > 0) {
                            g[1][0] = cpq * g[0][0];
#pragma unroll
                            for (int e = 1; e <= LPQ; ++e)
                                g[1][e] = cpq * g[0][e] + e * b00 * g[0][e - 1];
#pragma unroll
                            for (int f = 1; f < LRS; ++f) {
                                g[f + 1][0] = cpq * g[f][0] + f * b01 * g[f - 1][0];
#pragma unroll
                                for (int e = 1; e <= LPQ; ++e)
                                    g[f + 1][e] = cpq * g[f][e] + f * b01 * g[f - 1][e] + e * b00 * g[f][e - 1];
                            }
                        }
                        double *o = sl_g2 + ra * G2SIZE;
#pragma unroll
                        for (int f = 0; f <= LRS; ++f)
#pragma unroll
                            for (int e = 0; e <= LPQ; ++e) o[f * (LPQ + 1) + e] = g[f][e];
                    }
                    __syncthreads();
                    // ---------------- phase 3: HRR (bra then ket), in registers ----------------
                    for (int r3 = tg; r3 < 3 * NROOTS * (LP + 1) * (LQ + 1); r3 += G) {
                        int ra = r3 / ((LP + 1) * (LQ + 1));
                        int ab2 = r3 - ra * ((LP + 1) * (LQ + 1));
                        int ea = ab2 / (LQ + 1);
                        int eb = ab2 - ea * (LQ + 1);
                        int axis = ra % 3;
                        double AB = sl[SLOT_AB + axis];
                        double CD = blk[BLK_RSGEOM + axis];
                        const double *in = sl_g2 + ra * G2SIZE + ea;
                        double v[LRS + 1];
#pragma unroll
                        for (int f = 0; f <= LRS; ++f) {
                            double u[LQ + 1];
#pragma unroll
                            for (int e = 0; e <= LQ; ++e) u[e] = (e <= eb) ? in[f * (LPQ + 1) + e] : 0.;
#pragma unroll
                            for (int k = 0; k < LQ; ++k) {
                                if (k < eb) {
#pragma unroll
                                    for (int e = 0; e < LQ - k; ++e) u[e] = u[e + 1] + AB * u[e];
                                }
                            }
                            v[f] = u[0];
                        }
                        double *o = sl_g + ra * GSIZE + ab2 * SUB;
#pragma unroll
                        for (int d = 0; d <= LS; ++d) {
#pragma unroll
                            for (int c = 0; c <= LR; ++c) o[c * (LS + 1) + d] = v[c];
                            if (d < LS) {
#pragma unroll
                                for (int c = 0; c < LRS - d; ++c) v[c] = v[c + 1] + CD * v[c];
                            }
                        }
                    }
                    __syncthreads();
                    // ---------------- phase 4: gout += gx * gy * gz ----------------
                    if (NCS == 1) {
#pragma unroll
                        for (int root = 0; root < NROOTS; ++root) {
                            const double *gr = sl_g + root * 3 * GSIZE;
                            double gx[SUB], gy[SUB], gz[SUB];
#pragma unroll
                            for (int i = 0; i < SUB; ++i) {
                                gx[i] = gr[gbx + i];
                                gy[i] = gr[gby + i];
                                gz[i] = gr[gbz + i];
                            }
#pragma unroll
                            for (int c = 0; c < NFR; ++c) {
#pragma unroll
                                for (int d = 0; d < NFS; ++d) {
                                    gout[c][d] += gx[cart_x(LR, c) * (LS + 1) + cart_x(LS, d)]
                                                * gy[cart_y(LR, c) * (LS + 1) + cart_y(LS, d)]
                                                * gz[cart_z(LR, c) * (LS + 1) + cart_z(LS, d)];
                                }
                            }
                        }
                    } else {
                        for (int root = 0; root < NROOTS; ++root) {
                            const double *gr = sl_g + root * 3 * GSIZE;
#pragma unroll
                            for (int n = 0; n < NC; ++n) {
                                double gx[LS + 1], gy[LS + 1], gz[LS + 1];
#pragma unroll
                                for (int i = 0; i <= LS; ++i) {
                                    gx[i] = gr[gbx + ox[n] + i];
                                    gy[i] = gr[gby + oy[n] + i];
                                    gz[i] = gr[gbz + oz[n] + i];
                                }
#pragma unroll
                                for (int d = 0; d < NFS; ++d) {
                                    gout[n][d] += gx[cart_x(LS, d)] * gy[cart_y(LS, d)] * gz[cart_z(LS, d)];
                                }
                            }
                        }
                    }
                } }

                // ---------------- digestion ----------------
                if (active) {
                    const size_t prow = (size_t)(p0 + a) * nao;
                    const size_t qrow = (size_t)(q0 + b) * nao;
#pragma unroll
                    for (int idm = 0; idm < NDM; ++idm) {
                        const double *dm = t.dm + idm * nao2;
                        if (do_j) {
                            double *vj = t.vj + idm * nao2;
                            double dpq = __ldg(dm + prow + q0 + b);
                            double s = 0.;
#pragma unroll
                            for (int n = 0; n < NC; ++n) {
                                if (c0 + n < NFR) {
#pragma unroll
                                    for (int d = 0; d < NFS; ++d) {
                                        s += gout[n][d] * drs[idm * NFRS + (c0 + n) * NFS + d];
                                        jacc[idm][n][d] += gout[n][d] * dpq;
                                    }
                                }
                            }
                            JK_ATOMIC_ADD(vj + prow + q0 + b, s);
                        }
                        if (do_k) {
                            double *vk = t.vk + idm * nao2;
                            double dqs[NFS], dps[NFS], kps[NFS], kqs[NFS];
#pragma unroll
                            for (int d = 0; d < NFS; ++d) {
                                dqs[d] = __ldg(dm + qrow + s0 + d);
                                dps[d] = __ldg(dm + prow + s0 + d);
                                kps[d] = 0.;
                                kqs[d] = 0.;
                            }
#pragma unroll
                            for (int n = 0; n < NC; ++n) {
                                if (c0 + n < NFR) {
                                    double dqr = __ldg(dm + qrow + r0 + c0 + n);
                                    double dpr = __ldg(dm + prow + r0 + c0 + n);
                                    double kpr = 0., kqr = 0.;
#pragma unroll
                                    for (int d = 0; d < NFS; ++d) {
                                        double gv = gout[n][d];
                                        kpr += gv * dqs[d];
                                        kqr += gv * dps[d];
                                        kps[d] += gv * dqr;
                                        kqs[d] += gv * dpr;
                                    }
                                    JK_ATOMIC_ADD(vk + prow + r0 + c0 + n, kpr);
                                    JK_ATOMIC_ADD(vk + qrow + r0 + c0 + n, kqr);
                                }
                            }
#pragma unroll
                            for (int d = 0; d < NFS; ++d) {
                                JK_ATOMIC_ADD(vk + prow + s0 + d, kps[d]);
                                JK_ATOMIC_ADD(vk + qrow + s0 + d, kqs[d]);
                            }
                        }
                    }
                }
            }   // steps

            // ---------------- keep the unprocessed remainder ----------------
            __syncthreads();
            const int rem = qn - nsteps * nq;
            if (tid == 0 && t.counter != nullptr) {
                atomicAdd(t.counter, (unsigned long long)(rem > 0 ? nsteps * nq : qn));
            }
            uint32_t keepv = 0;
            if (rem > 0 && tid < rem) keepv = queue[nsteps * nq + tid];
            __syncthreads();
            if (rem > 0 && tid < rem) queue[tid] = keepv;
            if (tid == 0) blk_ints[1] = rem > 0 ? rem : 0;
            __syncthreads();
            if (exhausted) break;
        }

        // ---------------- flush J_rs ----------------
        if (do_j) {
#pragma unroll
            for (int idm = 0; idm < NDM; ++idm) {
                double *vj = t.vj + idm * nao2;
#pragma unroll
                for (int n = 0; n < NC; ++n) {
                    if (c0 + n < NFR) {
#pragma unroll
                        for (int d = 0; d < NFS; ++d) {
                            double v = jacc[idm][n][d];
                            if (v != 0.) JK_ATOMIC_ADD(vj + (size_t)(r0 + c0 + n) * nao + s0 + d, v);
                        }
                    }
                }
            }
        }
    }
}
