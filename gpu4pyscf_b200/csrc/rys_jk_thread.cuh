// Register-resident variant of the J/K class kernel for small angular classes:
// ONE thread owns ONE shell quartet (pq|rs) -- roots, 2D recurrences, all
// nf_p*nf_q*nf_r*nf_s integrals and the six J/K digestions stay in registers;
// shared memory only holds the block-resident X pair and the survivor queue.
// Same screening, work distribution and output conventions as rys_jk_kernel
// (rys_jk_kernel.cuh); replaces the reference's unrolled rys_k_XXXX kernels
// (gpu4pyscf/lib/gvhf-rys/unrolled_rys_jk.cu) for classes with few components.
#pragma once
#include "common.cuh"
#include "rys_roots.cuh"
#include "rys_jk_kernel.cuh"

#ifndef JKT_THREADS
#define JKT_THREADS 128
#endif
// resident blocks per SM requested from the register allocator, by class size
__host__ __device__ constexpr int jkt_minblocks(int nf_total)
{
#ifdef JKT_MINBLOCKS_FIXED
    return JKT_MINBLOCKS_FIXED;
#endif
    // measured on Taxol (profiles/): 170 registers (3 blocks) beat 96/128 for everything but the (ss|..) classes;
    // from ~50 live doubles on, 255 registers (2 blocks) win: (20|30) with 70 spills 728 B per thread at 170
    // registers and stalls 71 % of the time on local-memory loads (profiles/ncu_r2/raw_r2j_20301111_metrics.txt)
    return nf_total <= 6 ? 5 : (nf_total <= 48 ? 3 : 2);
}

// A class whose integral block does not fit the register file CAN be evaluated in NPART passes over
// the Cartesian components c of shell r (NCP = NFR / NPART components per pass): roots and recurrences
// are recomputed per pass, the live integral block shrinks NPART times and the spills go away
// ((20|30): 264 -> 24 B of stack).  Measured on Taxol with JKT_PART_MIN = 48 (profiles/launch_table_r2l_threadsplit.txt):
// uncontracted launches gain 3-8 % ((20|30) 73 -> 69 ms, (11|30) 41 -> 38 ms), contracted launches lose up to
// 2x because the recurrences are redone per primitive quartet and pass ((11|11) p-contracted 29 -> 46 ms);
// +109 ms per build in total, so the split is compiled out by default.
#ifndef JKT_PART_MAX
#define JKT_PART_MAX 50          // integrals per pass
#endif
#ifndef JKT_PART_MIN
#define JKT_PART_MIN 1000000     // classes up to this many live doubles are not split (default: never split)
#endif
__host__ __device__ constexpr int jkt_nparts(int nfpq, int nfr, int nfs, int ndm)
{
    if (nfpq * nfr * nfs + nfr * nfs * ndm <= JKT_PART_MIN) return 1;
    for (int p = 2; p <= nfr; ++p)
        if (nfr % p == 0 && nfpq * (nfr / p) * nfs <= JKT_PART_MAX) return p;
    return 1;
}

// block-wide sum of per-thread values through shuffles + one smem stage
__device__ __forceinline__ double jkt_warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}

// shared-memory doubles of one warp: ints, X primitive pairs, D_rs block, 64-entry queue
__host__ __device__ inline int jkt_warp_doubles(int nrs, int ndm, int nfrs)
{
    return (2 + nrs * 5 + ndm * nfrs + 32 + 1) & ~1;
}

// One primitive quartet of (pq|rs): Rys roots and weights, VRR + TRR + both HRRs in
// registers, gout[((a*NFQ+b)*NFR+c)*NFS+d] += gx*gy*gz.  (zeta, P) is the Gaussian
// product of the (p,q) primitive pair with prefactor cpq, rr = {eta, Qx, Qy, Qz, coef}
// that of the (r,s) pair; A / C are the centres of p / r, AB = A - B, CD = C - D.
// Recurrences as in the reference (build_rys_gxyz.cuh:18-123), different data flow.
template <int LP, int LQ, int LR, int LS, int NCP = ncart(LR)>
__device__ __forceinline__ void jkt_prim_quartet(
    double (&gout)[ncart(LP) * ncart(LQ) * NCP * ncart(LS)],
    const double zeta, const double Px, const double Py, const double Pz, const double cpq,
    const double Ax, const double Ay, const double Az, const double ABx, const double ABy, const double ABz,
    const double *rr, const double Cx, const double Cy, const double Cz,
    const double CDx, const double CDy, const double CDz, const double omega, const double *rys_tab,
    const int c_lo = 0)
{
    // c_lo: first r component of this pass (a constant after unrolling the caller's pass loop);
    // gout holds the components c_lo .. c_lo + NCP - 1 of shell r
    typedef JKCfg<LP, LQ, LR, LS, 1> C;
    constexpr int NFP = C::NFP, NFQ = C::NFQ, NFS = C::NFS;
    constexpr int LPQ = C::LPQ, LRS = C::LRS, NROOTS = C::NROOTS;
    const double eta = rr[0];
    const double PQx = Px - rr[1], PQy = Py - rr[2], PQz = Pz - rr[3];
    const double ze = zeta + eta;
    const double zei = 1. / ze;
    const double theta = zeta * eta * zei;
    double x = theta * (PQx * PQx + PQy * PQy + PQz * PQz);
    double pref = cpq * rr[4] / (zeta * eta * sqrt(ze));
    double rscale = zei;
    if (omega != 0.) {                   // long-range erf(omega r)/r: theta -> theta * lrf
        const double lrf = omega * omega / (omega * omega + theta);
        x *= lrf;
        rscale *= lrf;
        pref *= sqrt(lrf);
    }
    const double hz = .5 / zeta, he = .5 / eta;
#pragma unroll
    for (int root = 0; root < NROOTS; ++root) {
        const double rt_aa = rys_value_s<NROOTS>(root, x, rys_tab) * rscale;
        const double wt = rys_value_s<NROOTS>(NROOTS + root, x, rys_tab) * pref;
        const double rt_aij = rt_aa * eta;
        const double rt_akl = rt_aa * zeta;
        const double b00 = .5 * rt_aa;
        const double b10 = hz * (1. - rt_aij);
        const double b01 = he * (1. - rt_akl);
        double g4[3][C::GSIZE];
#pragma unroll
        for (int axis = 0; axis < 3; ++axis) {
            const double PQ = axis == 0 ? PQx : (axis == 1 ? PQy : PQz);
            const double PA = axis == 0 ? Px - Ax : (axis == 1 ? Py - Ay : Pz - Az);
            const double QC = axis == 0 ? rr[1] - Cx : (axis == 1 ? rr[2] - Cy : rr[3] - Cz);
            const double AB = axis == 0 ? ABx : (axis == 1 ? ABy : ABz);
            const double CD = axis == 0 ? CDx : (axis == 1 ? CDy : CDz);
            const double c0p = PA - rt_aij * PQ;
            const double c0q = QC + rt_akl * PQ;
            double g[LRS + 1][LPQ + 1];
            g[0][0] = (axis == 2) ? wt : 1.;
            if (LPQ > 0) {
                g[0][1] = c0p * g[0][0];
#pragma unroll
                for (int e = 1; e < LPQ; ++e)
                    g[0][e + 1] = c0p * g[0][e] + e * b10 * g[0][e - 1];
            }
            if (LRS > 0) {
                g[1][0] = c0q * g[0][0];
#pragma unroll
                for (int e = 1; e <= LPQ; ++e)
                    g[1][e] = c0q * g[0][e] + e * b00 * g[0][e - 1];
#pragma unroll
                for (int f = 1; f < LRS; ++f) {
                    g[f + 1][0] = c0q * g[f][0] + f * b01 * g[f - 1][0];
#pragma unroll
                    for (int e = 1; e <= LPQ; ++e)
                        g[f + 1][e] = c0q * g[f][e] + f * b01 * g[f - 1][e] + e * b00 * g[f][e - 1];
                }
            }
            // bra HRR into u[b][a][f], then ket HRR
            double u[LQ + 1][LP + 1][LRS + 1];
#pragma unroll
            for (int f = 0; f <= LRS; ++f) {
                double w[LPQ + 1];
#pragma unroll
                for (int a = 0; a <= LPQ; ++a) w[a] = g[f][a];
#pragma unroll
                for (int b = 0; b <= LQ; ++b) {
#pragma unroll
                    for (int a = 0; a <= LP; ++a) u[b][a][f] = w[a];
                    if (b < LQ) {
#pragma unroll
                        for (int a = 0; a < LPQ - b; ++a) w[a] = w[a + 1] + AB * w[a];
                    }
                }
            }
#pragma unroll
            for (int a = 0; a <= LP; ++a)
#pragma unroll
                for (int b = 0; b <= LQ; ++b) {
                    double v[LRS + 1];
#pragma unroll
                    for (int f = 0; f <= LRS; ++f) v[f] = u[b][a][f];
#pragma unroll
                    for (int d = 0; d <= LS; ++d) {
#pragma unroll
                        for (int c = 0; c <= LR; ++c)
                            g4[axis][((a * (LQ + 1) + b) * (LR + 1) + c) * (LS + 1) + d] = v[c];
                        if (d < LS) {
#pragma unroll
                            for (int c = 0; c < LRS - d; ++c) v[c] = v[c + 1] + CD * v[c];
                        }
                    }
                }
        }
#pragma unroll
        for (int a = 0; a < NFP; ++a)
#pragma unroll
        for (int b = 0; b < NFQ; ++b)
#pragma unroll
        for (int cc = 0; cc < NCP; ++cc)
#pragma unroll
        for (int d = 0; d < NFS; ++d) {
            const int c = c_lo + cc;
            gout[((a * NFQ + b) * NCP + cc) * NFS + d] +=
                g4[0][((cart_x(LP, a) * (LQ + 1) + cart_x(LQ, b)) * (LR + 1) + cart_x(LR, c)) * (LS + 1) + cart_x(LS, d)]
              * g4[1][((cart_y(LP, a) * (LQ + 1) + cart_y(LQ, b)) * (LR + 1) + cart_y(LR, c)) * (LS + 1) + cart_y(LS, d)]
              * g4[2][((cart_z(LP, a) * (LQ + 1) + cart_z(LQ, b)) * (LR + 1) + cart_z(LR, c)) * (LS + 1) + cart_z(LS, d)];
        }
    }
}

template <int LP, int LQ, int LR, int LS, int NDM>
__global__ void
__launch_bounds__(JKT_THREADS,
                  jkt_minblocks(ncart(LP) * ncart(LQ) * ncart(LR) * ncart(LS)
                                / jkt_nparts(ncart(LP) * ncart(LQ), ncart(LR), ncart(LS), NDM)
                                + ncart(LR) * ncart(LS) * NDM
                                + (jkt_nparts(ncart(LP) * ncart(LQ), ncart(LR), ncart(LS), NDM) > 1
                                   ? NDM * (ncart(LP) * ncart(LQ) + (ncart(LP) + ncart(LQ)) * ncart(LS)) : 0)))
rys_jk_thread_kernel(JKTask t)
{
    constexpr int NPART = jkt_nparts(ncart(LP) * ncart(LQ), ncart(LR), ncart(LS), NDM);
    constexpr int NCP = ncart(LR) / NPART;
    typedef JKCfg<LP, LQ, LR, LS, 1> C;
    constexpr int NFP = C::NFP, NFQ = C::NFQ, NFR = C::NFR, NFS = C::NFS;
    constexpr int NFPQ = C::NFPQ, NFRS = C::NFRS;
    constexpr int LPQ = C::LPQ, LRS = C::LRS, NROOTS = C::NROOTS;

#ifdef CPU_EMUL
    double *smem = (double *)emul_dynamic_smem;
#else
    extern __shared__ double smem[];
#endif
    const int tid = threadIdx.x;
    const int NT = blockDim.x;
    const int npq = t.nprim_p * t.nprim_q;
    const int nrs = t.nprim_r * t.nprim_s;
    const int nterm = t.ratio2 != 0. ? 2 : 1;
    const int nbas = t.nbas;
    const int nao = t.nao;
    const size_t nao2 = (size_t)nao * nao;
    const bool do_j = t.vj != nullptr;
    const bool do_k = t.vk != nullptr;

    // every warp is an independent worker: own X pair, own survivor queue
    const int lane = tid & 31;
    const int warp = tid >> 5;
    double *rys_tab = smem;                          // Rys tables of this NROOTS (rys_roots.cuh)
    double *wsm = smem + RysSmem<NROOTS>::DOUBLES + warp * jkt_warp_doubles(nrs, NDM, NFRS);
    int *w_ints = (int *)wsm;                        // [0] ix
    double *rsprim = wsm + 2;                        // nrs * 5: eta, Qx, Qy, Qz, coef
    double *drs = rsprim + nrs * 5;
    uint32_t *queue = (uint32_t *)(drs + NDM * NFRS);  // 64 entries
    rys_stage_tables<NROOTS>(rys_tab, t.rys_cheb, t.rys_large, tid, NT);
    __syncthreads();

    const int *bas = t.bas;
    const double *env = t.env;
    const int *ao_loc = t.ao_loc;

    while (true) {
        __syncwarp();
        if (lane == 0) w_ints[0] = atomicAdd(t.head, 1) * t.world + t.rank;
        __syncwarp();
        // work item = (X pair ix, chunk of the q-sorted Y list), chunk-major
        const int item = w_ints[0];
        if (item >= t.npair_x * t.nychunk) break;
        const int chunk = item / t.npair_x;
        const int ix = item - chunk * t.npair_x;
        const float thr = t.cutoff - t.q_x[ix];
        int ypos = chunk * t.ychunk;
        if (t.dedupe == 1 && ypos < ix) ypos = ix;
        const int y_end = min(t.npair_y, (chunk + 1) * t.ychunk);
        if (ypos >= y_end || t.q_y[ypos] + t.dm_penalty + t.q_margin < thr) continue;   // dead chunk
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
            double rr2 = CDx * CDx + CDy * CDy + CDz * CDz;
            for (int i = lane; i < nrs; i += 32) {
                int ir = i / t.nprim_s;
                int is = i - ir * t.nprim_s;
                double ar = er[ir], as = es[is];
                double eta = ar + as;
                double f = as / eta;
                double *o = rsprim + i * 5;
                o[0] = eta;
                o[1] = Cx - f * CDx;
                o[2] = Cy - f * CDy;
                o[3] = Cz - f * CDz;
                o[4] = cr[ir] * cs[is] * exp(-ar * f * rr2);
            }
            for (int i = lane; i < NDM * NFRS; i += 32) {
                int idm = i / NFRS;
                int cd = i - idm * NFRS;
                int c = cd / NFS;
                int d = cd - c * NFS;
                drs[i] = __ldg(t.dm + idm * nao2 + (size_t)(r0 + c) * nao + s0 + d);
            }
        }
        // 16-byte loads of D rows when the r / s blocks start on an even AO index
        const bool vec_r = NFR > 1 && ((r0 | nao) & 1) == 0;
        const bool vec_s = NFS > 1 && ((s0 | nao) & 1) == 0;
        double jacc[NDM][NFRS];
#pragma unroll
        for (int idm = 0; idm < NDM; ++idm)
#pragma unroll
            for (int n = 0; n < NFRS; ++n) jacc[idm][n] = 0.;

        const float *dc_r = t.dm_cond + (size_t)rsh * nbas;
        const float *dc_s = t.dm_cond + (size_t)ssh * nbas;
        const float d_rs = dc_r[ssh];
        int qn = 0;
        bool exhausted = false;
        __syncwarp();

        while (true) {
            // ---------------- screen Y pairs into the warp's queue (ballot compaction) ----------------
            while (!exhausted && qn < 32) {
                const int y = ypos + lane;
                bool keep = false, stop = false;
                if (y >= y_end) {
                    stop = true;
                } else {
                    float qy = t.q_y[y];
                    if (qy + t.dm_penalty + t.q_margin < thr) {
                        stop = true;                  // sorted: nothing further survives
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
                            // short-range exchange filter; the X-pair data are re-read there so that the
                            // common path carries no extra live registers
                            if (t.sr_s_x != nullptr && keep) keep = sr_keep(t, ix, y, dmax);
                        }
                    }
                }
                const unsigned m = __ballot_sync(0xffffffffu, keep);
                if (keep) queue[qn + __popc(m & ((1u << lane) - 1u))] = (uint32_t)y;
                qn += __popc(m);
                exhausted = __any_sync(0xffffffffu, stop);
                ypos += 32;
                __syncwarp();
            }
            if (qn == 0) break;
            const int nb = qn < 32 ? qn : 32;

            if (lane < nb) {
                const int iy = queue[lane];
                // Gaussian products of the Y pair: from its precomputed record when the
                // caller provides one, else from bas/env (reference-signature entry)
                const double *rec = t.rec_y != nullptr ? t.rec_y + (size_t)iy * t.rec_stride : nullptr;
                int p0, q0;
                double Ax, Ay, Az, ABx, ABy, ABz, fac = PI_FAC * t.scale, rr2 = 0.;
                const double *ep = nullptr, *eq = nullptr, *cp = nullptr, *cq = nullptr;
                if (rec != nullptr) {
                    const double2 h0 = *(const double2 *)(rec + 0);
                    const double2 h1 = *(const double2 *)(rec + 2);
                    const double2 h2 = *(const double2 *)(rec + 4);
                    const double2 h3 = *(const double2 *)(rec + 6);
                    ABx = h0.x; ABy = h0.y; ABz = h1.x;
                    Ax = h1.y; Ay = h2.x; Az = h2.y;
                    p0 = ((const int *)&h3.x)[0];
                    q0 = ((const int *)&h3.x)[1];
                    fac *= h3.y;                               // 1/2 when psh == qsh
                    if (t.dedupe == 1 && iy == ix) fac *= .5;  // same list: (pq) == (rs)
                } else {
                    const uint32_t pair_pq = t.pair_y[iy];
                    const int psh = pair_pq / nbas;
                    const int qsh = pair_pq - psh * nbas;
                    p0 = ao_loc[psh];
                    q0 = ao_loc[qsh];
                    const double *Rp = env + bas[psh * BAS_SLOTS + PTR_BAS_COORD];
                    const double *Rq = env + bas[qsh * BAS_SLOTS + PTR_BAS_COORD];
                    ep = env + bas[psh * BAS_SLOTS + PTR_EXP];
                    eq = env + bas[qsh * BAS_SLOTS + PTR_EXP];
                    cp = env + bas[psh * BAS_SLOTS + PTR_COEFF];
                    cq = env + bas[qsh * BAS_SLOTS + PTR_COEFF];
                    Ax = Rp[0]; Ay = Rp[1]; Az = Rp[2];
                    ABx = Ax - Rq[0]; ABy = Ay - Rq[1]; ABz = Az - Rq[2];
                    rr2 = ABx * ABx + ABy * ABy + ABz * ABz;
                    if (psh == qsh) fac *= .5;
                    if (pair_pq == pair_rs) fac *= .5;
                }
                if (rsh == ssh) fac *= .5;

                // sums that run over all r components: complete after the last pass
                double jpq[NDM][NFPQ], kps[NDM][NFP][NFS], kqs[NDM][NFQ][NFS];
#pragma unroll
                for (int idm = 0; idm < NDM; ++idm) {
#pragma unroll
                    for (int n = 0; n < NFPQ; ++n) jpq[idm][n] = 0.;
#pragma unroll
                    for (int n = 0; n < NFP * NFS; ++n) kps[idm][n / NFS][n % NFS] = 0.;
#pragma unroll
                    for (int n = 0; n < NFQ * NFS; ++n) kqs[idm][n / NFS][n % NFS] = 0.;
                }

#pragma unroll
                for (int part = 0; part < NPART; ++part) {
                const int c_lo = part * NCP;
                double gout[NFPQ * NCP * NFS];
#pragma unroll
                for (int n = 0; n < NFPQ * NCP * NFS; ++n) gout[n] = 0.;

                for (int ipq = 0; ipq < npq; ++ipq) {
                    double zeta, Px, Py, Pz, cpq;
                    if (rec != nullptr) {
                        const double *pr = rec + PAIR_REC_HDR + ipq * 5;
                        zeta = pr[0]; Px = pr[1]; Py = pr[2]; Pz = pr[3];
                        cpq = fac * pr[4];
                    } else {
                        const int ip = ipq / t.nprim_q;
                        const int iq = ipq - ip * t.nprim_q;
                        const double ap = ep[ip], aq = eq[iq];
                        zeta = ap + aq;
                        const double fq = aq / zeta;
                        Px = Ax - fq * ABx; Py = Ay - fq * ABy; Pz = Az - fq * ABz;
                        cpq = fac * cp[ip] * cq[iq] * exp(-ap * fq * rr2);
                    }
                    // operator terms (JKTask::ratio2) share ONE call site: measured in registers and stack, this
                    // form costs the single-term path nothing (a second inlined call or a per-term table did)
                    for (int it = 0; it < nrs * nterm; ++it) {
                        const int term = it / nrs;
                        const int irs = it - term * nrs;
                        jkt_prim_quartet<LP, LQ, LR, LS, NCP>(gout, zeta, Px, Py, Pz, term ? cpq * t.ratio2 : cpq,
                                                              Ax, Ay, Az, ABx, ABy, ABz,
                                                              rsprim + irs * 5, Cx, Cy, Cz, CDx, CDy, CDz,
                                                              term ? t.omega2 : t.omega, rys_tab, c_lo);
                    }
                }

                // ---------------- digestion of this pass (thread-local) ----------------
#pragma unroll
                for (int idm = 0; idm < NDM; ++idm) {
                    const double *dm = t.dm + idm * nao2;
                    if (do_j) {
#pragma unroll
                        for (int a = 0; a < NFP; ++a)
#pragma unroll
                        for (int b = 0; b < NFQ; ++b) {
                            const double dpq = __ldg(dm + (size_t)(p0 + a) * nao + q0 + b);
                            double s = 0.;
#pragma unroll
                            for (int cd = 0; cd < NCP * NFS; ++cd) {
                                const double gv = gout[(a * NFQ + b) * NCP * NFS + cd];
                                s += gv * drs[idm * NFRS + c_lo * NFS + cd];
                                jacc[idm][c_lo * NFS + cd] += gv * dpq;
                            }
                            jpq[idm][a * NFQ + b] += s;
                        }
                    }
                    if (do_k) {
                        double *vk = t.vk + idm * nao2;
                        // K_pr (r components of this pass), K_ps (partial), rows of p
                        double dqr[NFQ][NCP], dqs[NFQ][NFS];
#pragma unroll
                        for (int b = 0; b < NFQ; ++b) {
                            if (NPART == 1) {
                                ld_row<NCP>(dqr[b], dm + (size_t)(q0 + b) * nao + r0, vec_r);
                            } else {
#pragma unroll
                                for (int cc = 0; cc < NCP; ++cc) dqr[b][cc] = __ldg(dm + (size_t)(q0 + b) * nao + r0 + c_lo + cc);
                            }
                            ld_row<NFS>(dqs[b], dm + (size_t)(q0 + b) * nao + s0, vec_s);
                        }
#pragma unroll
                        for (int a = 0; a < NFP; ++a) {
                            double kpr[NCP];
#pragma unroll
                            for (int cc = 0; cc < NCP; ++cc) kpr[cc] = 0.;
#pragma unroll
                            for (int b = 0; b < NFQ; ++b)
#pragma unroll
                            for (int cc = 0; cc < NCP; ++cc)
#pragma unroll
                            for (int d = 0; d < NFS; ++d) {
                                const double gv = gout[((a * NFQ + b) * NCP + cc) * NFS + d];
                                kpr[cc] += gv * dqs[b][d];
                                kps[idm][a][d] += gv * dqr[b][cc];
                            }
#pragma unroll
                            for (int cc = 0; cc < NCP; ++cc)
                                JK_ATOMIC_ADD(vk + (size_t)(p0 + a) * nao + r0 + c_lo + cc, kpr[cc]);
                        }
                        // K_qr (r components of this pass), K_qs (partial), rows of q
                        double dpr[NFP][NCP], dps[NFP][NFS];
#pragma unroll
                        for (int a = 0; a < NFP; ++a) {
                            if (NPART == 1) {
                                ld_row<NCP>(dpr[a], dm + (size_t)(p0 + a) * nao + r0, vec_r);
                            } else {
#pragma unroll
                                for (int cc = 0; cc < NCP; ++cc) dpr[a][cc] = __ldg(dm + (size_t)(p0 + a) * nao + r0 + c_lo + cc);
                            }
                            ld_row<NFS>(dps[a], dm + (size_t)(p0 + a) * nao + s0, vec_s);
                        }
#pragma unroll
                        for (int b = 0; b < NFQ; ++b) {
                            double kqr[NCP];
#pragma unroll
                            for (int cc = 0; cc < NCP; ++cc) kqr[cc] = 0.;
#pragma unroll
                            for (int a = 0; a < NFP; ++a)
#pragma unroll
                            for (int cc = 0; cc < NCP; ++cc)
#pragma unroll
                            for (int d = 0; d < NFS; ++d) {
                                const double gv = gout[((a * NFQ + b) * NCP + cc) * NFS + d];
                                kqr[cc] += gv * dps[a][d];
                                kqs[idm][b][d] += gv * dpr[a][cc];
                            }
#pragma unroll
                            for (int cc = 0; cc < NCP; ++cc)
                                JK_ATOMIC_ADD(vk + (size_t)(q0 + b) * nao + r0 + c_lo + cc, kqr[cc]);
                        }
                    }
                }
                }   // passes over the r components

                // ---------------- the sums over all r components leave the thread ----------------
#pragma unroll
                for (int idm = 0; idm < NDM; ++idm) {
                    if (do_j) {
                        double *vj = t.vj + idm * nao2;
#pragma unroll
                        for (int a = 0; a < NFP; ++a)
#pragma unroll
                        for (int b = 0; b < NFQ; ++b)
                            JK_ATOMIC_ADD(vj + (size_t)(p0 + a) * nao + q0 + b, jpq[idm][a * NFQ + b]);
                    }
                    if (do_k) {
                        double *vk = t.vk + idm * nao2;
#pragma unroll
                        for (int a = 0; a < NFP; ++a)
#pragma unroll
                        for (int d = 0; d < NFS; ++d) JK_ATOMIC_ADD(vk + (size_t)(p0 + a) * nao + s0 + d, kps[idm][a][d]);
#pragma unroll
                        for (int b = 0; b < NFQ; ++b)
#pragma unroll
                        for (int d = 0; d < NFS; ++d) JK_ATOMIC_ADD(vk + (size_t)(q0 + b) * nao + s0 + d, kqs[idm][b][d]);
                    }
                }
            }   // lane < nb

            // ---------------- keep the unprocessed remainder of the queue ----------------
            __syncwarp();
            if (lane == 0 && t.counter != nullptr) atomicAdd(t.counter, (unsigned long long)nb);
            const int rem = qn - nb;
            uint32_t keepv = 0;
            if (lane < rem) keepv = queue[nb + lane];
            __syncwarp();
            if (lane < rem) queue[lane] = keepv;
            qn = rem;
            __syncwarp();
            if (exhausted && qn == 0) break;
        }

        // ---------------- flush J_rs: warp reduction, one atomic per warp ----------------
        if (do_j) {
#pragma unroll
            for (int idm = 0; idm < NDM; ++idm) {
                double *vj = t.vj + idm * nao2;
#pragma unroll
                for (int cd = 0; cd < NFRS; ++cd) {
                    double v = jkt_warp_sum(jacc[idm][cd]);
                    if (lane == 0 && v != 0.) JK_ATOMIC_ADD(vj + (size_t)(r0 + cd / NFS) * nao + s0 + cd % NFS, v);
                }
            }
        }
    }
}
