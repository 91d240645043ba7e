// Warp-team variant of the direct-SCF J/K class kernel (sm_100a, FP64).
//
// Same arithmetic and output conventions as rys_jk_kernel (rys_jk_kernel.cuh),
// different execution model: the unit of work is a TEAM of TW warps that takes
// (X = (r,s) shell pair, chunk of the q-ordered Y = (p,q) list) work items from a
// device counter, screens its chunk itself (ballot compaction into a private
// queue) and evaluates NSW quartets at once (TW == 1: lanes = NSW x G, with
// G = nf_p*nf_q*NCS lanes per quartet, or nf_p*NCS in the b-loop variant;
// TW > 1: one quartet over TW warps).  Teams never synchronise with each other --
// the only barriers are __syncwarp (TW == 1) or a named barrier over the team's
// warps -- so the dependent chains of the root and recurrence phases of one team
// overlap the FMA-dense contraction/digestion phases of the other resident teams.
//
// Per pass a team holds NSW*K "g-sets" (roots, weights, 2D integrals of one
// primitive quartet) in shared memory:
//   uncontracted launch (1 primitive quartet): K different quartets per lane slot
//   contracted launch: K primitive quartets of the same quartet
// and runs
//   phase 1  one root or weight per lane (Rys series from the shared-memory tables)
//   phase 2  one (g-set, root, axis) per lane: VRR + TRR + bra HRR in registers,
//            result to shared memory in 16-byte segments
//   phase 3  per lane: ket HRR in registers, gout += gx*gy*gz on its register tile
//   phase 4  J/K digestion (K rows shuffle-reduced over the lanes of a quartet,
//            then FP64 atomics; J_rs accumulated in registers over the work item)
//
// Replaces (with rys_jk_thread_kernel for the light classes and rys_jk_kernel for
// classes needing more than 256 lanes) the reference's rys_jk_kernel /
// rys_k_kernel / unrolled kernels (gpu4pyscf/lib/gvhf-rys/rys_contract_jk.cu:33-333,
// unrolled_rys_jk.cu) and its task generator _fill_vjk_tasks
// (create_tasks.cu:113-180).
#pragma once
#include "common.cuh"
#include "rys_roots.cuh"
#include "rys_jk_kernel.cuh"

// BL = 1 ("b loop", uncontracted launches only): a lane owns a Cartesian component a of
// shell p and loops over the nf_q components b itself, so a quartet needs nf_p*NCS lanes
// instead of nf_p*nf_q*NCS and a warp carries nf_q times more quartets.
#ifndef JKW_SMEM_RED
#define JKW_SMEM_RED 1
#endif
template <int LP, int LQ, int LR, int LS, int NCS, int BL = 0>
struct JKWCfg : JKCfg<LP, LQ, LR, LS, NCS> {
    typedef JKCfg<LP, LQ, LR, LS, NCS> B;
    static constexpr int G = (BL ? B::NFP : B::NFPQ) * NCS;   // lanes per quartet
    static constexpr int NBL = BL ? B::NFQ : 1;               // b components per lane
    static constexpr int TW = (G + 31) / 32;             // warps per team
    static constexpr int TEAM = TW * 32;
    static constexpr int NSW = TW == 1 ? 32 / G : 1;     // quartets per team in the contraction phase
    // strides (in doubles) are even with an odd half: segments are 16-byte aligned
    // for 128-bit shared-memory accesses and consecutive segments / (root,axis)
    // blocks / g-sets start in different 16-byte bank groups
    static constexpr int oddhalf(int n) { return (n + (n & 1)) + ((((n + (n & 1)) / 2) % 2 == 0) ? 2 : 0); }
    // NCS == 1: a segment holds the bra-transferred g(ea,eb; f), f = 0..LR+LS, and the
    // ket transfer runs in registers in the contraction phase (fewer shared-memory
    // bytes per FMA); NCS > 1: a segment holds the final (c,d) block
    static constexpr int SEG = NCS == 1 ? B::LRS + 1 : B::SUBSIZE;
    static constexpr int SUBP = oddhalf(SEG);
    static constexpr int GSIZEP = oddhalf((LP + 1) * (LQ + 1) * SUBP);
    static constexpr int GSET0 = oddhalf(2 * B::NROOTS + 3 * B::NROOTS * GSIZEP);
    // Bank conflicts of the 16-byte accesses to the g-sets for a g-set stride of `gset` doubles:
    // (a) contraction loads -- lane (quartet slot, a, b) reads the segment of its (ea, eb) on every
    // axis; (b) stores of the merged root/recurrence phase -- lane (g-set, root).  A quarter warp
    // (8 lanes) is served in one wavefront when its distinct addresses fall in distinct 16-byte
    // bank groups; returns the excess wavefronts.
    static constexpr int gset_conflicts(int gset)
    {
        int excess = 0;
        for (int mode = 0; mode < 4; ++mode) {                 // 0..2: loads on axis x, y, z; 3: stores
            for (int qw = 0; qw < 4; ++qw) {
                int addr[8] = {0}, n = 0;
                for (int l = qw * 8; l < qw * 8 + 8; ++l) {
                    int off = -1;
                    if (mode < 3) {
                        const int s4 = TW == 1 ? l / G : 0;
                        if (TW == 1 && s4 >= NSW) continue;
                        const int tg = TW == 1 ? l - s4 * G : l;
                        if (tg >= G) continue;
                        const int ab = tg / NCS;
                        const int a = BL ? ab : ab / B::NFQ;
                        const int b = BL ? 0 : ab - a * B::NFQ;
                        const int ea = mode == 0 ? cart_x(LP, a) : (mode == 1 ? cart_y(LP, a) : cart_z(LP, a));
                        const int eb = mode == 0 ? cart_x(LQ, b) : (mode == 1 ? cart_y(LQ, b) : cart_z(LQ, b));
                        off = s4 * gset + 2 * B::NROOTS + mode * GSIZEP + (ea * (LQ + 1) + eb) * SUBP;
                    } else {
                        const int g = l / B::NROOTS, root = l - g * B::NROOTS;
                        if (g >= NSW * 4) continue;
                        off = g * gset + 2 * B::NROOTS + root * 3 * GSIZEP;
                    }
                    bool dup = false;
                    for (int i = 0; i < n; ++i) dup = dup || addr[i] == off / 2;
                    if (!dup) addr[n++] = off / 2;
                }
                int cnt[8] = {0}, worst = 0;
                for (int i = 0; i < n; ++i) {
                    const int grp = addr[i] & 7;
                    cnt[grp]++;
                    worst = cnt[grp] > worst ? cnt[grp] : worst;
                }
                excess += worst > 1 ? worst - 1 : 0;
            }
        }
        return excess;
    }
    // smallest padding (< 16 doubles) of the g-set stride with the fewest conflicts
    static constexpr int best_gset()
    {
        int best = GSET0, cost = gset_conflicts(GSET0);
        for (int pad = 2; pad < 16; pad += 2) {
            const int c = gset_conflicts(GSET0 + pad);
            if (c < cost) {
                cost = c;
                best = GSET0 + pad;
            }
        }
        return best;
    }
    static constexpr int GSET = best_gset();
    // registers: gout + J_rs accumulators + one root's gx/gy/gz tile
    static constexpr int TILE = B::NC * B::NFS;
    // row-stream (uncontracted) kernels: quartets per lane slot such that the (g-set, root) lanes of
    // the merged root/recurrence phase fill the team, within 16 KB of g-sets per team
    static constexpr int krow()
    {
        int k = TEAM / (NSW * B::NROOTS);
        k = k > 4 ? 4 : (k < 1 ? 1 : k);
        while (k > 1 && NSW * k * GSET * 8 > 16384) --k;
        return k;
    }
    static constexpr int KROW = krow();
    // K rows that several lanes of a single-warp team hold partial sums of -- K_q* over the p
    // components a, K_p* over the q components b -- are summed through a shared-memory transpose
    // (REDV values per lane and set, one 32-lane row per value): REDV stores and NFP (NFQ) loads per
    // quartet instead of REDV * log2(NFP) shuffle + select steps
    // The scratch (REDV rows of REDS doubles: odd stride, so that the lanes of one column fall in
    // different banks) is the g-set area of the quartet batch that has just been contracted; only
    // the b-loop variant, which re-reads its g-sets for the next b, and classes whose batch is
    // smaller than the scratch get a dedicated buffer.
    static constexpr int REDV = B::NC + B::NFS;
    static constexpr int REDS = 33;
    // Measured on Taxol (profiles/launch_table_r2h_smemred.txt vs launch_table_r2e.txt): the transpose
    // wins where the shuffle tree is deep and the tile large -- lanes (d,p), (f,s), (f,p) against
    // (d,d)/(f,d)/(f,f) tiles: -4..-10 % -- and loses 5..20 % for p-shell lanes (two shuffle steps).
    static constexpr bool RED_WINS = B::NFP >= 10 || (B::NFP >= 6 && B::NFR * B::NFS >= 36);
    static constexpr bool RED_Q = JKW_SMEM_RED && TW == 1 && B::NFP > 1 && RED_WINS;
    __host__ __device__ static constexpr bool red_p(int row) { return RED_Q && !row && !BL && B::NFQ > 1; }
    __host__ __device__ static constexpr int redd(int row)
    {
        (void)row;
        return (RED_Q && (BL || NSW * GSET < REDV * REDS)) ? ((REDV * REDS + 1) & ~1) : 0;
    }
};

#define JKW_SLOT_HDR 8       // AB[3], A[3], fac, (p0,q0)

// shared-memory doubles of one team (used by the kernel and by the launcher)
// rowd: doubles of the row buffers of the row-stream variant (D_p* block + K_p* accumulators), else 0
// redd: doubles of the reduction scratch (JKWCfg::redd), even
__host__ __device__ inline int jkw_team_doubles(int team_threads, int nsw, int K, int npq, int nrs,
                                                int ndm, int nfrs, int gset, int rowd = 0, int redd = 0,
                                                int nterm = 1)
{
    int ngs = nsw * K;
    int nqs = (npq * nrs * nterm == 1) ? ngs : nsw;
    int qcap = nqs + team_threads;
    int n = 8 + nrs * 5 + ndm * nfrs + rowd + (qcap + 1) / 2 + nqs * (JKW_SLOT_HDR + 5 * npq);
    n = (n + 1) & ~1;                    // g-sets start 16-byte aligned
    return n + redd + ngs * gset;
}

template <int TW>
__device__ __forceinline__ void jkw_team_sync(int team)
{
#ifdef CPU_EMUL
    if (TW == 1) __syncwarp();
    else emul_barrier(1 + team, TW * 32);
#else
    if (TW == 1) __syncwarp();
    else asm volatile("bar.sync %0, %1;" :: "r"(team + 1), "r"(TW * 32) : "memory");
#endif
}

// 16-byte shared-memory vector access of an N-double segment (p is 16-byte aligned)
template <int N>
__device__ __forceinline__ void jkw_ld_seg(double (&dst)[N], const double *p)
{
#pragma unroll
    for (int i = 0; i + 1 < N; i += 2) {
        double2 v = *(const double2 *)(p + i);
        dst[i] = v.x;
        dst[i + 1] = v.y;
    }
    if (N & 1) dst[N - 1] = p[N - 1];
}
template <int N>
__device__ __forceinline__ void jkw_st_seg(double *p, const double (&src)[N])
{
#pragma unroll
    for (int i = 0; i + 1 < N; i += 2) {
        double2 v;
        v.x = src[i];
        v.y = src[i + 1];
        *(double2 *)(p + i) = v;
    }
    if (N & 1) p[N - 1] = src[N - 1];
}

#ifndef JKW_MAXTHREADS
#define JKW_MAXTHREADS 256
#endif
// resident blocks per SM asked from the register allocator, by register-tile size
// Block size and resident blocks by the size d of the per-lane register tile (doubles that
// stay live across a work item: gout + J_rs accumulators).  Measured on Taxol (profiles/):
//   d <= 24      : 256 threads, 3 blocks (85 registers, 24 warps)
//   24 < d <= 44 : 128 threads, 3 blocks (170 registers, 12 warps) -- X = (d,p) tiles
//   d > 44       : 256 threads, 1 block  (255 registers, 8 warps)
__host__ __device__ constexpr int jkw_max_threads(int d) { return (d > 24 && d <= 44) ? 128 : JKW_MAXTHREADS; }
__host__ __device__ constexpr int jkw_block_threads(int team, int d)
{
    return team > jkw_max_threads(d) ? team : (jkw_max_threads(d) / team) * team;
}
// Launch bounds.  The mid-size tiles (170 registers) may also run as ONE block of 384 threads per SM
// instead of three blocks of 128: same register budget, but the Rys table slice is staged once per SM,
// which lets 12 warps fit where shared memory admits only two 128-thread blocks (the launcher decides).
// Only for the single-warp-team row-stream kernels: elsewhere the 384-thread bound made ptxas spill
// more at the same register count ((21|33) rows: 160 -> 248 B of stack) and no launch needed it.
__host__ __device__ constexpr bool jkw_big_ok(int team, int d, int row) { return row && team == 32 && d > 24 && d <= 44; }
__host__ __device__ constexpr int jkw_bound_threads(int team, int d, int row)
{
    return jkw_big_ok(team, d, row) ? 384 : jkw_block_threads(team, d);
}
__host__ __device__ constexpr int jkw_minblocks(int tile, int ndm, int sub, int team);
__host__ __device__ constexpr int jkw_bound_blocks(int tile, int ndm, int sub, int team, int row)
{
    return jkw_big_ok(team, tile * (1 + ndm), row) ? 1 : jkw_minblocks(tile, ndm, sub, team);
}
__host__ __device__ constexpr int jkw_minblocks(int tile, int ndm, int sub, int team)
{
    int d = tile * (1 + ndm);
    int threads = jkw_block_threads(team, d);
    int warps = d <= 24 ? 24 : (d <= 44 ? 12 : 8);
    (void)sub;
    int nb = warps * 32 / threads;
    return nb < 1 ? 1 : nb;
}

// ROW = 1: row-stream variant.  The Y list is grouped in rows of one shell p (JKTask::row_ptr /
// row_p); a team owns (X pair, block of rows) and walks the rows one at a time, so p, r, s are
// team-uniform: D_pr / D_ps are staged in shared memory once per row, K_pr / K_ps accumulate in
// registers over the row and leave the team through a shared-memory panel at the row end, and
// the q-dependent blocks are addressed with q as the column index (dmT / vkT), see rys_jk_rowt.cuh.
template <int LP, int LQ, int LR, int LS, int NCS, int NDM, int BL = 0, int ROW = 0>
__global__ void
__launch_bounds__(jkw_bound_threads(JKWCfg<LP, LQ, LR, LS, NCS, BL>::TEAM,
                                    JKWCfg<LP, LQ, LR, LS, NCS, BL>::TILE * (1 + NDM), ROW),
                  jkw_bound_blocks(JKWCfg<LP, LQ, LR, LS, NCS, BL>::TILE, NDM,
                                NCS == 1 ? JKWCfg<LP, LQ, LR, LS, NCS, BL>::SUBSIZE : LS + 1,
                                JKWCfg<LP, LQ, LR, LS, NCS, BL>::TEAM, ROW))
rys_jk_warp_kernel(JKTask t)
{
    typedef JKWCfg<LP, LQ, LR, LS, NCS, BL> C;
    constexpr int NBL = C::NBL;
    constexpr int NBLANE = BL ? 1 : C::NFQ;                // q components spread over lanes
    constexpr int NFQ = C::NFQ, NFR = C::NFR, NFS = C::NFS, NFRS = C::NFRS;
    constexpr int LPQ = C::LPQ, LRS = C::LRS, NROOTS = C::NROOTS;
    constexpr int NC = C::NC, G = C::G;
    constexpr int SUB = C::SUBSIZE, SUBP = C::SUBP, GSIZEP = C::GSIZEP;
    constexpr int TW = C::TW, TEAM = C::TEAM, NSW = C::NSW, GSET = C::GSET;

#ifdef CPU_EMUL
    double *smem = (double *)emul_dynamic_smem;
#else
    extern __shared__ double smem[];
#endif
    const int tid = threadIdx.x;
    const int team = tid / TEAM;
    const int lt = tid - team * TEAM;
    const int lane = tid & 31;

    double *rys_tab = smem;
    rys_stage_tables<NROOTS>(rys_tab, t.rys_cheb, t.rys_large, tid, blockDim.x);
    __syncthreads();

    // ROW kernels serve uncontracted launches only: every shape below is a compile-time constant
    constexpr bool SNG = ROW != 0;
    const int npq = SNG ? 1 : t.nprim_p * t.nprim_q;
    const int nrs = SNG ? 1 : t.nprim_r * t.nprim_s;
    // a second operator term (JKTask::ratio2) doubles the primitive-quartet list: entry pi of
    // [nprimq, 2 nprimq) is primitive quartet pi - nprimq evaluated with omega2 and weight * ratio2
    const int nterm = (!SNG && t.ratio2 != 0.) ? 2 : 1;
    const int nprimq = npq * nrs * nterm;
    const bool single = SNG || nprimq == 1;
    const int K = SNG ? C::KROW : t.nq;      // g-sets per lane slot
    const int ngs = NSW * K;
    const int nqs = single ? ngs : NSW;      // quartets per batch
    const int qcap = nqs + TEAM;
    const int slot_d = JKW_SLOT_HDR + 5 * npq;
    const int npass = single ? 1 : (nprimq + K - 1) / K;

    double *tm = smem + RysSmem<NROOTS>::DOUBLES + team * t.slot_doubles;
    int *tm_ints = (int *)tm;                // [0..1] work item (64 bit when ROW), [2] qn, [3] exhausted
    double *rsgeom = tm + 2;                 // CD[3], C[3]
    double *rsprim = tm + 8;                 // nrs * 5: eta, Qx, Qy, Qz, coef
    double *drs = rsprim + nrs * 5;
    constexpr int NKP = C::NFP * (NFR + NFS);              // K_pr | K_ps values of one density matrix
    constexpr int ROWD = ROW ? 2 * NDM * NKP : 0;
    double *dpx = drs + NDM * NFRS;                        // ROW: D_pr | D_ps of the current row
    double *kpanel = dpx + NDM * NKP;                      // ROW: K_pr | K_ps of the current row
    uint32_t *queue = (uint32_t *)(drs + NDM * NFRS + ROWD);
    double *slots = drs + NDM * NFRS + ROWD + (qcap + 1) / 2;
    double *red_own = tm + ((8 + nrs * 5 + NDM * NFRS + ROWD + (qcap + 1) / 2 + nqs * slot_d + 1) & ~1);
    double *gsets = red_own + C::redd(ROW);
    constexpr bool RED_Q = C::RED_Q, RED_P = C::red_p(ROW);
    constexpr int REDS = C::REDS;

    const int nbas = t.nbas;
    const int nao = t.nao;
    const size_t nao2 = (size_t)nao * nao;
    const bool do_j = t.vj != nullptr;
    const bool do_k = t.vk != nullptr;
    const int *bas = t.bas;
    const double *env = t.env;
    const int *ao_loc = t.ao_loc;

    // contraction-phase role of this lane
    const int s4 = (TW == 1) ? lane / G : 0;
    const bool lane_on = (TW == 1) ? (s4 < NSW) : (lt < G);
    const int tg = lane_on ? ((TW == 1) ? lane - s4 * G : lt) : 0;
    const int ab = tg / NCS;                      // BL: ab is the p component a itself
    const int cpart = tg - ab * NCS;
    const int a = BL ? ab : ab / NFQ;
    const int b = BL ? 0 : ab - a * NFQ;          // BL: b is a loop index, not a lane property
    const int c0 = cpart * NC;
    const int eax = cart_x(LP, a) * (LQ + 1), eay = cart_y(LP, a) * (LQ + 1), eaz = cart_z(LP, a) * (LQ + 1);
    const int ebx0 = cart_x(LQ, b), eby0 = cart_y(LQ, b), ebz0 = cart_z(LQ, b);
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
        // ---------------- next X pair of this team ----------------
        jkw_team_sync<TW>(team);
        if (lt == 0) {
            if (ROW) {
                *(long long *)tm = (long long)atomicAdd((unsigned long long *)t.head, 1ull) * t.world + t.rank;
            } else {
                tm_ints[0] = atomicAdd(t.head, 1) * t.world + t.rank;
            }
            tm_ints[2] = 0;
            tm_ints[3] = 0;
        }
        jkw_team_sync<TW>(team);
        // flat: work item = (X pair ix, chunk of the q-sorted Y list); chunk-major order
        // so that the dense leading chunks of every X pair are handed out first
        // ROW : work item = (X pair ix, block of rows), row-block major
        int ix, ypos = 0, y_end = 0, row_begin = 0, row_end = 1;
        if (ROW) {
            const long long item = *(const long long *)tm;
            if (item >= (long long)t.npair_x * t.nrowblk) break;
            const int rb = (int)(item / t.npair_x);
            ix = (int)(item - (long long)rb * t.npair_x);
            row_begin = rb * t.rows_per_item;
            row_end = min(t.nrows, row_begin + t.rows_per_item);
        } else {
            const int item = tm_ints[0];
            if (item >= t.npair_x * t.nychunk) break;
            const int chunk = item / t.npair_x;
            ix = item - chunk * t.npair_x;
            ypos = chunk * t.ychunk;
            if (t.dedupe == 1 && ypos < ix) ypos = ix;
            y_end = min(t.npair_y, (chunk + 1) * t.ychunk);
        }
        const float thr = t.cutoff - t.q_x[ix];
        if (ROW) {
            bool any_row = false;
            for (int row = row_begin + lt; row < row_end; row += TEAM)
                any_row = any_row || (t.row_qmax[row] + t.dm_penalty >= thr);
            if (TW == 1) {
                if (!__any_sync(0xffffffffu, any_row)) continue;
            } else {
                if (any_row) tm_ints[2] = 1;
                jkw_team_sync<TW>(team);
                const int alive = tm_ints[2];
                jkw_team_sync<TW>(team);
                if (lt == 0) tm_ints[2] = 0;
                if (!alive) continue;
            }
        } else {
            if (ypos >= y_end || t.q_y[ypos] + t.dm_penalty + t.q_margin < thr) continue;   // dead chunk
        }
        const uint32_t pair_rs = t.pair_x[ix];
        const int rsh = pair_rs / nbas;
        const int ssh = pair_rs - rsh * nbas;
        const int r0 = ao_loc[rsh];
        const int s0 = ao_loc[ssh];
        {
            const double *Rr = env + bas[rsh * BAS_SLOTS + PTR_BAS_COORD];
            const double *Rs = env + bas[ssh * BAS_SLOTS + PTR_BAS_COORD];
            const double *er = env + bas[rsh * BAS_SLOTS + PTR_EXP];
            const double *es = env + bas[ssh * BAS_SLOTS + PTR_EXP];
            const double *cr = env + bas[rsh * BAS_SLOTS + PTR_COEFF];
            const double *cs = env + bas[ssh * BAS_SLOTS + PTR_COEFF];
            double dx = Rr[0] - Rs[0], dy = Rr[1] - Rs[1], dz = Rr[2] - Rs[2];
            double rr2 = dx * dx + dy * dy + dz * dz;
            for (int i = lt; i < nrs; i += TEAM) {
                int ir = i / t.nprim_s;
                int is = i - ir * t.nprim_s;
                double ar = er[ir], as = es[is];
                double eta = ar + as;
                double f = as / eta;
                double *o = rsprim + i * 5;
                o[0] = eta;
                o[1] = Rr[0] - f * dx;
                o[2] = Rr[1] - f * dy;
                o[3] = Rr[2] - f * dz;
                o[4] = cr[ir] * cs[is] * exp(-ar * f * rr2);
            }
            if (lt < 3) {
                rsgeom[lt] = Rr[lt] - Rs[lt];
                rsgeom[3 + lt] = Rr[lt];
            }
            for (int i = lt; i < NDM * NFRS; i += TEAM) {
                int idm = i / NFRS;
                int cd = i - idm * NFRS;
                int c = cd / NFS;
                int d = cd - c * NFS;
                drs[i] = __ldg(t.dm + idm * nao2 + (size_t)(r0 + c) * nao + s0 + d);
            }
        }
        const bool vec_r = NC > 1 && ((r0 | nao) & 1) == 0;
        const bool vec_s = NFS > 1 && ((s0 | nao) & 1) == 0;
        double jacc[NDM][NC][NFS];
#pragma unroll
        for (int idm = 0; idm < NDM; ++idm)
#pragma unroll
            for (int n = 0; n < NC; ++n)
#pragma unroll
                for (int d = 0; d < NFS; ++d) jacc[idm][n][d] = 0.;

        const float *dc_r = t.dm_cond + (size_t)rsh * nbas;
        const float *dc_s = t.dm_cond + (size_t)ssh * nbas;
        const float d_rs = dc_r[ssh];
        // K_pr / K_ps partial sums of this lane: per quartet (flat) or over the whole row (ROW)
        double accp[NDM][NC], accs[NDM][NFS];
        int psh_row = 0;
        float d_uni = -1e30f;
        const float *dc_prow = t.dm_cond;
        jkw_team_sync<TW>(team);

        for (int row = row_begin; row < row_end; ++row) {
        if (ROW) {
            // ---------------- next row: shell p, its partners [ypos, y_end) ----------------
            if (t.row_qmax[row] + t.dm_penalty < thr) continue;
            psh_row = t.row_p[row];
            const int p0 = ao_loc[psh_row];
            ypos = t.row_ptr[row];
            y_end = t.row_ptr[row + 1];
            dc_prow = t.dm_cond + (size_t)psh_row * nbas;
            d_uni = -1e30f;                               // team-uniform part of the density bound
            if (do_j) d_uni = d_rs;
            if (do_k) d_uni = fmaxf(d_uni, fmaxf(dc_r[psh_row], dc_s[psh_row]));
            jkw_team_sync<TW>(team);
            if (do_k) {
                for (int i = lt; i < NDM * NKP; i += TEAM) {
                    const int idm = i / NKP;
                    const int ak = i - idm * NKP;
                    const int ia = ak / (NFR + NFS);
                    const int k = ak - ia * (NFR + NFS);
                    const int col = k < NFR ? r0 + k : s0 + k - NFR;
                    dpx[i] = __ldg(t.dm + idm * nao2 + (size_t)(p0 + ia) * nao + col);
                    kpanel[i] = 0.;
                }
            }
#pragma unroll
            for (int idm = 0; idm < NDM; ++idm) {
#pragma unroll
                for (int n = 0; n < NC; ++n) accp[idm][n] = 0.;
#pragma unroll
                for (int d = 0; d < NFS; ++d) accs[idm][d] = 0.;
            }
            if (TW > 1 && lt == 0) {
                tm_ints[2] = 0;
                tm_ints[3] = 0;
            }
        }
        int qn = 0;
        int exhausted = 0;
        jkw_team_sync<TW>(team);

        while (true) {
            // ---------------- screen Y pairs into the team queue ----------------
            while (!exhausted && qn < nqs) {
                const int y = ypos + lt;
                bool keep = false, stop = false;
                if (y >= y_end) {
                    stop = true;
                } else if (ROW) {
                    // rows are not q-sorted: every partner of p is tested
                    const float qy = t.q_y[y];
                    keep = qy + t.dm_penalty >= thr;
                    if (keep) {
                        const int qsh = t.pair_y[y] - (uint32_t)psh_row * nbas;
                        if (t.dedupe) {
                            const uint32_t canon = psh_row >= qsh ? (uint32_t)psh_row * nbas + qsh
                                                                  : (uint32_t)qsh * nbas + psh_row;
                            keep = canon >= pair_rs;
                        }
                        float dmax = d_uni;
                        if (do_j) dmax = fmaxf(dmax, dc_prow[qsh]);
                        if (do_k) dmax = fmaxf(dmax, fmaxf(dc_r[qsh], dc_s[qsh]));
                        keep = keep && (dmax + qy > thr);
                        if (t.sr_s_x != nullptr && keep) keep = sr_keep(t, ix, y, dmax);
                    }
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
                            if (t.sr_s_x != nullptr && keep) keep = sr_keep(t, ix, y, dmax);
                        }
                    }
                }
                const unsigned m = __ballot_sync(0xffffffffu, keep);
                const int cnt = __popc(m);
                int base = qn;
                if (TW > 1) {
                    base = 0;
                    if (lane == 0 && cnt) base = atomicAdd(&tm_ints[2], cnt);
                    base = __shfl_sync(0xffffffffu, base, 0);
                }
                if (keep) queue[base + __popc(m & ((1u << lane) - 1u))] = (uint32_t)y;
                if (TW == 1) {
                    qn += cnt;
                    exhausted = __any_sync(0xffffffffu, stop);
                    __syncwarp();
                } else {
                    if (stop) tm_ints[3] = 1;
                    jkw_team_sync<TW>(team);
                    qn = tm_ints[2];
                    exhausted = tm_ints[3];
                    jkw_team_sync<TW>(team);
                }
                ypos += TEAM;
            }
            if (qn == 0) break;                       // exhausted and nothing queued
            const int nb = qn < nqs ? qn : nqs;

            // ---------------- quartet slots: Gaussian products of the Y pairs ----------------
            for (int i = lt; i < nqs * npq; i += TEAM) {
                const int j = i / npq;
                const int ip = i - j * npq;
                const bool active = j < nb;
                const int iy = queue[active ? j : 0];
                double *sl = slots + j * slot_d;
                double *o = sl + JKW_SLOT_HDR + ip * 5;
                if (t.rec_y != nullptr) {
                    // precomputed record of the pair (B200_pair_records)
                    const double *rec = t.rec_y + (size_t)iy * t.rec_stride;
                    const double *pr = rec + PAIR_REC_HDR + ip * 5;
                    o[0] = pr[0]; o[1] = pr[1]; o[2] = pr[2]; o[3] = pr[3]; o[4] = pr[4];
                    if (ip == 0) {
                        sl[0] = rec[0]; sl[1] = rec[1]; sl[2] = rec[2];
                        sl[3] = rec[3]; sl[4] = rec[4]; sl[5] = rec[5];
                        double fac = PI_FAC * t.scale * rec[7];
                        if (rsh == ssh) fac *= .5;
                        if (ROW) {
                            if (t.dedupe) {
                                const int qsh = t.pair_y[iy] - (uint32_t)psh_row * nbas;
                                const uint32_t canon = psh_row >= qsh ? (uint32_t)psh_row * nbas + qsh
                                                                      : (uint32_t)qsh * nbas + psh_row;
                                if (canon == pair_rs) fac *= .5;
                            }
                        } else if (t.dedupe == 1 && iy == ix) {
                            fac *= .5;
                        }
                        sl[6] = active ? fac : 0.;
                        sl[7] = rec[6];                    // (p0, q0)
                    }
                    continue;
                }
                const uint32_t pair_pq = t.pair_y[iy];
                const int psh = pair_pq / nbas;
                const int qsh = pair_pq - psh * nbas;
                const double *Rp = env + bas[psh * BAS_SLOTS + PTR_BAS_COORD];
                const double *Rq = env + bas[qsh * BAS_SLOTS + PTR_BAS_COORD];
                const double *ep = env + bas[psh * BAS_SLOTS + PTR_EXP];
                const double *eq = env + bas[qsh * BAS_SLOTS + PTR_EXP];
                const double *cp = env + bas[psh * BAS_SLOTS + PTR_COEFF];
                const double *cq = env + bas[qsh * BAS_SLOTS + PTR_COEFF];
                double dx = Rp[0] - Rq[0], dy = Rp[1] - Rq[1], dz = Rp[2] - Rq[2];
                double rr2 = dx * dx + dy * dy + dz * dz;
                const int ipp = ip / t.nprim_q;
                const int iq = ip - ipp * t.nprim_q;
                double ap = ep[ipp], aq = eq[iq];
                double zeta = ap + aq;
                double f = aq / zeta;
                o[0] = zeta;
                o[1] = Rp[0] - f * dx;
                o[2] = Rp[1] - f * dy;
                o[3] = Rp[2] - f * dz;
                o[4] = cp[ipp] * cq[iq] * exp(-ap * f * rr2);
                if (ip == 0) {
                    sl[0] = dx; sl[1] = dy; sl[2] = dz;
                    sl[3] = Rp[0]; sl[4] = Rp[1]; sl[5] = Rp[2];
                    double fac = PI_FAC * t.scale;
                    if (psh == qsh) fac *= .5;
                    if (rsh == ssh) fac *= .5;
                    if (pair_pq == pair_rs) fac *= .5;
                    sl[6] = active ? fac : 0.;
                    int *pi = (int *)(sl + 7);
                    pi[0] = ao_loc[psh];
                    pi[1] = ao_loc[qsh];
                }
            }
            // 2D integrals of one (g-set, root, axis): VRR + TRR in registers, bra HRR (and the ket
            // HRR when the tile is split), result to shared memory as
            // g[root*3+axis][(ea, eb)][segment], (ea, eb) segments SUBP apart
            auto recur_axis = [&](double *gs, const double *sl, const double P, const double Q, const int ra,
                                  const int axis, const double g00, const double rt_aij, const double rt_akl,
                                  const double b00, const double b10, const double b01) {
                const double PQ = P - Q;
                const double c0p = (P - sl[3 + axis]) - rt_aij * PQ;
                const double cpq = (Q - rsgeom[3 + axis]) + rt_akl * PQ;
                double gg[LRS + 1][LPQ + 1];
                gg[0][0] = g00;
                if (LPQ > 0) {
                    gg[0][1] = c0p * gg[0][0];
#pragma unroll
                    for (int e = 1; e < LPQ; ++e)
                        gg[0][e + 1] = c0p * gg[0][e] + e * b10 * gg[0][e - 1];
                }
                if (LRS > 0) {
                    gg[1][0] = cpq * gg[0][0];
#pragma unroll
                    for (int e = 1; e <= LPQ; ++e)
                        gg[1][e] = cpq * gg[0][e] + e * b00 * gg[0][e - 1];
#pragma unroll
                    for (int f = 1; f < LRS; ++f) {
                        gg[f + 1][0] = cpq * gg[f][0] + f * b01 * gg[f - 1][0];
#pragma unroll
                        for (int e = 1; e <= LPQ; ++e)
                            gg[f + 1][e] = cpq * gg[f][e] + f * b01 * gg[f - 1][e] + e * b00 * gg[f][e - 1];
                    }
                }
                const double AB = sl[axis];
                const double CD = rsgeom[axis];
                double *o = gs + 2 * NROOTS + ra * GSIZEP;
#pragma unroll
                for (int eb = 0; eb <= LQ; ++eb) {
                    // gg[f][ea] now holds the bra-transferred value for (ea, eb)
#pragma unroll
                    for (int ea = 0; ea <= LP; ++ea) {
                        double v[LRS + 1];
#pragma unroll
                        for (int f = 0; f <= LRS; ++f) v[f] = gg[f][ea];
                        if (NCS == 1) {
                            jkw_st_seg<LRS + 1>(o + (ea * (LQ + 1) + eb) * SUBP, v);
                        } else {
                            double seg[SUB];
#pragma unroll
                            for (int d = 0; d <= LS; ++d) {
#pragma unroll
                                for (int c = 0; c <= LR; ++c) seg[c * (LS + 1) + d] = v[c];
                                if (d < LS) {
#pragma unroll
                                    for (int c = 0; c < LRS - d; ++c) v[c] = v[c + 1] + CD * v[c];
                                }
                            }
                            jkw_st_seg<SUB>(o + (ea * (LQ + 1) + eb) * SUBP, seg);
                        }
                    }
                    if (eb < LQ) {
#pragma unroll
                        for (int f = 0; f <= LRS; ++f)
#pragma unroll
                            for (int e = 0; e < LPQ - eb; ++e) gg[f][e] = gg[f][e + 1] + AB * gg[f][e];
                    }
                }
            };
            // roots, weights and 2D integrals of every g-set of the team, ONE phase: a lane owns a
            // (g-set, root) -- the scalars of the primitive quartet (one rsqrt, one division), both
            // Rys series, then the three axes.  Uncontracted launches: g-set = quartet slot;
            // contracted: g-set k*NSW+j = primitive quartet pass*K+k of quartet slot j.
            auto build_gsets = [&](const int pass) {
                if (pass > 0) jkw_team_sync<TW>(team);      // the previous pass has been contracted by every lane
                for (int it = lt; it < ngs * NROOTS; it += TEAM) {
                    const int g = it / NROOTS;
                    const int root = it - g * NROOTS;
                    int j = g, pi = 0;
                    if (!single) {
                        const int k = g / NSW;
                        j = g - k * NSW;
                        pi = pass * K + k;
                    }
                    const bool valid = pi < nprimq;
                    if (!valid) pi = 0;
                    const bool term2 = pi >= npq * nrs;
                    if (term2) pi -= npq * nrs;
                    const double om = term2 ? t.omega2 : t.omega;
                    const int ipq = pi / nrs;
                    const int irs = pi - ipq * nrs;
                    const double *sl = slots + j * slot_d;
                    const double *pp = sl + JKW_SLOT_HDR + ipq * 5;
                    const double *rr = rsprim + irs * 5;
                    const double zeta = pp[0], eta = rr[0];
                    const double dx = pp[1] - rr[1], dy = pp[2] - rr[2], dz = pp[3] - rr[3];
                    const double rsq = rsqrt(zeta + eta);
                    const double rze = rsq * rsq;                   // 1 / (zeta + eta)
                    const double ize = 1. / (zeta * eta);
                    const double theta = zeta * eta * rze;
                    double x = theta * (dx * dx + dy * dy + dz * dz);
                    double rscale = rze;
                    double w = valid ? sl[6] * pp[4] * rr[4] * ize * rsq : 0.;
                    if (term2) w *= t.ratio2;
                    if (om != 0.) {                                 // long range: theta -> theta * lrf
                        const double lrf = om * om / (om * om + theta);
                        x *= lrf;
                        rscale *= lrf;
                        w *= sqrt(lrf);
                    }
                    const double rt_aa = rys_value_s<NROOTS>(root, x, rys_tab) * rscale;
                    const double wt = rys_value_s<NROOTS>(NROOTS + root, x, rys_tab) * w;
                    const double rt_aij = rt_aa * eta;
                    const double rt_akl = rt_aa * zeta;
                    const double b00 = .5 * rt_aa;
                    const double b10 = .5 * eta * ize * (1. - rt_aij);
                    const double b01 = .5 * zeta * ize * (1. - rt_akl);
                    double *gs = gsets + g * GSET;
#pragma unroll
                    for (int axis = 0; axis < 3; ++axis)
                        recur_axis(gs, sl, pp[1 + axis], rr[1 + axis], root * 3 + axis, axis,
                                   axis == 2 ? wt : 1., rt_aij, rt_akl, b00, b10, b01);
                }
                jkw_team_sync<TW>(team);
            };
            // phase 3: gout += gx * gy * gz over the roots of g-set g (register tile of this lane)
            auto contract = [&](double (&gout)[NC][NFS], const int g, const int bb) {
                const double *gk = gsets + g * GSET + 2 * NROOTS;
                // BL: bb is a compile-time constant after unrolling; else the lane's own b
                const int gbx = (eax + (BL ? cart_x(LQ, bb) : ebx0)) * SUBP;
                const int gby = (eay + (BL ? cart_y(LQ, bb) : eby0)) * SUBP + GSIZEP;
                const int gbz = (eaz + (BL ? cart_z(LQ, bb) : ebz0)) * SUBP + 2 * GSIZEP;
                const double cdx = rsgeom[0], cdy = rsgeom[1], cdz = rsgeom[2];
                if (lane_on) {
                    if (NCS == 1) {
#pragma unroll
                        for (int root = 0; root < NROOTS; ++root) {
                            const double *gr = gk + root * 3 * GSIZEP;
                            double gx[SUB], gy[SUB], gz[SUB];
                            {   // ket transfer (r <- s) of this lane's three bra-transferred rows
                                double hx[LRS + 1], hy[LRS + 1], hz[LRS + 1];
                                jkw_ld_seg<LRS + 1>(hx, gr + gbx);
                                jkw_ld_seg<LRS + 1>(hy, gr + gby);
                                jkw_ld_seg<LRS + 1>(hz, gr + gbz);
#pragma unroll
                                for (int d = 0; d <= LS; ++d) {
#pragma unroll
                                    for (int c = 0; c <= LR; ++c) {
                                        gx[c * (LS + 1) + d] = hx[c];
                                        gy[c * (LS + 1) + d] = hy[c];
                                        gz[c * (LS + 1) + d] = hz[c];
                                    }
                                    if (d < LS) {
#pragma unroll
                                        for (int c = 0; c < LRS - d; ++c) {
                                            hx[c] = hx[c + 1] + cdx * hx[c];
                                            hy[c] = hy[c + 1] + cdy * hy[c];
                                            hz[c] = hz[c + 1] + cdz * hz[c];
                                        }
                                    }
                                }
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
                            const double *gr = gk + root * 3 * GSIZEP;
#pragma unroll
                            for (int n = 0; n < NC; ++n) {
                                double gx[LS + 1], gy[LS + 1], gz[LS + 1];
                                if ((LS + 1) % 2 == 0) {
                                    jkw_ld_seg<LS + 1>(gx, gr + gbx + ox[n]);
                                    jkw_ld_seg<LS + 1>(gy, gr + gby + oy[n]);
                                    jkw_ld_seg<LS + 1>(gz, gr + gbz + oz[n]);
                                } else {
#pragma unroll
                                    for (int i = 0; i <= LS; ++i) {
                                        gx[i] = gr[gbx + ox[n] + i];
                                        gy[i] = gr[gby + oy[n] + i];
                                        gz[i] = gr[gbz + oz[n] + i];
                                    }
                                }
#pragma unroll
                                for (int d = 0; d < NFS; ++d) {
                                    gout[n][d] += gx[cart_x(LS, d)] * gy[cart_y(LS, d)] * gz[cart_z(LS, d)];
                                }
                            }
                        }
                    }
                }
            };
            // J/K digestion of quartet slot j with this lane's tile (reference: rys_contract_jk.cu:304-329)
            // Within a single-warp team the lanes (a, b) of one quartet hold partial sums of the
            // same K rows: K_p* over b, K_q* over a.  They are combined with segmented
            // shuffle reductions so that one lane per row issues the atomic.
            constexpr bool SHFL_RED = (TW == 1);
            auto red_b = [&](double v) {          // sum over b (lane stride NCS), result in the b == 0 lane
#pragma unroll
                for (int off = 1; off < NFQ; off <<= 1) {
                    const double tv = __shfl_down_sync(0xffffffffu, v, off * NCS);
                    if (b + off < NFQ) v += tv;
                }
                return v;
            };
            auto red_a = [&](double v) {          // sum over a (lane stride NFQ*NCS), result in the a == 0 lane
#pragma unroll
                for (int off = 1; off < C::NFP; off <<= 1) {
                    const double tv = __shfl_down_sync(0xffffffffu, v, off * (BL ? 1 : NFQ) * NCS);
                    if (a + off < C::NFP) v += tv;
                }
                return v;
            };
            // digestion of the (a, bb) tile of quartet slot j.  accp / accs collect K_pr / K_ps of
            // row a over the b loop (BL) and are flushed when `last`.
            auto digest = [&](double (&gout)[NC][NFS], const int j, const int bb,
                              double (&accp)[NDM][NC], double (&accs)[NDM][NFS], const bool last) {
                const bool on = lane_on && j < nb;
                const int jbase = j - s4;             // first quartet slot of this pass
                int p0 = 0, q0 = 0;
                if (on) {
                    const int *pi = (const int *)(slots + j * slot_d + 7);
                    p0 = pi[0];
                    q0 = pi[1];
                }
                const size_t prow = (size_t)(p0 + a) * nao;
                const size_t qrow = (size_t)(q0 + bb) * nao;
#pragma unroll
                for (int idm = 0; idm < NDM; ++idm) {
                    const double *dm = t.dm + idm * nao2;
                    if (do_j && on) {
                        double *vj = t.vj + idm * nao2;
                        double dpq = __ldg(dm + prow + q0 + bb);
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
                        JK_ATOMIC_ADD(vj + prow + q0 + bb, s);
                    }
                    if (do_k) {
                        double *vk = t.vk + idm * nao2;
                        double kqr[NC], kqs[NFS];
#pragma unroll
                        for (int n = 0; n < NC; ++n) kqr[n] = 0.;
#pragma unroll
                        for (int d = 0; d < NFS; ++d) kqs[d] = 0.;
                        if (on) {
                            double dqs[NFS], dps[NFS];
                            double dqr[NC], dpr[NC];
                            if (ROW) {
                                // q is the column index: neighbouring partners of p -> neighbouring addresses;
                                // the p rows come from the block staged at the row start
                                const double *dT = t.dmT + idm * nao2 + q0 + bb;
                                const double *dp = dpx + idm * NKP + a * (NFR + NFS);
#pragma unroll
                                for (int d = 0; d < NFS; ++d) {
                                    dqs[d] = __ldg(dT + (size_t)(s0 + d) * nao);
                                    dps[d] = dp[NFR + d];
                                }
#pragma unroll
                                for (int n = 0; n < NC; ++n) {
                                    const int c = c0 + n < NFR ? c0 + n : NFR - 1;
                                    dqr[n] = __ldg(dT + (size_t)(r0 + c) * nao);
                                    dpr[n] = dp[c];
                                }
                            } else {
                            ld_row<NFS>(dqs, dm + qrow + s0, vec_s);
                            ld_row<NFS>(dps, dm + prow + s0, vec_s);
                            if (NCS == 1) {
                                ld_row<NC>(dqr, dm + qrow + r0, vec_r);
                                ld_row<NC>(dpr, dm + prow + r0, vec_r);
                            } else {
#pragma unroll
                                for (int n = 0; n < NC; ++n) {
                                    const int c = c0 + n < NFR ? c0 + n : NFR - 1;
                                    dqr[n] = __ldg(dm + qrow + r0 + c);
                                    dpr[n] = __ldg(dm + prow + r0 + c);
                                }
                            }
                            }
#pragma unroll
                            for (int n = 0; n < NC; ++n) {
                                if (c0 + n < NFR) {
#pragma unroll
                                    for (int d = 0; d < NFS; ++d) {
                                        const double gv = gout[n][d];
                                        accp[idm][n] += gv * dqs[d];
                                        kqr[n] += gv * dps[d];
                                        accs[idm][d] += gv * dqr[n];
                                        kqs[d] += gv * dpr[n];
                                    }
                                }
                            }
                        }
                        if (RED_Q) {
                            // ---- shared-memory transpose reduction: every lane parks its partial K_q* sums
                            // (and, when they are complete, its K_p* sums); one lane per output row element
                            // then adds the NFP (NFQ) partials and issues the atomic
                            const bool stage_p = RED_P && last;
                            double *red = C::redd(ROW) > 0 ? red_own : gsets + jbase * GSET;
                            __syncwarp();
                            if (lane_on) {
#pragma unroll
                                for (int n = 0; n < NC; ++n) red[n * REDS + lane] = kqr[n];
#pragma unroll
                                for (int d = 0; d < NFS; ++d) red[(NC + d) * REDS + lane] = kqs[d];
                            }
                            __syncwarp();
                            constexpr int NKK = NFR + NFS;
                            constexpr int NOQ = NSW * NBLANE * NKK;       // K_q* outputs: (slot, b, k)
                            for (int o = lane; o < NOQ; o += 32) {
                                const int so = o / (NBLANE * NKK);
                                const int rem = o - so * (NBLANE * NKK);
                                const int bo = rem / NKK;
                                const int k = rem - bo * NKK;
                                const int jo = jbase + so;
                                if (jo >= nb) continue;
                                const int lane0 = so * G + bo * NCS;
                                double sum = 0., sum1 = 0.;        // two chains: half the dependent adds
                                if (k < NFR) {
                                    const int cp = k / NC;
                                    const double *rp = red + (k - cp * NC) * REDS + lane0 + cp;
#pragma unroll
                                    for (int ia = 0; ia + 1 < C::NFP; ia += 2) {
                                        sum += rp[ia * NBLANE * NCS];
                                        sum1 += rp[(ia + 1) * NBLANE * NCS];
                                    }
                                    if (C::NFP & 1) sum += rp[(C::NFP - 1) * NBLANE * NCS];
                                } else {
                                    const double *rp = red + (NC + k - NFR) * REDS + lane0;
#pragma unroll
                                    for (int ia = 0; ia < C::NFP; ++ia)
#pragma unroll
                                        for (int cp = 0; cp < NCS; ++cp) {
                                            if ((ia * NCS + cp) & 1) sum1 += rp[ia * NBLANE * NCS + cp];
                                            else sum += rp[ia * NBLANE * NCS + cp];
                                        }
                                }
                                sum += sum1;
                                if (sum != 0.) {
                                    const int qo = ((const int *)(slots + jo * slot_d + 7))[1] + (BL ? bb : bo);
                                    const int col = k < NFR ? r0 + k : s0 + k - NFR;
                                    if (ROW) JK_ATOMIC_ADD(t.vkT + idm * nao2 + (size_t)col * nao + qo, sum);
                                    else JK_ATOMIC_ADD(vk + (size_t)qo * nao + col, sum);
                                }
                            }
                            if (stage_p) {
                                // same scratch, second round: the completed K_p* partial sums, summed over b
                                __syncwarp();
                                if (lane_on) {
#pragma unroll
                                    for (int n = 0; n < NC; ++n) red[n * REDS + lane] = accp[idm][n];
#pragma unroll
                                    for (int d = 0; d < NFS; ++d) red[(NC + d) * REDS + lane] = accs[idm][d];
                                }
                                __syncwarp();
                                constexpr int NOP = NSW * C::NFP * NKK;   // K_p* outputs: (slot, a, k)
                                for (int o = lane; o < NOP; o += 32) {
                                    const int so = o / (C::NFP * NKK);
                                    const int rem = o - so * (C::NFP * NKK);
                                    const int ao = rem / NKK;
                                    const int k = rem - ao * NKK;
                                    const int jo = jbase + so;
                                    if (jo >= nb) continue;
                                    const int lane0 = so * G + ao * NFQ * NCS;
                                    double sum = 0., sum1 = 0.;
                                    if (k < NFR) {
                                        const int cp = k / NC;
                                        const double *rp = red + (k - cp * NC) * REDS + lane0 + cp;
#pragma unroll
                                        for (int ib = 0; ib + 1 < NFQ; ib += 2) {
                                            sum += rp[ib * NCS];
                                            sum1 += rp[(ib + 1) * NCS];
                                        }
                                        if (NFQ & 1) sum += rp[(NFQ - 1) * NCS];
                                    } else {
                                        const double *rp = red + (NC + k - NFR) * REDS + lane0;
#pragma unroll
                                        for (int ib = 0; ib < NFQ; ++ib)
#pragma unroll
                                            for (int cp = 0; cp < NCS; ++cp) {
                                                if ((ib * NCS + cp) & 1) sum1 += rp[ib * NCS + cp];
                                                else sum += rp[ib * NCS + cp];
                                            }
                                    }
                                    sum += sum1;
                                    if (sum != 0.) {
                                        const int po = ((const int *)(slots + jo * slot_d + 7))[0] + ao;
                                        const int col = k < NFR ? r0 + k : s0 + k - NFR;
                                        JK_ATOMIC_ADD(vk + (size_t)po * nao + col, sum);
                                    }
                                }
                            } else if (last && !ROW) {
                                // b-loop variant or s-type q shell: the K_p* rows are complete in the lane
                                if (on) {
#pragma unroll
                                    for (int n = 0; n < NC; ++n)
                                        if (c0 + n < NFR) JK_ATOMIC_ADD(vk + prow + r0 + c0 + n, accp[idm][n]);
#pragma unroll
                                    for (int d = 0; d < NFS; ++d) JK_ATOMIC_ADD(vk + prow + s0 + d, accs[idm][d]);
                                }
                            }
                        } else {
                        // K_q* rows: sum over a across lanes, one atomic from the a == 0 lane
                        bool lead_q = true;
                        if (SHFL_RED && C::NFP > 1) {
#pragma unroll
                            for (int n = 0; n < NC; ++n) kqr[n] = red_a(kqr[n]);
#pragma unroll
                            for (int d = 0; d < NFS; ++d) kqs[d] = red_a(kqs[d]);
                            lead_q = a == 0;
                        }
                        if (on && lead_q) {
                            if (ROW) {
                                double *kT = t.vkT + idm * nao2 + q0 + bb;
#pragma unroll
                                for (int n = 0; n < NC; ++n)
                                    if (c0 + n < NFR) JK_ATOMIC_ADD(kT + (size_t)(r0 + c0 + n) * nao, kqr[n]);
#pragma unroll
                                for (int d = 0; d < NFS; ++d) JK_ATOMIC_ADD(kT + (size_t)(s0 + d) * nao, kqs[d]);
                            } else {
#pragma unroll
                            for (int n = 0; n < NC; ++n)
                                if (c0 + n < NFR) JK_ATOMIC_ADD(vk + qrow + r0 + c0 + n, kqr[n]);
#pragma unroll
                            for (int d = 0; d < NFS; ++d) JK_ATOMIC_ADD(vk + qrow + s0 + d, kqs[d]);
                            }
                        }
                        // K_p* rows: complete after the b loop (BL) or after the shuffle over b;
                        // ROW: they keep accumulating until the row ends
                        if (last && !ROW) {
                            bool lead_p = true;
                            if (!BL && SHFL_RED && NFQ > 1) {
#pragma unroll
                                for (int n = 0; n < NC; ++n) accp[idm][n] = red_b(accp[idm][n]);
#pragma unroll
                                for (int d = 0; d < NFS; ++d) accs[idm][d] = red_b(accs[idm][d]);
                                lead_p = b == 0;
                            }
                            if (on && lead_p) {
#pragma unroll
                                for (int n = 0; n < NC; ++n)
                                    if (c0 + n < NFR) JK_ATOMIC_ADD(vk + prow + r0 + c0 + n, accp[idm][n]);
#pragma unroll
                                for (int d = 0; d < NFS; ++d) JK_ATOMIC_ADD(vk + prow + s0 + d, accs[idm][d]);
                            }
                        }
                        }   // !RED_Q
                    }
                }
            };
            jkw_team_sync<TW>(team);
            if (single) {
                // K different quartets per lane slot, one primitive quartet each
                build_gsets(0);
                for (int k = 0; k < K; ++k) {
                    if (!ROW) {
#pragma unroll
                        for (int idm = 0; idm < NDM; ++idm) {
#pragma unroll
                            for (int n = 0; n < NC; ++n) accp[idm][n] = 0.;
#pragma unroll
                            for (int d = 0; d < NFS; ++d) accs[idm][d] = 0.;
                        }
                    }
#pragma unroll
                    for (int bi = 0; bi < NBL; ++bi) {
                        const int bb = BL ? bi : b;
                        double gout[NC][NFS];
#pragma unroll
                        for (int n = 0; n < NC; ++n)
#pragma unroll
                            for (int d = 0; d < NFS; ++d) gout[n][d] = 0.;
                        contract(gout, k * NSW + s4, bb);
                        digest(gout, k * NSW + s4, bb, accp, accs, bi == NBL - 1);
                    }
                }
            } else {
                // contracted class (BL == 0 only): K primitive quartets of the same quartet per pass
                double gout[NC][NFS];
                if (!ROW) {
#pragma unroll
                    for (int idm = 0; idm < NDM; ++idm) {
#pragma unroll
                        for (int n = 0; n < NC; ++n) accp[idm][n] = 0.;
#pragma unroll
                        for (int d = 0; d < NFS; ++d) accs[idm][d] = 0.;
                    }
                }
#pragma unroll
                for (int n = 0; n < NC; ++n)
#pragma unroll
                    for (int d = 0; d < NFS; ++d) gout[n][d] = 0.;
                for (int pass = 0; pass < npass; ++pass) {
                    build_gsets(pass);
                    for (int k = 0; k < K; ++k) contract(gout, k * NSW + s4, b);
                }
                digest(gout, s4, b, accp, accs, true);
            }

            // ---------------- keep the unprocessed remainder of the queue ----------------
            jkw_team_sync<TW>(team);
            if (lt == 0 && t.counter != nullptr) atomicAdd(t.counter, (unsigned long long)nb);
            const int rem = qn - nb;
            uint32_t keepv = 0;
            if (lt < rem) keepv = queue[nb + lt];
            jkw_team_sync<TW>(team);
            if (lt < rem) queue[lt] = keepv;
            qn = rem;
            if (TW > 1) {
                if (lt == 0) tm_ints[2] = rem;
            }
            jkw_team_sync<TW>(team);
            if (exhausted && qn == 0) break;
        }

        if (ROW && do_k) {
            // ---------------- row end: K_pr / K_ps leave the team ----------------
            // lanes holding the same p component a add to the same panel entries: one round per
            // (quartet slot, b, c part), so that every round touches each entry at most once
            constexpr int NROUND = NSW * NBLANE * NCS;
            const int myround = (s4 * NBLANE + (BL ? 0 : b)) * NCS + cpart;
            for (int rd = 0; rd < NROUND; ++rd) {
                if (lane_on && myround == rd) {
                    double *kp = kpanel + a * (NFR + NFS);
#pragma unroll
                    for (int idm = 0; idm < NDM; ++idm) {
#pragma unroll
                        for (int n = 0; n < NC; ++n)
                            if (c0 + n < NFR) kp[idm * NKP + c0 + n] += accp[idm][n];
#pragma unroll
                        for (int d = 0; d < NFS; ++d) kp[idm * NKP + NFR + d] += accs[idm][d];
                    }
                }
                jkw_team_sync<TW>(team);
            }
            const int p0 = ao_loc[psh_row];
            for (int i = lt; i < NDM * NKP; i += TEAM) {
                const double v = kpanel[i];
                if (v != 0.) {
                    const int idm = i / NKP;
                    const int ak = i - idm * NKP;
                    const int ia = ak / (NFR + NFS);
                    const int k = ak - ia * (NFR + NFS);
                    const int col = k < NFR ? r0 + k : s0 + k - NFR;
                    JK_ATOMIC_ADD(t.vk + idm * nao2 + (size_t)(p0 + ia) * nao + col, v);
                }
            }
        }
        }   // rows of the work item (one pass when !ROW)

        // ---------------- flush J_rs ----------------
        if (do_j) {
#pragma unroll
            for (int idm = 0; idm < NDM; ++idm) {
                double *vj = t.vj + idm * nao2;
#pragma unroll
                for (int n = 0; n < NC; ++n) {
#pragma unroll
                    for (int d = 0; d < NFS; ++d) {
                        double v = jacc[idm][n][d];
                        if (TW == 1 && NCS == 1) {
                            // every lane holds a partial sum of the same (c,d): butterfly, one atomic per warp
#pragma unroll
                            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                            if (lane == 0 && v != 0.) JK_ATOMIC_ADD(vj + (size_t)(r0 + n) * nao + s0 + d, v);
                        } else if (lane_on && c0 + n < NFR && v != 0.) {
                            JK_ATOMIC_ADD(vj + (size_t)(r0 + c0 + n) * nao + s0 + d, v);
                        }
                    }
                }
            }
        }
    }
}
