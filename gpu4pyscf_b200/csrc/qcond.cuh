// Schwarz bounds q_ij = 1/2 log max_ab |(ab|ab)| per shell pair, FP64 throughout.
// Counterpart of the reference's int2e_qcond_kernel
// (gpu4pyscf/lib/gvhf-rys/nr_sr_estimator.cu:99-396), which evaluates the same
// quantity in FP32; here one thread owns one shell pair and only the diagonal
// elements g(a,b,a,b) of the 2D integrals are transferred.
#pragma once
#include "common.cuh"
#include "rys_roots.cuh"

// (ij|ij) diagonal of one shell pair for the operator c_full / r + c_lr * erf(omega r) / r
// (the same two terms as the J/K kernels, reference: rys_roots_for_k.cu:25-60);
// returns 1/2 log max |(ab|ab)|, -700 for a vanishing pair.
template <int LI, int LJ>
__device__ float qcond_pair(const uint32_t pair, const int *bas, const double *env, const int nbas,
                            const int nprim_i, const int nprim_j,
                            const double *cheb, const double *large,
                            const double c_full, const double c_lr, const double omega)
{
    constexpr int L2 = LI + LJ;
    constexpr int NROOTS = L2 + 1;
    constexpr int NFI = ncart(LI), NFJ = ncart(LJ);
    int ish = pair / nbas;
    int jsh = pair - ish * nbas;
    const double *Ri = env + bas[ish * BAS_SLOTS + PTR_BAS_COORD];
    const double *Rj = env + bas[jsh * BAS_SLOTS + PTR_BAS_COORD];
    const double *ei = env + bas[ish * BAS_SLOTS + PTR_EXP];
    const double *ej = env + bas[jsh * BAS_SLOTS + PTR_EXP];
    const double *ci = env + bas[ish * BAS_SLOTS + PTR_COEFF];
    const double *cj = env + bas[jsh * BAS_SLOTS + PTR_COEFF];
    double AB[3] = {Ri[0] - Rj[0], Ri[1] - Rj[1], Ri[2] - Rj[2]};
    double rr2 = AB[0] * AB[0] + AB[1] * AB[1] + AB[2] * AB[2];
    double val[NFI][NFJ];
    for (int a = 0; a < NFI; ++a)
        for (int b = 0; b < NFJ; ++b) val[a][b] = 0.;
    int npp = nprim_i * nprim_j;
    for (int ij = 0; ij < npp; ++ij) {
        int ip = ij / nprim_j, jp = ij - ip * nprim_j;
        double ai = ei[ip], aj = ej[jp];
        double zeta = ai + aj;
        double f1 = aj / zeta;
        double P[3] = {Ri[0] - f1 * AB[0], Ri[1] - f1 * AB[1], Ri[2] - f1 * AB[2]};
        double cij = ci[ip] * cj[jp] * exp(-ai * f1 * rr2);
        for (int kl = 0; kl < npp; ++kl) {
            int kp = kl / nprim_j, lp = kl - kp * nprim_j;
            double ak = ei[kp], al = ej[lp];
            double eta = ak + al;
            double f2 = al / eta;
            double Q[3] = {Ri[0] - f2 * AB[0], Ri[1] - f2 * AB[1], Ri[2] - f2 * AB[2]};
            double ckl = ci[kp] * cj[lp] * exp(-ak * f2 * rr2);
            double ze = zeta + eta;
            double dx = P[0] - Q[0], dy = P[1] - Q[1], dz = P[2] - Q[2];
            double x = zeta * eta / ze * (dx * dx + dy * dy + dz * dz);
            const double fac0 = PI_FAC * cij * ckl / (zeta * eta * sqrt(ze));
            const double x0 = x;
            for (int term = 0; term < 2; ++term) {
            // term 0: c_full / r; term 1: c_lr * erf(omega r) / r (theta -> theta * lrf, rys_roots.cu:91-119)
            double fac = fac0 * (term == 0 ? c_full : c_lr);
            if (fac == 0.) continue;
            double rscale = 1. / ze;
            x = x0;
            if (term == 1) {
                const double theta = zeta * eta / ze;
                const double lrf = omega * omega / (omega * omega + theta);
                x *= lrf;
                rscale *= lrf;
                fac *= sqrt(lrf);
            }
            for (int root = 0; root < NROOTS; ++root) {
                double rt_aa = rys_value<NROOTS>(root, x, cheb, large) * rscale;
                double w = rys_value<NROOTS>(NROOTS + root, x, cheb, large) * fac;
                double rt_aij = rt_aa * eta, rt_akl = rt_aa * zeta;
                double b00 = .5 * rt_aa;
                double b10 = .5 / zeta * (1. - rt_aij);
                double b01 = .5 / eta * (1. - rt_akl);
                double diag[3][LI + 1][LJ + 1];
                for (int axis = 0; axis < 3; ++axis) {
                    double PQ = P[axis] - Q[axis];
                    double c0p = (P[axis] - Ri[axis]) - rt_aij * PQ;
                    double cpq = (Q[axis] - Ri[axis]) + rt_akl * PQ;
                    double g[L2 + 1][L2 + 1];    // [f][e]
                    g[0][0] = (axis == 2) ? w : 1.;
                    for (int e = 0; e < L2; ++e)
                        g[0][e + 1] = c0p * g[0][e] + (e > 0 ? e * b10 * g[0][e - 1] : 0.);
                    for (int f = 0; f < L2; ++f)
                        for (int e = 0; e <= L2; ++e)
                            g[f + 1][e] = cpq * g[f][e] + (f > 0 ? f * b01 * g[f - 1][e] : 0.)
                                        + (e > 0 ? e * b00 * g[f][e - 1] : 0.);
                    for (int ea = 0; ea <= LI; ++ea) {
                        for (int eb = 0; eb <= LJ; ++eb) {
                            double v[L2 + 1];
                            for (int f = 0; f <= L2; ++f) {
                                double u[LJ + 1];
                                for (int e = 0; e <= eb; ++e) u[e] = g[f][ea + e];
                                for (int k = 0; k < eb; ++k)
                                    for (int e = 0; e < eb - k; ++e) u[e] = u[e + 1] + AB[axis] * u[e];
                                v[f] = u[0];
                            }
                            // ket transfer to (c = ea, d = eb)
                            for (int d = 0; d < eb; ++d)
                                for (int c = 0; c < L2 - d; ++c) v[c] = v[c + 1] + AB[axis] * v[c];
                            diag[axis][ea][eb] = v[ea];
                        }
                    }
                }
                for (int a = 0; a < NFI; ++a)
                    for (int b = 0; b < NFJ; ++b)
                        val[a][b] += diag[0][cart_x(LI, a)][cart_x(LJ, b)]
                                   * diag[1][cart_y(LI, a)][cart_y(LJ, b)]
                                   * diag[2][cart_z(LI, a)][cart_z(LJ, b)];
            }
            }   // operator terms
        }
    }
    double vmax = 0.;
    for (int a = 0; a < NFI; ++a)
        for (int b = 0; b < NFJ; ++b) vmax = fmax(vmax, fabs(val[a][b]));
    return vmax > 1e-300 ? (float)(.5 * log(vmax)) : -700.f;
}

// one thread per pair of a list of one (li, lj, nprim_i, nprim_j) class
template <int LI, int LJ>
__global__ void qcond_kernel(float *q_out, const uint32_t *pairs, int npairs,
                             const int *bas, const double *env, int nbas,
                             int nprim_i, int nprim_j,
                             const double *cheb, const double *large)
{
    int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= npairs) return;
    q_out[n] = qcond_pair<LI, LJ>(pairs[n], bas, env, nbas, nprim_i, nprim_j, cheb, large, 1., 0., 0.);
}

// Batched form behind the reference-signature entry int2e_qcond_estimator
// (nr_sr_estimator.cu:99-131, 416-430): block b owns the pairs
// [shl_pair_offsets[b], shl_pair_offsets[b+1]) of the concatenated list, all of one
// (l_i, l_j, nprim_i, nprim_j) class which is read from the batch's first pair, as the
// reference kernel does.  The <LI, LJ> instantiation serves the batches of its class only.
template <int LI, int LJ>
__global__ void qcond_batch_kernel(float *q_out, const uint32_t *bas_ij_idx, const int *shl_pair_offsets,
                                   const int *bas, const double *env, int nbas,
                                   const double *cheb, const double *large,
                                   double c_full, double c_lr, double omega)
{
    const int sp0 = shl_pair_offsets[blockIdx.x];
    const int sp1 = shl_pair_offsets[blockIdx.x + 1];
    if (sp0 >= sp1) return;
    const uint32_t first = bas_ij_idx[sp0];
    const int ish0 = first / nbas;
    const int jsh0 = first - ish0 * nbas;
    const int l0 = bas[ish0 * BAS_SLOTS + ANG_OF], l1 = bas[jsh0 * BAS_SLOTS + ANG_OF];
    // (ij|ij) is symmetric in i <-> j: a batch with l_i < l_j is evaluated as (ji|ji)
    const bool swap = LI != LJ && l0 == LJ && l1 == LI;
    if (!swap && (l0 != LI || l1 != LJ)) return;
    const int nprim_i = bas[(swap ? jsh0 : ish0) * BAS_SLOTS + NPRIM_OF];
    const int nprim_j = bas[(swap ? ish0 : jsh0) * BAS_SLOTS + NPRIM_OF];
    for (int n = sp0 + threadIdx.x; n < sp1; n += blockDim.x) {
        uint32_t pair = bas_ij_idx[n];
        if (swap) {
            const uint32_t i = pair / nbas, j = pair - i * nbas;
            pair = j * nbas + i;
        }
        q_out[n] = qcond_pair<LI, LJ>(pair, bas, env, nbas, nprim_i, nprim_j, cheb, large, c_full, c_lr, omega);
    }
}

// Overlap prescreen + enumeration of the shell pairs ish >= jsh (reference:
// ovlp_mask_estimation_kernel, lib/pbc/overlap.cu:2163-2223, and PBCsort_pair_ij,
// lib/pbc/supmol_sr_estimator.cu:461-486, which produce a mask that CuPy then compacts).
// Here one kernel does estimate + compaction: every thread owns one (ish, jsh), evaluates the
// reference's FP32 bound from the most diffuse primitives and, if the pair survives, appends
// it to the list of its group pair gp = gi*(gi+1)/2 + gj.  Run twice: pairs == NULL counts
// (counts[gp]), then the fill pass appends at offsets[gp] + atomicAdd(cursor[gp]).  The order
// inside a group pair is arbitrary; the caller sorts by the Schwarz bound anyway.
__global__ void pair_prescreen_kernel(uint32_t *pairs, int *counts, const int *offsets,
                                      const int *bas, const double *env, int nbas,
                                      const float *exps, const float *log_coeff, const int *shl_group,
                                      float log_cutoff)
{
    const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int ish = (int)(n / nbas);
    const int jsh = (int)(n - (long long)ish * nbas);
    if (ish >= nbas || jsh > ish) return;
    const int li = bas[ish * BAS_SLOTS + ANG_OF];
    const int lj = bas[jsh * BAS_SLOTS + ANG_OF];
    const double *ri = env + bas[ish * BAS_SLOTS + PTR_BAS_COORD];
    const double *rj = env + bas[jsh * BAS_SLOTS + PTR_BAS_COORD];
    const float dx = (float)(rj[0] - ri[0]), dy = (float)(rj[1] - ri[1]), dz = (float)(rj[2] - ri[2]);
    const float ai = exps[ish], aj = exps[jsh];
    const float aij = ai + aj;
    const float fi = ai / aij, fj = aj / aij;
    const float theta = ai * fj;
    const float rr = dx * dx + dy * dy + dz * dz;
    if (theta * rr > 50.f) return;                     // REMOTE_THRESHOLD
    const float dr = sqrtf(rr);
    const float dri = fj * dr, drj = fi * dr;
    const float log_ovlp = log_coeff[ish] + log_coeff[jsh] + 1.717f - 1.5f * logf(aij) - theta * rr
                         + .5f * li * logf(.5f * li / aij + dri * dri + 1e-9f)
                         + .5f * lj * logf(.5f * lj / aij + drj * drj + 1e-9f);
    if (!(log_ovlp > log_cutoff)) return;
    const int gi = shl_group[ish], gj = shl_group[jsh];
    const int gp = gi * (gi + 1) / 2 + gj;
    const int pos = atomicAdd(&counts[gp], 1);
    if (pairs != nullptr) pairs[offsets[gp] + pos] = (uint32_t)ish * nbas + jsh;
}

// One thread per shell pair: Gaussian-product data of all primitive pairs, so that
// the J/K kernels read one contiguous record per streamed pair instead of chasing
// bas/env pointers (replaces nothing in the reference, whose kernels recompute
// these per quartet, rys_contract_jk.cu:146-160; it plays the role the reference
// reserves for bas_pair_cache extensions).
__global__ void pair_record_kernel(double *out, const uint32_t *pairs, int npairs,
                                   const int *bas, const double *env, const int *ao_loc, int nbas,
                                   int nprim_i, int nprim_j)
{
    int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= npairs) return;
    const uint32_t pij = pairs[n];
    const int ish = pij / nbas;
    const int jsh = pij - ish * nbas;
    const double *Ri = env + bas[ish * BAS_SLOTS + PTR_BAS_COORD];
    const double *Rj = env + bas[jsh * BAS_SLOTS + PTR_BAS_COORD];
    const double *ei = env + bas[ish * BAS_SLOTS + PTR_EXP];
    const double *ej = env + bas[jsh * BAS_SLOTS + PTR_EXP];
    const double *ci = env + bas[ish * BAS_SLOTS + PTR_COEFF];
    const double *cj = env + bas[jsh * BAS_SLOTS + PTR_COEFF];
    double *rec = out + (size_t)n * pair_rec_stride(nprim_i * nprim_j);
    const double dx = Ri[0] - Rj[0], dy = Ri[1] - Rj[1], dz = Ri[2] - Rj[2];
    const double rr2 = dx * dx + dy * dy + dz * dz;
    rec[0] = dx; rec[1] = dy; rec[2] = dz;
    rec[3] = Ri[0]; rec[4] = Ri[1]; rec[5] = Ri[2];
    int *pi = (int *)(rec + 6);
    pi[0] = ao_loc[ish];
    pi[1] = ao_loc[jsh];
    rec[7] = ish == jsh ? .5 : 1.;
    for (int ip = 0; ip < nprim_i; ++ip)
        for (int jp = 0; jp < nprim_j; ++jp) {
            const double ai = ei[ip], aj = ej[jp];
            const double zeta = ai + aj;
            const double f = aj / zeta;
            double *o = rec + PAIR_REC_HDR + (ip * nprim_j + jp) * 5;
            o[0] = zeta;
            o[1] = Ri[0] - f * dx;
            o[2] = Ri[1] - f * dy;
            o[3] = Ri[2] - f * dz;
            o[4] = ci[ip] * cj[jp] * exp(-ai * f * rr2);
        }
}

// Short-range pair estimator (reference: fill_s_estimator_kernel, nr_sr_estimator.cu:40-97, same FP32
// formula): log of a bound on the short-range pair density built from the most diffuse primitives,
// with the distance guess r_guess = 6 / sqrt(a_ij omega^2 / (omega^2 + a_ij)).  When `geom` is given it
// also receives what the quartet filter needs per pair (create_tasks.cu:274-352): the centre of the
// product of the two diffuse primitives and their exponent sum.
__global__ void s_estimator_kernel(float *s_out, float4 *geom, const uint32_t *pairs, int npairs,
                                   const int *bas, const double *env, int nbas,
                                   const float *diffuse_exps, const float *diffuse_ctr_coef, float omega)
{
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= npairs) return;
    const uint32_t pair = pairs[n];
    const int ish = pair / nbas;
    const int jsh = pair - ish * nbas;
    const int li = bas[ish * BAS_SLOTS + ANG_OF];
    const int lj = bas[jsh * BAS_SLOTS + ANG_OF];
    const double *ri = env + bas[ish * BAS_SLOTS + PTR_BAS_COORD];
    const double *rj = env + bas[jsh * BAS_SLOTS + PTR_BAS_COORD];
    const float xi = (float)ri[0], yi = (float)ri[1], zi = (float)ri[2];
    const float xjxi = (float)rj[0] - xi, yjyi = (float)rj[1] - yi, zjzi = (float)rj[2] - zi;
    const float rr_ij = xjxi * xjxi + yjyi * yjyi + zjzi * zjzi;
    const float ai = diffuse_exps[ish], aj = diffuse_exps[jsh];
    const float aij = ai + aj;
    const float aj_aij = aj / aij, ai_aij = ai / aij;
    const float theta_ij = ai * aj / aij;
    const float cicj = diffuse_ctr_coef[ish] * diffuse_ctr_coef[jsh];
    const float omega2 = omega * omega;
    const float fac_guess = .5f - logf(omega2) / 4;
    const float omega_aij = omega2 / (omega2 + aij);
    const float r_guess = 6.f / sqrtf(aij * omega_aij);
    float norm = 1.f;
    if (li >= 2) norm *= (2 * li + 1.f) / (4 * 3.14159265358979f);
    if (lj >= 2) norm *= (2 * lj + 1.f) / (4 * 3.14159265358979f);
    const float log_fac = logf(fabsf(cicj) * sqrtf(norm)) + 1.7171f - 1.5f * logf(aij) + fac_guess;
    const float dri = aj_aij * r_guess, drj = ai_aij * r_guess;
    const float dri_fac = .5f * li * logf(.5f * li / aij + dri * dri + 1e-9f);
    const float drj_fac = .5f * lj * logf(.5f * lj / aij + drj * drj + 1e-9f);
    const float est = dri_fac + drj_fac - theta_ij * rr_ij + log_fac;
    s_out[n] = est > -700.f ? est : -700.f;
    if (geom != nullptr) {
        float4 g;
        g.x = xi + xjxi * aj_aij;
        g.y = yi + yjyi * aj_aij;
        g.z = zi + zjzi * aj_aij;
        g.w = aij;
        geom[n] = g;
    }
}
