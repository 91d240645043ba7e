/*
 * C-ABI of libgvhf_rys_b200.so -- the sm_100a direct-SCF J/K path.
 *
 * Drop-in boundary for the reference's libgvhf_rys entry points used by
 * gpu4pyscf/scf/jk.py (loaded there with numpy.ctypeslib / ctypes,
 * gpu4pyscf/scf/jk.py:42-47).  All pointers are raw DEVICE pointers unless
 * marked HOST; no ownership is transferred; outputs are accumulated (+=) into
 * caller-zeroed buffers; every function returns 0 on success, non-zero on
 * failure (the Python side raises RuntimeError, jk.py:584-585).  Launches are
 * asynchronous on the stream given (or the legacy default stream for the
 * reference-signature entry points, as in the reference).
 */
#ifndef GVHF_RYS_B200_H
#define GVHF_RYS_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Same field order as the reference struct (lib/gvhf-rys/vhf.cuh:47-54,
 * gpu4pyscf/gto/mole.py:1054-1082): a HOST struct holding device pointers. */
typedef struct {
    int natm;
    int nbas;
    int *atm;
    int *bas;
    double *env;
    int *ao_loc;
} RysIntEnvVars;

/* Upload the Rys root/weight tables to the current device (idempotent,
 * per-device).  Called implicitly by the entry points below.
 * Replaces the static __device__ tables of lib/gvhf-rys/rys_roots_dat.cu. */
int RYS_b200_init(void);

/* Highest angular momentum the GPU path evaluates (reference: LMAX = 4,
 * lib/gvhf-rys/vhf.cuh:26, scf/jk.py:49). */
int RYS_b200_lmax(void);

/* Measurement hook: when non-NULL, every J/K launch adds the number of shell
 * quartets it evaluated to *dev_counter (device uint64).  Used by bench.py for
 * the algorithmic-flop count; NULL (default) disables it. */
int B200_set_counter(unsigned long long *dev_counter);

/* Replaces cuda_version() of lib/gvhf-rys/rys_jk_driver.cu:51. */
int cuda_version(void);

/* Reference signature: RYS_build_jk, lib/gvhf-rys/rys_contract_jk.cu:344-454,
 * called from gpu4pyscf/scf/jk.py:564-583.  One launch per (i,j,k,l) group of
 * (l, nprim)-sorted shells; quartets are de-duplicated by
 * pair_ij >= pair_kl (pair id = ish*nbas+jsh); vj/vk receive the partial
 * (un-symmetrised) sums of jk.py:594-616.  pool must hold at least one int;
 * bas and env are HOST copies (only bas is read, for l and nprim).
 * env[8] = omega (HOST, libcint convention as in the reference's rys_roots_rs,
 * lib/gvhf-rys/rys_roots.cu:91-119): 0 full range, > 0 long range erf(omega r)/r,
 * < 0 short range erfc(|omega| r)/r (evaluated as full - long: two launches).
 * s_cond_*, diffuse_exps and shm_size are accepted for signature compatibility
 * and ignored (the full-range q_cond is a valid, looser Schwarz bound). */
int RYS_build_jk(double *vj, double *vk, double *dm, int n_dm, int nao,
                 RysIntEnvVars *envs, int *shls_slice, int shm_size,
                 int npairs_ij, int npairs_kl,
                 uint32_t *pair_ij_mapping, uint32_t *pair_kl_mapping,
                 float *q_cond_ij, float *q_cond_kl,
                 float *s_cond_ij, float *s_cond_kl,
                 float *diffuse_exps, float *dm_cond, float cutoff, float dm_penalty,
                 uint32_t *pool, int *bas, double *env);

/* Reference signature: RYS_build_k, lib/gvhf-rys/rys_contract_k.cu:605-709, called from
 * gpu4pyscf/scf/jk.py:845-877.  K only, operator
 * lr_factor * erf(|omega| r)/r + sr_factor * erfc(|omega| r)/r with the reference's rules
 * (rys_roots_for_k.cu:25-60): omega == 0 -> lr_factor * K[1/r] (so (0,0,0) adds nothing);
 * sr_factor == 0 -> lr_factor * K[erf]; otherwise sr_factor * K[1/r] +
 * (lr_factor - sr_factor) * K[erf] (one or two launches).
 * Same group / pair-list / de-duplication semantics as RYS_build_jk; s_cond_*,
 * diffuse_exps and shm_size are accepted and ignored (the full-range q_cond is a valid
 * Schwarz bound for the attenuated operators). */
int RYS_build_k(double *vk, double *dm, int n_dm, int nao, double omega, double lr_factor, double sr_factor,
                RysIntEnvVars *envs, int *shls_slice, int shm_size, int npairs_ij, int npairs_kl,
                uint32_t *pair_ij_mapping, uint32_t *pair_kl_mapping,
                float *q_cond_ij, float *q_cond_kl, float *s_cond_ij, float *s_cond_kl,
                float *diffuse_exps, float *dm_cond, float cutoff, float dm_penalty,
                uint32_t *pool, int *bas);

/* Native entry: one angular class (pq|rs).
 *   lprim  HOST int[8] = {lp, lq, lr, ls, nprim_p, nprim_q, nprim_r, nprim_s}
 *   X = (r,s) pair list (block-resident side, nf_r*nf_s >= nf_p*nf_q),
 *   Y = (p,q) pair list (streamed); both sorted by q descending.
 *   dedupe 0: all X x Y quartets; 1: X and Y are the same list, keep iy >= ix;
 *          2/3: keep pair_x >= pair_y / pair_y >= pair_x (reference semantics).
 *   q_margin: 0 for exactly sorted q_y, 4 for the reference's binned order.
 *   rank/world: the work items (X pair, chunk of the Y list) with item % world == rank are
 *   processed (multi-GPU partition); head: zeroed device int work counter.
 *   rec_y: optional per-pair records of the Y list (B200_pair_records), NULL to
 *   derive the Gaussian products from bas/env inside the kernel.
 *   omega, scale, ratio2, omega2: the operator, see B200_class_two_terms below (omega = 0 for 1/r, > 0
 *   for the long-range operator erf(omega r)/r, reference rys_roots_rs, lib/gvhf-rys/rys_roots.cu:91-119;
 *   negative values are rejected: the short-range operator is a two-term operator).
 * vj or vk may be NULL (K-only / J-only).  n_dm is arbitrary. */
int B200_build_jk_class(double *vj, double *vk, const double *dm, int n_dm, int nao,
                        const RysIntEnvVars *envs, const int *lprim,
                        const uint32_t *pair_x, const float *q_x, int npair_x,
                        const uint32_t *pair_y, const float *q_y, int npair_y,
                        const float *dm_cond, float cutoff, float dm_penalty, float q_margin,
                        int dedupe, int rank, int world, int *head, void *stream, const double *rec_y,
                        double omega, double scale, double ratio2, double omega2,
                        const float *sr_s_x, const void *sr_g_x, const float *sr_s_y, const void *sr_g_y,
                        double sr_omega);

/* Native entry, row-stream kernels (csrc/rys_jk_rowt.cuh, ROW variant of rys_jk_warp.cuh): one
 * angular class with the streamed list grouped in ROWS of one shell p.
 *   lprim  HOST int[8] = {lp, lq, lr, ls, nprim_p, nprim_q, nprim_r, nprim_s}; (p,q) is the
 *          kernel order of the streamed pair: p the row shell, q its partners (lp >= lq)
 *   X = (r,s) flat pair list sorted by q descending (ids r*nbas+s, r >= s)
 *   Y: pair_y[k] = p*nbas+q, q_y[k], records rec_y (B200_pair_records on pair_y, required),
 *      row_ptr[nrows+1] offsets, row_p[nrows] the shell of each row, row_qmax[nrows] its largest q_y
 *   dmT / vkT: transposed density matrices and the buffer that receives the q-row K blocks
 *      transposed (vkT[k][x][q] += K_k[q][x]); both alias dm / vk for symmetric D
 *   dedupe != 0: the two lists come from the same group pair, a quartet is kept when
 *      canonical(pair_y) >= pair_x (1/2 on equality)
 *   head: zeroed device uint64 work counter; omega as in B200_build_jk_class.
 * The warp-team kernels of this entry serve uncontracted launches only (B200_class_kind). */
int B200_build_jk_rows(double *vj, double *vk, double *vkT, const double *dm, const double *dmT,
                       int n_dm, int nao, const RysIntEnvVars *envs, const int *lprim,
                       const uint32_t *pair_x, const float *q_x, int npair_x,
                       const uint32_t *pair_y, const float *q_y, const double *rec_y,
                       const int *row_ptr, const int *row_p, const float *row_qmax, int nrows,
                       const float *dm_cond, float cutoff, float dm_penalty,
                       int dedupe, int rank, int world, unsigned long long *head, void *stream,
                       double omega, double scale, double ratio2, double omega2,
                       const float *sr_s_x, const void *sr_g_x, const float *sr_s_y, const void *sr_g_y,
                        double sr_omega);

/* Short-range exchange filter of the native launches (sr_* arguments, sr_omega = |omega|; all NULL = off; only honoured by K-only
 * launches): per pair of the X / Y list the estimator s and a float4 (centre, exponent sum) record, both
 * from B200_sr_pair_data.  A quartet is kept only if dm_penalty and its largest exchange density bound
 * exceed  cutoff - s_x - s_y + log(R^2 + 1) + theta R^2  -- the reference's _fill_sr_vk_tasks
 * (create_tasks.cu:245-369), applied, like there, to the pure short-range operator (lr_factor == 0).
 * fill_s_estimator: reference signature (nr_sr_estimator.cu:400-414, scf/jk.py:1095-1105), the estimator only. */
int B200_sr_pair_data(float *s_estimator, void *geom, const RysIntEnvVars *envs, const uint32_t *bas_ij_idx,
                      const float *diffuse_exps, const float *diffuse_ctr_coef, int npairs, double omega,
                      void *stream);
int fill_s_estimator(float *s_estimator, RysIntEnvVars *envs, uint32_t *bas_ij_idx, float *diffuse_exps,
                     float *diffuse_ctr_coef, int npairs, double omega);

/* Operator of a native launch: every integral is  scale * ( O[omega] + ratio2 * O[omega2] )  with O[0] = 1/r and
 * O[w > 0] = erf(w r)/r; ratio2 = 0 is a single term.  The short-range operator erfc(w r)/r is (omega = 0,
 * ratio2 = -1, omega2 = w): the counterpart of the reference's doubled root set (rys_roots_for_k.cu:45-60).
 * B200_class_two_terms: 1 if the kernel family that serves the class takes both terms in one pass (else the
 * caller issues one launch per term). */
int B200_class_two_terms(int lp, int lq, int lr, int ls, int row);

/* 1 if a row-stream kernel is instantiated for rows of (lp,lq) pairs against resident (lr,ls) pairs. */
int B200_has_row_class(int lp, int lq, int lr, int ls);

/* Kernel family that serves a class (row = 0: B200_build_jk_class, row = 1: B200_build_jk_rows):
 * 0 none, 1 thread per quartet, 2 warp team, 3 block slots.  Plays the role of the reference's
 * kernel-selection switch (rys_contract_jk.cu:395-440) for the host-side launch planner. */
int B200_class_kind(int lp, int lq, int lr, int ls, int row);

/* 1 if a kernel is instantiated with (lp,lq) as the lane side (Y) and (lr,ls)
 * as the register-tile side (X); exactly one orientation of every unordered
 * angular class exists.  Plays the role of the reference's kernel-selection
 * switch (rys_contract_jk.cu:395-440). */
int B200_has_class(int lp, int lq, int lr, int ls);

/* Per-pair Gaussian-product records for a pair list of one (nprim_i, nprim_j) class,
 * B200_pair_rec_stride() doubles each, in list order (layout: csrc/common.cuh).  The
 * reference recomputes these per quartet (rys_contract_jk.cu:146-160); precomputing
 * them once per _VHFOpt.build is what its bas_pair_cache exists for (scf/jk.py:1040-1120). */
int B200_pair_rec_stride(int nprim_i, int nprim_j);
int B200_pair_records(double *out, const RysIntEnvVars *envs, const uint32_t *pairs, int npairs,
                      int nprim_i, int nprim_j, void *stream);

/* Schwarz bounds q = 1/2 log max|(ij|ij)| for a list of shell pairs of one
 * (li, lj, nprim_i, nprim_j) class.  Replaces int2e_qcond_estimator,
 * lib/gvhf-rys/nr_sr_estimator.cu:416-430 (called from scf/jk.py:1076-1086). */
int B200_qcond(float *q_out, const RysIntEnvVars *envs, const uint32_t *pairs, int npairs,
               int li, int lj, int nprim_i, int nprim_j, void *stream);

/* Reference signature: int2e_qcond_estimator, lib/gvhf-rys/nr_sr_estimator.cu:416-430, called from
 * gpu4pyscf/scf/jk.py:1075-1088.  q_out[n] = 1/2 log max|(ij|ij)| for the concatenated pair list
 * bas_ij_idx; block b of the launch owns [shl_pair_offsets[b], shl_pair_offsets[b+1]), all of one
 * (l_i, l_j, nprim_i, nprim_j) class (read from the batch's first pair, as in the reference).
 * (omega, lr_factor, sr_factor): operator lr_factor*erf(|omega| r)/r + sr_factor*erfc(|omega| r)/r with
 * the rules of RYS_build_k.  shm_size and gout_stride_lookup are accepted and ignored.  Legacy
 * default stream, like the reference. */
int int2e_qcond_estimator(float *q_out, RysIntEnvVars *envs, int shm_size,
                          int nbatches_shl_pair, uint32_t *bas_ij_idx,
                          int *shl_pair_offsets, int *gout_stride_lookup,
                          double omega, double lr_factor, double sr_factor);

/* Overlap prescreen + enumeration of the shell pairs ish >= jsh in one kernel.  Replaces
 * PBCovlp_mask_estimation (lib/pbc/overlap.cu:2163-2260), PBCsort_pair_ij
 * (lib/pbc/supmol_sr_estimator.cu:461-486) and the CuPy mask compaction of scf/jk.py:1025-1053.
 *   exps, log_coeff  float[nbas]: exponent and log|coefficient| of the most diffuse primitive of each shell
 *   shl_group        int[nbas]: (l, nprim) group of each shell; group pair gp = gi*(gi+1)/2 + gj
 *   counts           zeroed int[ngroups*(ngroups+1)/2]: receives the number of surviving pairs per gp
 *   pairs == NULL    count only; otherwise pair ids ish*nbas+jsh are appended at offsets[gp] + running count
 *   log_cutoff       log(precision^2), as the reference passes (scf/jk.py:1025). */
int B200_pair_prescreen(uint32_t *pairs, int *counts, const int *offsets, const RysIntEnvVars *envs,
                        const float *exps, const float *log_coeff, const int *shl_group,
                        float log_cutoff, void *stream);

/* Reference names of the J-only entry: RYS_build_j / RYS_init_rysj_constant,
 * lib/gvhf-rys/rys_contract_j.cu:824-956, called from gpu4pyscf/scf/jk.py:681-746.  DIFFERENCE: the
 * reference kernel takes density and J blocks folded to the xyz monomials of every shell pair (its
 * host-side cart2xyz layout, described by `scheme`); this entry takes plain (n_dm, nao, nao)
 * matrices in the sorted Cartesian basis and runs the J/K class kernels with a NULL K output, so a
 * caller that switches libraries drops the two folding transforms.  `scheme` is ignored.
 * Everything else (group slices, pair lists, de-duplication, env[8] = omega) as RYS_build_jk. */
int RYS_build_j(double *vj, double *dm, int n_dm, int nao,
                RysIntEnvVars *envs, int *scheme, int *shls_slice,
                int npairs_ij, int npairs_kl, uint32_t *pair_ij_mapping, uint32_t *pair_kl_mapping,
                float *q_cond_ij, float *q_cond_kl, float *s_cond_ij, float *s_cond_kl,
                float *diffuse_exps, float *dm_cond, float cutoff, float dm_penalty,
                uint32_t *pool, int *bas, double *env);
int RYS_init_rysj_constant(void);

/* Sparse AO transform  out (n, nrow, nrow) = A in A^T  for a CSR matrix A (nrow x ncol_in; row_ptr[nrow+1],
 * col, val) and n square inputs (ncol_in x ncol_in); tmp = scratch of n*nrow*ncol_in doubles.  A = C gives
 * D_sorted = C D C^T, A = C^T gives V = C^T V_sorted C, where C (nao_sorted x nao) is the recontraction x
 * cart->sph map of the sorted basis.  Replaces the bra_/ket_ {sph,cart}2sorted / sorted2{sph,cart} entries
 * of lib/gvhf-rys/mole_helper.cu:1735-1908 (gto/mole.py:790-938) with one entry on a CSR data model. */
int B200_sparse_A_mat_AT(double *out, double *tmp, const double *in, int n, int nrow, int ncol_in,
                         const int *row_ptr, const int *col, const double *val, void *stream);

/* dm_cond[i*nbas+j] = log(max |D_n[a,b]| + 1e-300) over the AO block of shell
 * pair (i,j) and all n_dm matrices.  Replaces the CuPy RawKernel behind
 * condense('absmax', ...) (gpu4pyscf/lib/cupy_helper.py:1068-1184, used at
 * scf/jk.py:510-514). */
int B200_condense(float *out, const double *dm, int n_dm, int nao, const int *ao_loc, int nbas,
                  void *stream);

/* Measurement only: FP64 FMA burn (8 independent chains per thread, `iters`
 * rounds; 16 flops per thread and round) -- the DFMA-peak denominator of the
 * roofline is timed with CUDA events around this launch (bench.py). */
int B200_dfma_burn(double *sink, int iters, int blocks, int threads, void *stream);

#ifdef __cplusplus
}
#endif
#endif
