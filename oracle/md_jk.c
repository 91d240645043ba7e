/*
 * ORACLE -- TEST INFRASTRUCTURE ONLY.  Never imported by the product path.
 *
 * CPU restatement of what PySCF's CPU get_jk returns for the reference's J/K
 * plugin (gpu4pyscf/scf/jk.py:60-151 compares against pyscf.scf.hf.get_jk in
 * gpu4pyscf/scf/tests/test_scf_jk.py:22,45 and test_rhf.py:166-194).
 * PySCF's libcvhf/libcint (pyscf==2.14.0, requirements.txt:4) are third-party
 * and absent from /root/reference and from this image, so the algorithm is
 * restated from its published definition:
 *   J_ij = sum_kl (ij|kl) D_lk ,  K_il = sum_jk (ij|kl) D_jk      
 * with libcint's Cartesian conventions (angular factor only for s and p;
 * reference mirror gto/mole.py:1120-1133).  The ERIs are evaluated with the
 * McMurchie-Davidson scheme (Hermite expansion + Boys function) -- deliberately
 * NOT the Rys quadrature the CUDA path uses, so that a table or recurrence bug
 * in the product cannot cancel in the comparison.
 *
 * Pinned against the reference's golden fingerprints in tests/test_oracle.py
 * (test_rhf.py:166-194: fp(vj)=179.14526555375858, fp(vk)=-34.851182918643005).
 *
 * 8-fold symmetry is used over shell quartets; optional Schwarz screening
 * (libcvhf CVHFnr_int2e_q_cond style) for the timed CPU baseline.
 */
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <stdio.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ATOM_OF 0
#define ANG_OF 1
#define NPRIM_OF 2
#define NCTR_OF 3
#define PTR_EXP 5
#define PTR_COEFF 6
#define BAS_SLOTS 8
#define ATM_SLOTS 6
#define PTR_COORD 1

#define LMAX_O 6
#define L4MAX (4*LMAX_O)
#define NCART_MAX ((LMAX_O+1)*(LMAX_O+2)/2)

static const double SQRTPI = 1.7724538509055160273;

/* Boys function F_0..F_m(x) */
static void boys(double *f, int m, double x)
{
    if (x < 35. + m) {
        /* series for F_m, then downward recursion */
        double ex = exp(-x);
        double term = 1. / (2 * m + 1);
        double sum = term;
        for (int k = 1; k < 400; k++) {
            term *= 2 * x / (2 * m + 2 * k + 1);
            sum += term;
            if (term < 1e-18 * sum) break;
        }
        f[m] = ex * sum;
        for (int n = m; n > 0; n--) {
            f[n - 1] = (2 * x * f[n] + ex) / (2 * n - 1);
        }
    } else {
        double ex = exp(-x);
        f[0] = 0.5 * SQRTPI / sqrt(x) * erf(sqrt(x));
        for (int n = 0; n < m; n++) {
            f[n + 1] = ((2 * n + 1) * f[n] - ex) / (2 * x);
        }
    }
}

/* Hermite expansion coefficients E[i][j][t], i<=la, j<=lb, t<=i+j (no exp prefactor) */
static void hermite_E(double *E, int la, int lb, double p, double PA, double PB)
{
    int nb = lb + 1, nt = la + lb + 2;
#define EE(i,j,t) E[((i)*nb+(j))*nt+(t)]
    memset(E, 0, sizeof(double) * (la + 1) * nb * nt);
    double h = 0.5 / p;
    EE(0,0,0) = 1.;
    for (int i = 0; i < la; i++) {
        for (int t = 0; t <= i + 1; t++) {
            double v = PA * EE(i,0,t) + (t + 1) * EE(i,0,t+1);
            if (t > 0) v += h * EE(i,0,t-1);
            EE(i+1,0,t) = v;
        }
    }
    for (int i = 0; i <= la; i++) {
        for (int j = 0; j < lb; j++) {
            for (int t = 0; t <= i + j + 1; t++) {
                double v = PB * EE(i,j,t) + (t + 1) * EE(i,j,t+1);
                if (t > 0) v += h * EE(i,j,t-1);
                EE(i,j+1,t) = v;
            }
        }
    }
#undef EE
}

static void cart_comp(int l, int *cx, int *cy, int *cz)
{
    int n = 0;
    for (int lx = l; lx >= 0; lx--)
        for (int ly = l - lx; ly >= 0; ly--) {
            cx[n] = lx; cy[n] = ly; cz[n] = l - lx - ly; n++;
        }
}

static double ang_fac(int l)
{
    if (l == 0) return 0.28209479177387814;
    if (l == 1) return 0.48860251190291992;
    return 1.;
}

typedef struct {
    int l, nprim, nctr, nf;
    const double *exps, *coef; /* coef[ctr*nprim + prim] */
    const double *r;
} Shell;

static Shell get_shell(const int *atm, const int *bas, const double *env, int ib)
{
    Shell s;
    const int *b = bas + ib * BAS_SLOTS;
    s.l = b[ANG_OF]; s.nprim = b[NPRIM_OF]; s.nctr = b[NCTR_OF];
    s.nf = (s.l + 1) * (s.l + 2) / 2;
    s.exps = env + b[PTR_EXP];
    s.coef = env + b[PTR_COEFF];
    s.r = env + atm[b[ATOM_OF] * ATM_SLOTS + PTR_COORD];
    return s;
}

/*
 * Contracted Cartesian ERI block (ij|kl), layout
 *   out[(((ic*nfi+i)*dj + (jc*nfj+j))*dk + (kc*nfk+k))*dl + (lc*nfl+l)],
 * i.e. plain 4-index array over the shell's (nctr*nf) functions.
 * omega = 0: 1/r; omega > 0: erf(omega r)/r; omega < 0: erfc(|omega| r)/r.
 */
static void eri_block(double *out, Shell si, Shell sj, Shell sk, Shell sl, double omega)
{
    int li = si.l, lj = sj.l, lk = sk.l, ll = sl.l;
    int L = li + lj + lk + ll;
    int nfi = si.nf, nfj = sj.nf, nfk = sk.nf, nfl = sl.nf;
    int di = nfi * si.nctr, dj = nfj * sj.nctr, dk = nfk * sk.nctr, dl = nfl * sl.nctr;
    size_t nout = (size_t)di * dj * dk * dl;
    memset(out, 0, sizeof(double) * nout);

    int ix[NCART_MAX], iy[NCART_MAX], iz[NCART_MAX];
    int jx[NCART_MAX], jy[NCART_MAX], jz[NCART_MAX];
    int kx[NCART_MAX], ky[NCART_MAX], kz[NCART_MAX];
    int lx[NCART_MAX], ly[NCART_MAX], lz[NCART_MAX];
    cart_comp(li, ix, iy, iz); cart_comp(lj, jx, jy, jz);
    cart_comp(lk, kx, ky, kz); cart_comp(ll, lx, ly, lz);

    int nR = L + 1;
    /* Rn[n][t][u][v] */
    double *Rn = (double *)malloc(sizeof(double) * nR * nR * nR * nR);
#define R(n,t,u,v) Rn[(((n)*nR+(t))*nR+(u))*nR+(v)]
    int lab = li + lj, lcd = lk + ll;
    int nhb = lab + 1, nhk = lcd + 1;
    double *Exb = (double *)malloc(sizeof(double) * 3 * (li + 1) * (lj + 1) * (lab + 2));
    double *Eyb = Exb + (li + 1) * (lj + 1) * (lab + 2);
    double *Ezb = Eyb + (li + 1) * (lj + 1) * (lab + 2);
    double *Exk = (double *)malloc(sizeof(double) * 3 * (lk + 1) * (ll + 1) * (lcd + 2));
    double *Eyk = Exk + (lk + 1) * (ll + 1) * (lcd + 2);
    double *Ezk = Eyk + (lk + 1) * (ll + 1) * (lcd + 2);
    /* I[tuv][kl comp] */
    double *Imid = (double *)malloc(sizeof(double) * nhb * nhb * nhb * nfk * nfl);
    double *prim = (double *)malloc(sizeof(double) * nfi * nfj * nfk * nfl);
    double fboys[L4MAX + 2];
    double angf = ang_fac(li) * ang_fac(lj) * ang_fac(lk) * ang_fac(ll);

    double ABx = si.r[0] - sj.r[0], ABy = si.r[1] - sj.r[1], ABz = si.r[2] - sj.r[2];
    double CDx = sk.r[0] - sl.r[0], CDy = sk.r[1] - sl.r[1], CDz = sk.r[2] - sl.r[2];
    double rab2 = ABx * ABx + ABy * ABy + ABz * ABz;
    double rcd2 = CDx * CDx + CDy * CDy + CDz * CDz;

    for (int ip = 0; ip < si.nprim; ip++)
    for (int jp = 0; jp < sj.nprim; jp++) {
        double a = si.exps[ip], b = sj.exps[jp];
        double p = a + b;
        double Px = (a * si.r[0] + b * sj.r[0]) / p;
        double Py = (a * si.r[1] + b * sj.r[1]) / p;
        double Pz = (a * si.r[2] + b * sj.r[2]) / p;
        double Kab = exp(-a * b / p * rab2);
        hermite_E(Exb, li, lj, p, Px - si.r[0], Px - sj.r[0]);
        hermite_E(Eyb, li, lj, p, Py - si.r[1], Py - sj.r[1]);
        hermite_E(Ezb, li, lj, p, Pz - si.r[2], Pz - sj.r[2]);
        for (int kp = 0; kp < sk.nprim; kp++)
        for (int lp = 0; lp < sl.nprim; lp++) {
            double c = sk.exps[kp], d = sl.exps[lp];
            double q = c + d;
            double Qx = (c * sk.r[0] + d * sl.r[0]) / q;
            double Qy = (c * sk.r[1] + d * sl.r[1]) / q;
            double Qz = (c * sk.r[2] + d * sl.r[2]) / q;
            double Kcd = exp(-c * d / q * rcd2);
            hermite_E(Exk, lk, ll, q, Qx - sk.r[0], Qx - sl.r[0]);
            hermite_E(Eyk, lk, ll, q, Qy - sk.r[1], Qy - sl.r[1]);
            hermite_E(Ezk, lk, ll, q, Qz - sk.r[2], Qz - sl.r[2]);
            double alpha = p * q / (p + q);
            double X = Px - Qx, Y = Py - Qy, Z = Pz - Qz;
            double rpq2 = X * X + Y * Y + Z * Z;
            double pref = 2 * pow(M_PI, 2.5) / (p * q * sqrt(p + q)) * Kab * Kcd * angf;

            /* one or two Boys evaluations: full-range and/or long-range */
            memset(prim, 0, sizeof(double) * nfi * nfj * nfk * nfl);
            for (int pass = 0; pass < 2; pass++) {
                double al = alpha, scale = 1.;
                if (pass == 0) {
                    if (omega > 0) continue;      /* LR only: skip full-range */
                } else {
                    if (omega == 0) continue;
                    double w2 = omega * omega;
                    al = alpha * w2 / (alpha + w2);
                    scale = sqrt(w2 / (alpha + w2));
                    if (omega < 0) scale = -scale;  /* SR = full - LR */
                }
                boys(fboys, L, al * rpq2);
                double m2a = 1.;
                for (int n = 0; n <= L; n++) {
                    R(n,0,0,0) = m2a * fboys[n] * scale;
                    m2a *= -2 * al;
                }
                /* build R^n_{tuv} downward in n */
                for (int n = L - 1; n >= 0; n--) {
                    int M = L - n;
                    for (int t = 0; t <= M; t++)
                    for (int u = 0; u <= M - t; u++)
                    for (int v = 0; v <= M - t - u; v++) {
                        if (t + u + v == 0) continue;
                        double val;
                        if (t > 0) {
                            val = X * R(n+1,t-1,u,v);
                            if (t > 1) val += (t - 1) * R(n+1,t-2,u,v);
                        } else if (u > 0) {
                            val = Y * R(n+1,t,u-1,v);
                            if (u > 1) val += (u - 1) * R(n+1,t,u-2,v);
                        } else {
                            val = Z * R(n+1,t,u,v-1);
                            if (v > 1) val += (v - 1) * R(n+1,t,u,v-2);
                        }
                        R(n,t,u,v) = val;
                    }
                }
                /* ket contraction: I[tuv][kl] = sum_{tau,nu,phi} (-1)^{..} Ecd R[t+tau,u+nu,v+phi] */
                int ntk = lcd + 2, nlk = ll + 1;
                for (int t = 0; t <= lab; t++)
                for (int u = 0; u <= lab - t; u++)
                for (int v = 0; v <= lab - t - u; v++) {
                    double *Irow = Imid + ((t * nhb + u) * nhb + v) * nfk * nfl;
                    for (int kc = 0; kc < nfk; kc++)
                    for (int lc = 0; lc < nfl; lc++) {
                        const double *ex = Exk + (kx[kc] * nlk + lx[lc]) * ntk;
                        const double *ey = Eyk + (ky[kc] * nlk + ly[lc]) * ntk;
                        const double *ez = Ezk + (kz[kc] * nlk + lz[lc]) * ntk;
                        int tx = kx[kc] + lx[lc], ty = ky[kc] + ly[lc], tz = kz[kc] + lz[lc];
                        double s = 0;
                        for (int a1 = 0; a1 <= tx; a1++) {
                            double sy = 0;
                            for (int b1 = 0; b1 <= ty; b1++) {
                                double sz = 0;
                                for (int c1 = 0; c1 <= tz; c1++) {
                                    double sg = ((a1 + b1 + c1) & 1) ? -1. : 1.;
                                    sz += sg * ez[c1] * R(0,t+a1,u+b1,v+c1);
                                }
                                sy += ey[b1] * sz;
                            }
                            s += ex[a1] * sy;
                        }
                        Irow[kc * nfl + lc] = s;
                    }
                }
                /* bra contraction */
                int ntb = lab + 2, nlb = lj + 1;
                for (int ic = 0; ic < nfi; ic++)
                for (int jc = 0; jc < nfj; jc++) {
                    const double *ex = Exb + (ix[ic] * nlb + jx[jc]) * ntb;
                    const double *ey = Eyb + (iy[ic] * nlb + jy[jc]) * ntb;
                    const double *ez = Ezb + (iz[ic] * nlb + jz[jc]) * ntb;
                    int tx = ix[ic] + jx[jc], ty = iy[ic] + jy[jc], tz = iz[ic] + jz[jc];
                    double *po = prim + (ic * nfj + jc) * nfk * nfl;
                    for (int t = 0; t <= tx; t++)
                    for (int u = 0; u <= ty; u++)
                    for (int v = 0; v <= tz; v++) {
                        double e = ex[t] * ey[u] * ez[v];
                        const double *Irow = Imid + ((t * nhb + u) * nhb + v) * nfk * nfl;
                        for (int n = 0; n < nfk * nfl; n++) po[n] += e * Irow[n];
                    }
                }
            }
            /* scatter into contracted block */
            for (int ci = 0; ci < si.nctr; ci++)
            for (int cj = 0; cj < sj.nctr; cj++)
            for (int ck = 0; ck < sk.nctr; ck++)
            for (int cl = 0; cl < sl.nctr; cl++) {
                double cc = pref * si.coef[ci * si.nprim + ip] * sj.coef[cj * sj.nprim + jp]
                          * sk.coef[ck * sk.nprim + kp] * sl.coef[cl * sl.nprim + lp];
                if (cc == 0) continue;
                for (int ic = 0; ic < nfi; ic++)
                for (int jc = 0; jc < nfj; jc++)
                for (int kc = 0; kc < nfk; kc++) {
                    double *po = out + ((((size_t)(ci * nfi + ic)) * dj + (cj * nfj + jc)) * dk + (ck * nfk + kc)) * dl + cl * nfl;
                    const double *pp = prim + ((ic * nfj + jc) * nfk + kc) * nfl;
                    for (int lc = 0; lc < nfl; lc++) po[lc] += cc * pp[lc];
                }
            }
        }
    }
#undef R
    free(Rn); free(Exb); free(Exk); free(Imid); free(prim);
}

static int *make_ao_loc(const int *bas, int nbas)
{
    int *loc = (int *)malloc(sizeof(int) * (nbas + 1));
    loc[0] = 0;
    for (int i = 0; i < nbas; i++) {
        int l = bas[i * BAS_SLOTS + ANG_OF];
        loc[i + 1] = loc[i] + (l + 1) * (l + 2) / 2 * bas[i * BAS_SLOTS + NCTR_OF];
    }
    return loc;
}

/* sqrt(max |(ij|ij)|) per shell pair -- libcvhf CVHFnr_int2e_q_cond restated */
void oracle_q_cond(double *q, const int *atm, int natm, const int *bas, int nbas,
                   const double *env, double omega)
{
#pragma omp parallel
    {
        double *buf = NULL; size_t cap = 0;
#pragma omp for schedule(dynamic, 4)
        for (int ij = 0; ij < nbas * (nbas + 1) / 2; ij++) {
            int i = (int)((sqrt(8. * ij + 1) - 1) / 2);
            while (i * (i + 1) / 2 > ij) i--;
            while ((i + 1) * (i + 2) / 2 <= ij) i++;
            int j = ij - i * (i + 1) / 2;
            Shell si = get_shell(atm, bas, env, i), sj = get_shell(atm, bas, env, j);
            int di = si.nf * si.nctr, dj = sj.nf * sj.nctr;
            size_t n = (size_t)di * dj * di * dj;
            if (n > cap) { free(buf); buf = (double *)malloc(sizeof(double) * n); cap = n; }
            eri_block(buf, si, sj, si, sj, omega);
            double m = 0;
            for (int a = 0; a < di; a++)
            for (int b = 0; b < dj; b++) {
                double v = fabs(buf[(((size_t)a * dj + b) * di + a) * dj + b]);
                if (v > m) m = v;
            }
            q[i * nbas + j] = q[j * nbas + i] = sqrt(m);
        }
        free(buf);
    }
}

/*
 * J/K for n_dm general (not necessarily symmetric) Cartesian density matrices.
 *   vj[n,i,j] = sum_kl (ij|kl) dm[n,l,k];  vk[n,i,l] = sum_jk (ij|kl) dm[n,j,k]
 * screen != 0: skip shell quartets with q_ij*q_kl*max|D| < tol (libcvhf direct-SCF
 * prescreening, CVHFnrs8_prescreen restated); returns number of quartets computed.
 */
/* Algorithmic flop model of SURVEY.md section 8d (same formula as
 * gpu4pyscf_b200/scf/flops.py) accumulated over the quartets oracle_jk computes,
 * so that a CPU timing on a fragment can be scaled to the full workload. */
static double g_last_flops = 0.;
double oracle_last_flops(void) { return g_last_flops; }
static int nf_l(int l) { return (l + 1) * (l + 2) / 2; }
static double falg_quartet(int li, int lj, int lk, int ll, double nprim, int n_dm)
{
    int lij = li + lj, lkl = lk + ll;
    int nroots = (lij + lkl) / 2 + 1;
    double nfq = (double)nf_l(li) * nf_l(lj) * nf_l(lk) * nf_l(ll);
    double f2d = 4. * lij + 6. * (lij + 1) * lkl;
    for (int j = 0; j < lj; j++) f2d += 2. * (lkl + 1) * (lij - j);
    for (int l = 0; l < ll; l++) f2d += 2. * (li + 1) * (lj + 1) * (lkl - l);
    double fprim = 40. + 76. * nroots + nroots * (24. + 3. * f2d + 3. * nfq);
    return nprim * fprim + 12. * nfq * n_dm;
}

long oracle_jk(double *vj, double *vk, const double *dm, int n_dm, int nao,
               const int *atm, int natm, const int *bas, int nbas, const double *env,
               double omega, int screen, double tol, int with_j, int with_k)
{
    int *ao_loc = make_ao_loc(bas, nbas);
    if (ao_loc[nbas] != nao) { fprintf(stderr, "oracle_jk: nao mismatch %d vs %d\n", ao_loc[nbas], nao); free(ao_loc); return -1; }
    size_t nao2 = (size_t)nao * nao;
    double *q = NULL, *dmax = NULL;
    if (screen) {
        q = (double *)malloc(sizeof(double) * nbas * nbas);
        oracle_q_cond(q, atm, natm, bas, nbas, env, omega);
        dmax = (double *)calloc((size_t)nbas * nbas, sizeof(double));
        for (int n = 0; n < n_dm; n++)
        for (int i = 0; i < nbas; i++)
        for (int j = 0; j < nbas; j++) {
            double m = dmax[i * nbas + j];
            for (int a = ao_loc[i]; a < ao_loc[i + 1]; a++)
            for (int b = ao_loc[j]; b < ao_loc[j + 1]; b++) {
                double v = fabs(dm[n * nao2 + (size_t)a * nao + b]);
                double w = fabs(dm[n * nao2 + (size_t)b * nao + a]);
                if (v > m) m = v;
                if (w > m) m = w;
            }
            dmax[i * nbas + j] = m;
        }
    }
    memset(vj, 0, sizeof(double) * nao2 * n_dm);
    memset(vk, 0, sizeof(double) * nao2 * n_dm);
    long ncomputed = 0;
    double flops = 0.;
    int npair = nbas * (nbas + 1) / 2;
#pragma omp parallel reduction(+:ncomputed,flops)
    {
        double *tj = (double *)calloc(nao2 * n_dm, sizeof(double));
        double *tk = (double *)calloc(nao2 * n_dm, sizeof(double));
        double *buf = NULL; size_t cap = 0;
#pragma omp for schedule(dynamic, 1)
        for (int ijr = 0; ijr < npair; ijr++) {
            int ij = npair - 1 - ijr;   /* big pairs first */
            int i = (int)((sqrt(8. * ij + 1) - 1) / 2);
            while (i * (i + 1) / 2 > ij) i--;
            while ((i + 1) * (i + 2) / 2 <= ij) i++;
            int j = ij - i * (i + 1) / 2;
            Shell si = get_shell(atm, bas, env, i), sj = get_shell(atm, bas, env, j);
            for (int kl = 0; kl <= ij; kl++) {
                int k = (int)((sqrt(8. * kl + 1) - 1) / 2);
                while (k * (k + 1) / 2 > kl) k--;
                while ((k + 1) * (k + 2) / 2 <= kl) k++;
                int l = kl - k * (k + 1) / 2;
                if (screen) {
                    double qq = q[i * nbas + j] * q[k * nbas + l];
                    if (qq < tol) continue;
                    double dd = dmax[i*nbas+j];
                    double t;
                    t = dmax[k*nbas+l]; if (t > dd) dd = t;
                    t = dmax[i*nbas+k]; if (t > dd) dd = t;
                    t = dmax[i*nbas+l]; if (t > dd) dd = t;
                    t = dmax[j*nbas+k]; if (t > dd) dd = t;
                    t = dmax[j*nbas+l]; if (t > dd) dd = t;
                    if (qq * dd < tol) continue;
                }
                Shell sk = get_shell(atm, bas, env, k), sl = get_shell(atm, bas, env, l);
                int di = si.nf * si.nctr, dj = sj.nf * sj.nctr, dk = sk.nf * sk.nctr, dl = sl.nf * sl.nctr;
                size_t n = (size_t)di * dj * dk * dl;
                if (n > cap) { free(buf); buf = (double *)malloc(sizeof(double) * n); cap = n; }
                eri_block(buf, si, sj, sk, sl, omega);
                ncomputed++;
                {   /* class in the kernels' canonical order: higher l first within a pair */
                    int a = bas[i*8+1], b = bas[j*8+1], c = bas[k*8+1], d = bas[l*8+1];
                    double np = (double)bas[i*8+2] * bas[j*8+2] * bas[k*8+2] * bas[l*8+2]
                              * bas[i*8+3] * bas[j*8+3] * bas[k*8+3] * bas[l*8+3];
                    flops += falg_quartet(a > b ? a : b, a > b ? b : a, c > d ? c : d, c > d ? d : c, np, n_dm);
                }
                /* the (up to) 8 distinct index permutations of this quartet */
                int perm[8][4] = {{0,1,2,3},{1,0,2,3},{0,1,3,2},{1,0,3,2},
                                  {2,3,0,1},{3,2,0,1},{2,3,1,0},{3,2,1,0}};
                int sh[4] = {i, j, k, l};
                int dims[4] = {di, dj, dk, dl};
                size_t strd[4] = {(size_t)dj * dk * dl, (size_t)dk * dl, (size_t)dl, 1};
                int seen[8][4]; int nseen = 0;
                for (int pi = 0; pi < 8; pi++) {
                    int *pm = perm[pi];
                    int a = sh[pm[0]], b = sh[pm[1]], c = sh[pm[2]], d = sh[pm[3]];
                    int dup = 0;
                    for (int s = 0; s < nseen; s++)
                        if (seen[s][0] == a && seen[s][1] == b && seen[s][2] == c && seen[s][3] == d) dup = 1;
                    if (dup) continue;
                    seen[nseen][0] = a; seen[nseen][1] = b; seen[nseen][2] = c; seen[nseen][3] = d; nseen++;
                    int a0 = ao_loc[a], b0 = ao_loc[b], c0 = ao_loc[c], d0 = ao_loc[d];
                    int na = dims[pm[0]], nb = dims[pm[1]], nc = dims[pm[2]], nd = dims[pm[3]];
                    size_t sa = strd[pm[0]], sb = strd[pm[1]], sc = strd[pm[2]], sd = strd[pm[3]];
                    for (int m = 0; m < n_dm; m++) {
                        const double *D = dm + m * nao2;
                        double *J = tj + m * nao2, *K = tk + m * nao2;
                        for (int ia = 0; ia < na; ia++)
                        for (int ib = 0; ib < nb; ib++)
                        for (int ic = 0; ic < nc; ic++)
                        for (int id = 0; id < nd; id++) {
                            double g = buf[ia * sa + ib * sb + ic * sc + id * sd];
                            if (with_j) J[(size_t)(a0 + ia) * nao + b0 + ib] += g * D[(size_t)(d0 + id) * nao + c0 + ic];
                            if (with_k) K[(size_t)(a0 + ia) * nao + d0 + id] += g * D[(size_t)(b0 + ib) * nao + c0 + ic];
                        }
                    }
                }
            }
        }
#pragma omp critical
        {
            for (size_t n = 0; n < nao2 * n_dm; n++) { vj[n] += tj[n]; vk[n] += tk[n]; }
        }
        free(tj); free(tk); free(buf);
    }
    free(ao_loc); free(q); free(dmax);
    g_last_flops = flops;
    return ncomputed;
}

/* full Cartesian ERI tensor for tiny systems (debug / unit tests) */
void oracle_eri_shell(double *out, const int *atm, const int *bas, const double *env,
                      int i, int j, int k, int l, double omega)
{
    eri_block(out, get_shell(atm, bas, env, i), get_shell(atm, bas, env, j),
              get_shell(atm, bas, env, k), get_shell(atm, bas, env, l), omega);
}

/* ------------------------------------------------------------------------------------------
 * One-electron integrals (overlap S, kinetic T, nuclear attraction V) in the Cartesian AO
 * basis, McMurchie-Davidson.  Needed only by the TEST-side SCF loop (tests/scf_driver.py) that
 * checks E_SCF of a GPU J/K build against the CPU one; restates what PySCF's
 * mol.intor('int1e_ovlp' / 'int1e_kin' / 'int1e_nuc') returns (libcint conventions as above).
 * ------------------------------------------------------------------------------------------ */

/* Hermite Coulomb integrals R^0_{tuv}(p, PC) for t+u+v <= L */
static void hermite_R(double *R, int L, double p, double X, double Y, double Z)
{
    int n1 = L + 1;
    double f[4 * LMAX_O + 2];
    boys(f, L, p * (X * X + Y * Y + Z * Z));
    /* work[n][t][u][v] */
    double *w = (double *)calloc((size_t)n1 * n1 * n1 * n1, sizeof(double));
#define W(n,t,u,v) w[(((n)*n1+(t))*n1+(u))*n1+(v)]
    double m2p = 1.;
    for (int n = 0; n <= L; n++) { W(n,0,0,0) = m2p * f[n]; m2p *= -2 * p; }
    for (int n = L - 1; n >= 0; n--)
        for (int t = 0; t + n <= L; t++)
            for (int u = 0; t + u + n <= L; u++)
                for (int v = 0; t + u + v + n <= L; v++) {
                    if (t + u + v == 0) continue;
                    double val;
                    if (t > 0) val = X * W(n+1,t-1,u,v) + (t > 1 ? (t - 1) * W(n+1,t-2,u,v) : 0.);
                    else if (u > 0) val = Y * W(n+1,t,u-1,v) + (u > 1 ? (u - 1) * W(n+1,t,u-2,v) : 0.);
                    else val = Z * W(n+1,t,u,v-1) + (v > 1 ? (v - 1) * W(n+1,t,u,v-2) : 0.);
                    W(n,t,u,v) = val;
                }
    for (int t = 0; t <= L; t++)
        for (int u = 0; t + u <= L; u++)
            for (int v = 0; t + u + v <= L; v++) R[(t * n1 + u) * n1 + v] = W(0,t,u,v);
#undef W
    free(w);
}

/* s, t, v: [nao, nao] Cartesian AO matrices (zeroed by the caller) */
void oracle_int1e(double *s, double *t, double *v, int nao,
                  const int *atm, int natm, const int *bas, int nbas, const double *env)
{
    int *ao_loc = make_ao_loc(bas, nbas);
    for (int ish = 0; ish < nbas; ish++)
    for (int jsh = 0; jsh < nbas; jsh++) {
        Shell si = get_shell(atm, bas, env, ish), sj = get_shell(atm, bas, env, jsh);
        int li = si.l, lj = sj.l;
        int cxi[NCART_MAX], cyi[NCART_MAX], czi[NCART_MAX], cxj[NCART_MAX], cyj[NCART_MAX], czj[NCART_MAX];
        cart_comp(li, cxi, cyi, czi);
        cart_comp(lj, cxj, cyj, czj);
        double AB[3] = {si.r[0] - sj.r[0], si.r[1] - sj.r[1], si.r[2] - sj.r[2]};
        double rr = AB[0] * AB[0] + AB[1] * AB[1] + AB[2] * AB[2];
        int nbj = lj + 3, ntj = li + lj + 4;        /* E arrays up to lj + 2 for the kinetic term */
        double *E[3];
        for (int x = 0; x < 3; x++) E[x] = (double *)malloc(sizeof(double) * (li + 1) * nbj * ntj);
        int L = li + lj;
        double *R = (double *)malloc(sizeof(double) * (L + 1) * (L + 1) * (L + 1));
        for (int ip = 0; ip < si.nprim; ip++)
        for (int jp = 0; jp < sj.nprim; jp++) {
            double a = si.exps[ip], b = sj.exps[jp], p = a + b;
            double P[3], pref = exp(-a * b / p * rr) * ang_fac(li) * ang_fac(lj);
            for (int x = 0; x < 3; x++) {
                P[x] = (a * si.r[x] + b * sj.r[x]) / p;
                hermite_E(E[x], li, lj + 2, p, P[x] - si.r[x], P[x] - sj.r[x]);
            }
#define EX(x,i,j,tt) E[x][((i)*nbj+(j))*ntj+(tt)]
            double s3 = pow(M_PI / p, 1.5);
            for (int ic = 0; ic < si.nctr; ic++)
            for (int jc = 0; jc < sj.nctr; jc++) {
                double cc = si.coef[ic * si.nprim + ip] * sj.coef[jc * sj.nprim + jp] * pref;
                if (cc == 0.) continue;
                for (int i = 0; i < si.nf; i++)
                for (int j = 0; j < sj.nf; j++) {
                    int ix = cxi[i], iy = cyi[i], iz = czi[i], jx = cxj[j], jy = cyj[j], jz = czj[j];
                    size_t idx = (size_t)(ao_loc[ish] + ic * si.nf + i) * nao + ao_loc[jsh] + jc * sj.nf + j;
                    double sx = EX(0,ix,jx,0), sy = EX(1,iy,jy,0), sz = EX(2,iz,jz,0);
                    s[idx] += cc * s3 * sx * sy * sz;
                    /* kinetic: -1/2 d^2/dx^2 acting on the ket Gaussian */
                    double kx = -2. * b * b * EX(0,ix,jx+2,0) + b * (2 * jx + 1) * sx
                              - (jx > 1 ? .5 * jx * (jx - 1) * EX(0,ix,jx-2,0) : 0.);
                    double ky = -2. * b * b * EX(1,iy,jy+2,0) + b * (2 * jy + 1) * sy
                              - (jy > 1 ? .5 * jy * (jy - 1) * EX(1,iy,jy-2,0) : 0.);
                    double kz = -2. * b * b * EX(2,iz,jz+2,0) + b * (2 * jz + 1) * sz
                              - (jz > 1 ? .5 * jz * (jz - 1) * EX(2,iz,jz-2,0) : 0.);
                    t[idx] += cc * s3 * (kx * sy * sz + sx * ky * sz + sx * sy * kz);
                }
            }
            /* nuclear attraction: -Z_C * 2 pi / p * sum_tuv E_t E_u E_v R_tuv(p, P - C) */
            for (int ia = 0; ia < natm; ia++) {
                double Z = (double)atm[ia * ATM_SLOTS + 0];
                if (Z == 0.) continue;
                const double *C = env + atm[ia * ATM_SLOTS + PTR_COORD];
                hermite_R(R, L, p, P[0] - C[0], P[1] - C[1], P[2] - C[2]);
                int n1 = L + 1;
                for (int ic = 0; ic < si.nctr; ic++)
                for (int jc = 0; jc < sj.nctr; jc++) {
                    double cc = si.coef[ic * si.nprim + ip] * sj.coef[jc * sj.nprim + jp] * pref;
                    if (cc == 0.) continue;
                    for (int i = 0; i < si.nf; i++)
                    for (int j = 0; j < sj.nf; j++) {
                        int ix = cxi[i], iy = cyi[i], iz = czi[i], jx = cxj[j], jy = cyj[j], jz = czj[j];
                        double acc = 0.;
                        for (int tt = 0; tt <= ix + jx; tt++)
                        for (int u = 0; u <= iy + jy; u++)
                        for (int w = 0; w <= iz + jz; w++)
                            acc += EX(0,ix,jx,tt) * EX(1,iy,jy,u) * EX(2,iz,jz,w) * R[(tt * n1 + u) * n1 + w];
                        size_t idx = (size_t)(ao_loc[ish] + ic * si.nf + i) * nao + ao_loc[jsh] + jc * sj.nf + j;
                        v[idx] -= Z * cc * 2. * M_PI / p * acc;
                    }
                }
            }
#undef EX
        }
        for (int x = 0; x < 3; x++) free(E[x]);
        free(R);
    }
    free(ao_loc);
}
