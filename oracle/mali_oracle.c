/*
 * mali_oracle.c -- CPU restatement of Lightspinner's MALI hot path.  TEST INFRASTRUCTURE ONLY:
 * see mali_oracle.h for the scope, the reference file:line each function follows and the pinning.
 *
 * Build: gcc -O2 -fPIC -shared -fopenmp -ffp-contract=off -fno-fast-math (oracle/Makefile).
 * Every expression is written in the reference's own evaluation order; nothing may be
 * re-associated or contracted.
 */
#include "mali_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* constants.py:1-4,17 -- digit for digit */
static const double CLight = 2.99792458E+08;
static const double HPlanck = 6.6260755E-34;
static const double KBoltzmann = 1.380658E-23;
static const double NM_TO_M = 1.0E-09;
#define HC (HPlanck * CLight)
static const double PI = 3.141592653589793; /* == np.pi */

/* formal_solver.py:14-44 */
void lso_w2(double dtau, double w[2])
{
    if (dtau < 5e-4) {
        w[0] = dtau * (1.0 - 0.5 * dtau);
        /* numba lowers dtau**2 to dtau*dtau and keeps the true divide by 3.0 */
        w[1] = (dtau * dtau) * (0.5 - dtau / 3.0);
    } else if (dtau > 50.0) {
        w[0] = 1.0;
        w[1] = 1.0;
    } else {
        double expdt = exp(-dtau);
        w[0] = 1.0 - expdt;
        w[1] = w[0] - dtau * expdt;
    }
}

/* utils.py:17-22 (numba: the cube is y*y*y) */
double lso_planck(double temp, double wav)
{
    double hc_Tkla = HC / (KBoltzmann * NM_TO_M * wav) / temp;
    double y = NM_TO_M * wav;
    double twohnu3_c2 = (2.0 * HC) / (y * y * y);
    return twohnu3_c2 / (exp(hc_Tkla) - 1.0);
}

/* formal_solver.py:46-142 */
void lso_piecewise_1d_impl(double muz, int toFrom, double Istart, const double *z, const double *chi,
                           const double *S, int Nspace, double *I, double *PsiStar)
{
    double zmu = 1.0 / muz;
    int dk, kStart, kEnd;
    if (toFrom) {
        dk = -1;
        kStart = Nspace - 1;
        kEnd = 0;
    } else {
        dk = 1;
        kStart = 0;
        kEnd = Nspace - 1;
    }

    double dtau_uw = 0.5 * (chi[kStart] + chi[kStart + dk]) * zmu * fabs(z[kStart] - z[kStart + dk]);
    double dS_uw = (S[kStart] - S[kStart + dk]) / dtau_uw;

    double Iupw = Istart;
    double w[2] = {0.0, 0.0};
    I[kStart] = Iupw;
    PsiStar[kStart] = 0.0; /* LambdaStar for now */

    int k = kStart; /* the reference leaves k at its last loop value (kEnd-dk) for the final point */
    for (int kk = kStart + dk; kk != kEnd; kk += dk) {
        k = kk;
        lso_w2(dtau_uw, w);
        I[k] = Iupw * (1.0 - w[0]) + w[0] * S[k] + w[1] * dS_uw;
        PsiStar[k] = w[0] - w[1] / dtau_uw;
        double dtau_dw = 0.5 * (chi[k] + chi[k + dk]) * zmu * fabs(z[k] - z[k + dk]);
        double dS_dw = (S[k] - S[k + dk]) / dtau_dw;
        Iupw = I[k];
        dS_uw = dS_dw;
        dtau_uw = dtau_dw;
    }
    /* formal_solver.py:137-139: w is NOT recomputed and S[k] is S[kEnd-dk] */
    I[kEnd] = (1.0 - w[0]) * Iupw + w[0] * S[k] + w[1] * dS_uw;
    PsiStar[kEnd] = w[0] - w[1] / dtau_uw;

    for (int q = 0; q < Nspace; ++q)
        PsiStar[q] = PsiStar[q] / chi[q];
}

/* formal_solver.py:144-212 */
void lso_piecewise_linear_1d(const double *height, const double *temperature, int Nspace, double muz,
                             int toFrom, double wav, const double *chi, const double *S, double *I,
                             double *PsiStar)
{
    double zmu = 1.0 / muz;
    double Iupw;
    if (toFrom) {
        int kStart = Nspace - 1, dk = -1;
        double dtau_uw = zmu * (chi[kStart] + chi[kStart + dk]) * 0.5 * fabs(height[kStart] - height[kStart + dk]);
        double Bnu0 = lso_planck(temperature[Nspace - 2], wav);
        double Bnu1 = lso_planck(temperature[Nspace - 1], wav);
        Iupw = Bnu1 - (Bnu0 - Bnu1) / dtau_uw;
    } else {
        Iupw = 0.0;
    }
    lso_piecewise_1d_impl(muz, toFrom, Iupw, height, chi, S, Nspace, I, PsiStar);
}

/* ------------------------------------------------------------------------------------------ */

typedef struct {
    int atom, i, j, isLine, Nblue, Nlambda;
} trans_t;

static trans_t get_trans(const lso_model *m, int t)
{
    const int *p = m->trans + (size_t)t * LSO_TRANS_STRIDE;
    trans_t r = {p[0], p[1], p[2], p[3], p[4], p[5]};
    return r;
}

static int trans_active(const trans_t *t, int la)
{
    return la >= t->Nblue && la < t->Nblue + t->Nlambda;
}

/* rh_method.py:245-288 */
void lso_uv(const lso_model *m, const lso_column *c, int t, int la, int mu, int toFrom, const double *gij,
            double *Uji, double *Vij, double *Vji)
{
    trans_t tr = get_trans(m, t);
    int N = m->Nspace;
    int lt = la - tr.Nblue;
    if (tr.isLine) {
        const double *phi = c->phi + c->phioff[t] + (((size_t)lt * m->Nrays + mu) * 2 + (toFrom ? 1 : 0)) * N;
        double hc4piBij = m->lineconst[3 * t + 0];
        double AoverB = m->lineconst[3 * t + 1];
        for (int k = 0; k < N; ++k) {
            Vij[k] = hc4piBij * phi[k];
            Vji[k] = gij[k] * Vij[k];
            Uji[k] = AoverB * Vji[k];
        }
    } else {
        double a = m->alpha[m->toff[t] + lt];
        double c3 = m->twohc_l3[m->toff[t] + lt];
        for (int k = 0; k < N; ++k) {
            Vij[k] = a;
            Vji[k] = gij[k] * a;
            Uji[k] = c3 * Vji[k];
        }
    }
}

typedef struct {
    double *gij, *wla;       /* [Ntrans][N] */
    double *eta;             /* [Natom][N] */
    double *U, *chi;         /* [sumNlevel][N] */
    double *chiTot, *etaTot, *S, *I, *Psi, *Ieff, *Uji, *Vij, *Vji, *JDag;
    int *lvloff, *g2off;
} scratch_t;

static void scratch_alloc(const lso_model *m, scratch_t *s)
{
    int N = m->Nspace;
    int sumL = 0;
    s->lvloff = (int *)malloc(sizeof(int) * (m->Natom + 1));
    s->g2off = (int *)malloc(sizeof(int) * (m->Natom + 1));
    s->lvloff[0] = 0;
    s->g2off[0] = 0;
    for (int a = 0; a < m->Natom; ++a) {
        s->lvloff[a + 1] = s->lvloff[a] + m->Nlevel[a];
        s->g2off[a + 1] = s->g2off[a] + m->Nlevel[a] * m->Nlevel[a];
    }
    sumL = s->lvloff[m->Natom];
    s->gij = (double *)calloc((size_t)m->Ntrans * N, sizeof(double));
    s->wla = (double *)calloc((size_t)m->Ntrans * N, sizeof(double));
    s->eta = (double *)calloc((size_t)m->Natom * N, sizeof(double));
    s->U = (double *)calloc((size_t)sumL * N, sizeof(double));
    s->chi = (double *)calloc((size_t)sumL * N, sizeof(double));
    s->chiTot = (double *)calloc((size_t)N * 10, sizeof(double));
    s->etaTot = s->chiTot + N;
    s->S = s->chiTot + 2 * N;
    s->I = s->chiTot + 3 * N;
    s->Psi = s->chiTot + 4 * N;
    s->Ieff = s->chiTot + 5 * N;
    s->Uji = s->chiTot + 6 * N;
    s->Vij = s->chiTot + 7 * N;
    s->Vji = s->chiTot + 8 * N;
    s->JDag = (double *)calloc((size_t)m->Nspect * N, sizeof(double));
}

static void scratch_free(scratch_t *s)
{
    free(s->lvloff);
    free(s->g2off);
    free(s->gij);
    free(s->wla);
    free(s->eta);
    free(s->U);
    free(s->chi);
    free(s->chiTot);
    free(s->JDag);
}

static double nanmax(double cur, double v)
{
    /* numpy max propagates NaN */
    if (isnan(cur) || isnan(v))
        return NAN;
    return v > cur ? v : cur;
}

static double fs_gamma(const lso_model *m, lso_column *c, scratch_t *s)
{
    const int N = m->Nspace, Nrays = m->Nrays, Nspect = m->Nspect;

    /* rh_method.py:587-590  Gamma.fill(0); Gamma += C */
    size_t g2tot = (size_t)s->g2off[m->Natom] * N;
    for (size_t q = 0; q < g2tot; ++q)
        c->Gamma[q] = 0.0 + c->C[q];

    /* :592-593 */
    memcpy(s->JDag, c->J, sizeof(double) * (size_t)Nspect * N);
    memset(c->J, 0, sizeof(double) * (size_t)Nspect * N);

    for (int la = 0; la < Nspect; ++la) {
        double wav = m->wavelength[la];
        /* setup_wavelength, rh_method.py:425-455 */
        for (int t = 0; t < m->Ntrans; ++t) {
            trans_t tr = get_trans(m, t);
            double *gij = s->gij + (size_t)t * N, *wla = s->wla + (size_t)t * N;
            if (!trans_active(&tr, la)) {
                for (int k = 0; k < N; ++k) {
                    gij[k] = 0.0;
                    wla[k] = 0.0;
                }
                continue;
            }
            int lt = la - tr.Nblue;
            if (tr.isLine) {
                double g = m->lineconst[3 * t + 2];
                double wl = m->wlambda[m->toff[t] + lt];
                const double *wphi = c->wphi + (size_t)t * N;
                for (int k = 0; k < N; ++k) {
                    gij[k] = g;
                    wla[k] = wl * wphi[k] / HC;
                }
            } else {
                const double *gt = c->gijcont + (size_t)(m->toff[t] + lt) * N;
                double w = m->wlacont[m->toff[t] + lt];
                for (int k = 0; k < N; ++k) {
                    gij[k] = gt[k];
                    wla[k] = w;
                }
            }
        }

        for (int mu = 0; mu < Nrays; ++mu) {
            for (int toFrom = 0; toFrom < 2; ++toFrom) {
                for (int k = 0; k < N; ++k) {
                    s->chiTot[k] = 0.0;
                    s->etaTot[k] = 0.0;
                }
                /* zero_angle_dependent_vars */
                memset(s->eta, 0, sizeof(double) * (size_t)m->Natom * N);
                memset(s->U, 0, sizeof(double) * (size_t)s->lvloff[m->Natom] * N);
                memset(s->chi, 0, sizeof(double) * (size_t)s->lvloff[m->Natom] * N);

                /* :604-627 */
                for (int t = 0; t < m->Ntrans; ++t) {
                    trans_t tr = get_trans(m, t);
                    if (!trans_active(&tr, la))
                        continue;
                    const double *gij = s->gij + (size_t)t * N;
                    lso_uv(m, c, t, la, mu, toFrom, gij, s->Uji, s->Vij, s->Vji);
                    int lo = s->lvloff[tr.atom];
                    const double *ni = c->n + (size_t)(lo + tr.i) * N;
                    const double *nj = c->n + (size_t)(lo + tr.j) * N;
                    double *chi_i = s->chi + (size_t)(lo + tr.i) * N;
                    double *chi_j = s->chi + (size_t)(lo + tr.j) * N;
                    double *U_j = s->U + (size_t)(lo + tr.j) * N;
                    double *eta_a = s->eta + (size_t)tr.atom * N;
                    for (int k = 0; k < N; ++k) {
                        double chi = ni[k] * s->Vij[k] - nj[k] * s->Vji[k];
                        double eta = nj[k] * s->Uji[k];
                        chi_i[k] += chi;
                        chi_j[k] -= chi;
                        U_j[k] += s->Uji[k];
                        s->chiTot[k] += chi;
                        s->etaTot[k] += eta;
                        eta_a[k] += eta;
                    }
                }

                /* :630-632 */
                const double *bchi = c->bg_chi + (size_t)la * N;
                const double *beta = c->bg_eta + (size_t)la * N;
                const double *bsca = c->bg_sca + (size_t)la * N;
                const double *JD = s->JDag + (size_t)la * N;
                for (int k = 0; k < N; ++k) {
                    s->chiTot[k] += bchi[k];
                    s->S[k] = (s->etaTot[k] + beta[k] + bsca[k] * JD[k]) / s->chiTot[k];
                }

                /* :635 */
                lso_piecewise_linear_1d(c->height, c->temperature, N, m->muz[mu], toFrom, wav, s->chiTot, s->S,
                                        s->I, s->Psi);
                /* :638-640 */
                c->I[(size_t)la * Nrays + mu] = s->I[0];
                double hw = 0.5 * m->wmu[mu];
                double *Jl = c->J + (size_t)la * N;
                for (int k = 0; k < N; ++k)
                    Jl[k] += hw * s->I[k];

                /* :643-692 */
                for (int a = 0; a < m->Natom; ++a) {
                    const double *eta_a = s->eta + (size_t)a * N;
                    for (int k = 0; k < N; ++k)
                        s->Ieff[k] = s->I[k] - s->Psi[k] * eta_a[k];
                    int lo = s->lvloff[a];
                    int NL = m->Nlevel[a];
                    double *G = c->Gamma + (size_t)s->g2off[a] * N;
                    for (int t = 0; t < m->Ntrans; ++t) {
                        trans_t tr = get_trans(m, t);
                        if (tr.atom != a || !trans_active(&tr, la))
                            continue;
                        const double *gij = s->gij + (size_t)t * N;
                        const double *wla = s->wla + (size_t)t * N;
                        lso_uv(m, c, t, la, mu, toFrom, gij, s->Uji, s->Vij, s->Vji);
                        const double *chi_i = s->chi + (size_t)(lo + tr.i) * N;
                        const double *chi_j = s->chi + (size_t)(lo + tr.j) * N;
                        const double *U_i = s->U + (size_t)(lo + tr.i) * N;
                        const double *U_j = s->U + (size_t)(lo + tr.j) * N;
                        double *Gij = G + ((size_t)tr.i * NL + tr.j) * N;
                        double *Gji = G + ((size_t)tr.j * NL + tr.i) * N;
                        for (int k = 0; k < N; ++k) {
                            double wlamu = wla[k] * hw * 4 * PI;
                            double integrand = (s->Uji[k] + s->Vji[k] * s->Ieff[k]) - (chi_i[k] * s->Psi[k] * U_j[k]);
                            Gij[k] += integrand * wlamu;
                            integrand = (s->Vij[k] * s->Ieff[k]) - (chi_j[k] * s->Psi[k] * U_i[k]);
                            Gji[k] += integrand * wlamu;
                        }
                    }
                }
            }
        }
    }

    /* :698-703 */
    for (int a = 0; a < m->Natom; ++a) {
        int NL = m->Nlevel[a];
        double *G = c->Gamma + (size_t)s->g2off[a] * N;
        for (int k = 0; k < N; ++k) {
            for (int i = 0; i < NL; ++i)
                G[((size_t)i * NL + i) * N + k] = 0.0;
            for (int i = 0; i < NL; ++i) {
                double GamDiag = 0.0;
                for (int l = 0; l < NL; ++l)
                    GamDiag += G[((size_t)l * NL + i) * N + k];
                G[((size_t)i * NL + i) * N + k] = -GamDiag;
            }
        }
    }

    /* :705-708 */
    double dJMax = -INFINITY;
    size_t tot = (size_t)Nspect * N;
    for (size_t q = 0; q < tot; ++q)
        dJMax = nanmax(dJMax, fabs(1.0 - s->JDag[q] / c->J[q]));
    return dJMax;
}

double lso_formal_sol_gamma_matrices(const lso_model *m, lso_column *c)
{
    scratch_t s;
    scratch_alloc(m, &s);
    double r = fs_gamma(m, c, &s);
    scratch_free(&s);
    return r;
}

/* Dense solve, LU with partial pivoting (unblocked, right-looking).  A is n x n row-major, overwritten. */
static int lu_solve(double *A, double *b, int n)
{
    for (int j = 0; j < n; ++j) {
        int p = j;
        double best = fabs(A[j * n + j]);
        for (int i = j + 1; i < n; ++i) {
            double v = fabs(A[i * n + j]);
            if (v > best) {
                best = v;
                p = i;
            }
        }
        if (best == 0.0 || isnan(best))
            return 1;
        if (p != j) {
            for (int q = 0; q < n; ++q) {
                double tmp = A[j * n + q];
                A[j * n + q] = A[p * n + q];
                A[p * n + q] = tmp;
            }
            double tb = b[j];
            b[j] = b[p];
            b[p] = tb;
        }
        double piv = A[j * n + j];
        for (int i = j + 1; i < n; ++i) {
            double l = A[i * n + j] / piv;
            A[i * n + j] = l;
            for (int q = j + 1; q < n; ++q)
                A[i * n + q] -= l * A[j * n + q];
            b[i] -= l * b[j];
        }
    }
    for (int i = n - 1; i >= 0; --i) {
        double acc = b[i];
        for (int q = i + 1; q < n; ++q)
            acc -= A[i * n + q] * b[q];
        b[i] = acc / A[i * n + i];
    }
    return 0;
}

/* rh_method.py:710-745 */
double lso_stat_equil(const lso_model *m, lso_column *c, int *singular)
{
    const int N = m->Nspace;
    double maxRelChange = 0.0;
    int nsing = 0;
    int lvloff = 0, g2off = 0;
    for (int a = 0; a < m->Natom; ++a) {
        int NL = m->Nlevel[a];
        double *A = (double *)malloc(sizeof(double) * NL * NL);
        double *b = (double *)malloc(sizeof(double) * NL);
        double *n = c->n + (size_t)lvloff * N;
        const double *G = c->Gamma + (size_t)g2off * N;
        for (int k = 0; k < N; ++k) {
            int iEl = 0;
            for (int l = 1; l < NL; ++l)
                if (n[(size_t)l * N + k] > n[(size_t)iEl * N + k])
                    iEl = l;
            for (int i = 0; i < NL; ++i)
                for (int j = 0; j < NL; ++j)
                    A[i * NL + j] = (i == iEl) ? 1.0 : G[((size_t)i * NL + j) * N + k];
            for (int i = 0; i < NL; ++i)
                b[i] = 0.0;
            b[iEl] = c->nTotal[(size_t)a * N + k];
            if (lu_solve(A, b, NL)) {
                ++nsing;
                continue;
            }
            for (int l = 0; l < NL; ++l) {
                double nOld = n[(size_t)l * N + k];
                double change = fabs(1.0 - nOld / b[l]);
                maxRelChange = nanmax(maxRelChange, change);
                n[(size_t)l * N + k] = b[l];
            }
        }
        free(A);
        free(b);
        lvloff += NL;
        g2off += NL * NL;
    }
    if (singular)
        *singular = nsing;
    return maxRelChange;
}

/* test.py:20-29 / response_fn.py:11-21 for a batch of independent columns, fixed iteration count */
void lso_iterate_batch(const lso_model *m, lso_column *cols, int ncol, int niter, int start_iter,
                       double *dJ_out, double *dPops_out)
{
#pragma omp parallel
    {
        scratch_t s;
        scratch_alloc(m, &s);
#pragma omp for schedule(dynamic, 1)
        for (int ic = 0; ic < ncol; ++ic) {
            double dJ = 1.0, dPops = 1.0;
            for (int it = start_iter + 1; it <= start_iter + niter; ++it) {
                dJ = fs_gamma(m, &cols[ic], &s);
                if (it > 3) {
                    int sing = 0;
                    dPops = lso_stat_equil(m, &cols[ic], &sing);
                }
            }
            if (dJ_out)
                dJ_out[ic] = dJ;
            if (dPops_out)
                dPops_out[ic] = dPops;
        }
        scratch_free(&s);
    }
}

void lso_set_threads(int n)
{
#ifdef _OPENMP
    if (n > 0)
        omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int lso_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
