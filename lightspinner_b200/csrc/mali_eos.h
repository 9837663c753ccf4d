// mali_eos.h -- the column set-up in front of the MALI hot path, host/device source:
//   * the Wittmann equation of state as coded in the reference (witt.py:198-742): electron and gas pressure from
//     temperature and density (pe_from_rho, pg_from_rho and the iterations under them), partial densities of the
//     background absorbers (getBackgroundPartials);
//   * the background continuum opacity `cop` with its 20-odd sources (witt.py:778-1365);
// restated function by function with the reference's evaluation order (numba lowers x**2, x**3 to products and calls
// libm for exp / log / pow; the device's versions of those differ by <= 2 ulp, which is what the 1e-12 bar of
// tests/test_eos_host.py and tests/test_gpu_background.py allows for).  Compiled for the host by tests/eos_shim.cpp
// (unit test against the reference's own numbers) and for the device by mali_api.cu (background_kernel).
#pragma once
#include <cmath>
#include <cstdint>

#ifdef __CUDACC__
#define MALI_EOS_HD __host__ __device__ inline
#else
#define MALI_EOS_HD inline
#endif

namespace mali {
namespace eos {

// witt.py:41-49 (NIST values as typed there)
constexpr double BK = 1.3806488E-16, HH = 6.62606957E-27, EV = 1.602176565E-12;
constexpr int kNcontr = 28;      // witt.py:156
constexpr int kMaxStage = 16;

struct Tables {          // state of a witt() instance (model-level; host or device pointers)
    int32_t npf;
    const double *tpf;           // [npf] temperatures of the partition-function tables
    const double *pf;            // [sum nstage][npf]
    const double *eion;          // [sum nstage] ionisation energies in eV
    const int32_t *stageOff;     // [kNcontr + 1]
    const double *abund;         // [99] abundances as normalised by witt.__init__
    double avw, rho_from_H, ab_others, saha_fac, prec;
};

MALI_EOS_HD double acota(double x, double x0, double x1) { return x < x0 ? x0 : (x > x1 ? x1 : x); }
MALI_EOS_HD double acotasig(double x, double x0, double x1) { return x < 0 ? -acota(-x, x0, x1) : acota(x, x0, x1); }
MALI_EOS_HD double sign(double a, double b) { return fabs(a) * (b / fabs(b)); }

// witt.py:479-500
MALI_EOS_HD double itep1(const double *x, const double *y, int n, double xx)
{
    if (xx <= x[0]) return y[0];
    if (xx >= x[n - 1]) return y[n - 1];
    int p0 = 0;
    while (!(x[p0] > xx)) ++p0;
    const int p1 = p0 - 1;
    const double dx = x[p1] - x[p0];
    const double u1 = (xx - x[p0]) / dx;
    const double u0 = 1.0 - u1;
    return u0 * y[p0] + u1 * y[p1];
}

// witt.py:504-528: u[0..nn-1]; entries beyond the element's stages stay 0 (only = 3 on hydrogen)
MALI_EOS_HD int partition_f(const Tables &E, int n, double t, int only, double *res)
{
    int nn = E.stageOff[n + 1] - E.stageOff[n];
    const int len = only > 0 ? only : nn;
    if (only > 0 && nn > only) nn = only;
    for (int i = 0; i < len; ++i) res[i] = 0.0;
    for (int i = 0; i < nn; ++i) res[i] = itep1(E.tpf, E.pf + (size_t)(E.stageOff[n] + i) * E.npf, E.npf, t);
    return len;
}

// witt.py:198-206
MALI_EOS_HD double nsaha(const Tables &E, double t, double xne, double u0, double u1, double eion)
{
    return 2.0 * E.saha_fac * (u1 / u0) * pow(t, 1.5) * exp(-eion * EV / (t * BK)) / xne;
}
MALI_EOS_HD double saha(double theta, double eion, double u1, double u2, double pe)
{
    return u2 * exp(2.302585093 * (9.0804625434325867 - theta * eion)) / (u1 * pe * pow(theta, 2.5));
}

// witt.py:329-338 (the derivatives are never used)
MALI_EOS_HD void molecb(double X, double &Y0, double &Y1)
{
    Y0 = -11.206998 + X * (2.7942767 + X * (7.9196803E-2 - X * 2.4790744E-2));
    Y1 = -12.533505 + X * (4.9251644 + X * (-5.6191273E-2 + X * 3.2687661E-3));
}

// Everything pe_pg / gasc evaluate that depends on the temperature only -- the partition functions of the 28 electron
// donors, the exponentials and the power of theta inside saha(), the molecular equilibrium constants -- is formed once
// per (temperature) point: the EOS iterations call pe_pg / gasc dozens of times at one temperature, and the values
// are the very same each time (same expressions, same bits; only the products with pe are redone).
struct PointCache {
    double theta, th25;              // 5040 / t, theta**2.5
    double u[kNcontr][3];            // partition_f(ii, t, only=3)
    double ex1[kNcontr], ex2[kNcontr];   // exp(2.302585093 * (9.0804625434325867 - theta * eion[0 | 1]));  H: eion[0], 0.754
    double p10c[2], p10[2];          // 10**molecb (clamped to +-30 as pe_pg does / as gasc takes it)
};

MALI_EOS_HD void point_cache(const Tables &E, double t, PointCache &C)
{
    C.theta = 5040.0 / t;
    C.th25 = pow(C.theta, 2.5);
    for (int ii = 0; ii < kNcontr; ++ii) {
        partition_f(E, ii, t, 3, C.u[ii]);
        const double e0 = E.eion[E.stageOff[ii]];
        const double e1 = ii == 0 ? 0.754 : E.eion[E.stageOff[ii] + 1];
        C.ex1[ii] = exp(2.302585093 * (9.0804625434325867 - C.theta * e0));
        C.ex2[ii] = exp(2.302585093 * (9.0804625434325867 - C.theta * e1));
    }
    double c0, c1;
    molecb(C.theta, c0, c1);
    C.p10[0] = pow(10.0, c0);
    C.p10[1] = pow(10.0, c1);
    C.p10c[0] = pow(10.0, acota(c0, -30., 30.));
    C.p10c[1] = pow(10.0, acota(c1, -30., 30.));
}
// saha() with its temperature-only factors taken from the cache: u2 * ex / (u1 * pe * theta**2.5)
MALI_EOS_HD double saha_c(double ex, double u1, double u2, double pe, double th25) { return u2 * ex / (u1 * pe * th25); }

// witt.py:342-432
MALI_EOS_HD double pe_pg(const Tables &E, const PointCache &C, double pe, double pgas, double &fe_out)
{
    double g1 = 0.0;
    double g4, g5;
    if (pe < 0.0) {
        pe = 1.e-15;
        g4 = 0.0;
        g5 = 0.0;
    } else {
        g4 = pe * C.p10c[0];
        g5 = pe * C.p10c[1];
    }
    const double g2 = saha_c(C.ex1[0], C.u[0][0], C.u[0][1], pe, C.th25);
    double g3 = saha_c(C.ex2[0], 1.0, C.u[0][0], pe, C.th25);
    g3 = 1.0 / acota(g3, 1.e-30, 1.0e30);
    for (int ii = 1; ii < kNcontr; ++ii) {
        const double alfai = E.abund[ii] / E.abund[0];
        const double a = saha_c(C.ex1[ii], C.u[ii][0], C.u[ii][1], pe, C.th25);
        const double b = saha_c(C.ex2[ii], C.u[ii][1], C.u[ii][2], pe, C.th25);
        const double c = 1. + a * (1. + b);
        g1 += alfai / c * a * (1. + 2. * b);
    }
    double a = 1. + g2 + g3;
    const double b = 2. * (1. + g2 / g5 * g4);
    const double c = g5;
    double d = g2 - g3;
    const double e = g2 / g5 * g4;
    a = acotasig(a, 1.e-15, 1.e15);
    d = acotasig(d, 1.e-15, 1.e15);
    const double c1 = c * b * b + a * d * b - e * a * a;
    const double c2 = 2.0 * a * e - d * b + a * b * g1;
    const double c3 = -(e + b * g1);
    double f1 = 0.5 * c2 / c1;
    f1 = -f1 + sign(1., c1) * sqrt(f1 * f1 - c3 / c1);
    double f5 = (1. - a * f1) / b;
    double f4 = e * f5;
    const double f3 = g3 * f1;
    const double f2 = g2 * f1;
    double fe = acota(f2 - f3 + f4 + g1, 1.e-30, 1.e30);
    double phtot = pe / fe;
    if (f5 <= 1.e-4) {
        double diff = 1.0;
        const double const6 = g5 / pe * f1 * f1, const7 = f2 - f3 + g1;
        int it = 0;
        while ((diff > 1.e-5) && (it < 5)) {
            const double of5 = f5;
            f5 = phtot * const6;
            f4 = e * f5;
            fe = const7 + f4;
            phtot = pe / fe;
            diff = 0.5 * fabs(f5 - of5) / (f5 + of5);
            it += 1;
        }
    }
    pe = pgas / (1. + (f1 + f2 + f3 + f4 + f5 + E.ab_others) / fe);
    if (pe <= 0.0) pe = 1.e-15;
    fe_out = fe;
    return pe;
}

// witt.py:541-621.  pp: the hydrogen entries only (f1, f2, f5, f3, phtot, fe) -- the per-species entries are never read
MALI_EOS_HD double gasc(const Tables &E, const PointCache &C, double pe, double *pp6)
{
    const double g4 = C.p10[0], g5 = C.p10[1];
    const double g2 = saha_c(C.ex1[0], C.u[0][0], C.u[0][1], pe, C.th25);
    const double g3 = 1.0 / saha_c(C.ex2[0], 1.0, C.u[0][0], pe, C.th25);
    double g1 = 0.0;
    for (int ii = 1; ii < kNcontr; ++ii) {
        const double alfai = E.abund[ii] / E.abund[0];
        const double a = saha_c(C.ex1[ii], C.u[ii][0], C.u[ii][1], pe, C.th25);
        const double b = saha_c(C.ex2[ii], C.u[ii][1], C.u[ii][2], pe, C.th25);
        const double c = 1. + a * (1. + b);
        const double ppi = alfai / c;
        g1 += ppi * a * (1. + 2. * b);
    }
    const double a = 1. + g2 + g3;
    const double e = g2 / g5 * g4;
    const double b = 2.0 * (1.0 + e);
    const double c = g5;
    const double d = g2 - g3;
    const double c1 = c * b * b + a * d * b - e * a * a;
    const double c2 = 2. * a * e - d * b + a * b * g1;
    const double c3 = -(e + b * g1);
    double f1 = 0.5 * c2 / c1;
    f1 = -f1 + sign(1.0, c1) * sqrt(f1 * f1 - c3 / c1);
    double f5 = (1.0 - a * f1) / b;
    double f4 = e * f5;
    const double f3 = g3 * f1;
    const double f2 = g2 * f1;
    double fe = f2 - f3 + f4 + g1;
    double phtot = pe / fe;
    if (f5 <= 1.e-5) {
        double diff = 1.0;
        const double const6 = g5 / pe * f1 * f1, const7 = f2 - f3 + g1;
        int it = 0;
        while ((diff > 1.e-5) && (it < 5)) {
            const double of5 = f5;
            f5 = phtot * const6;
            f4 = e * f5;
            fe = const7 + f4;
            phtot = pe / fe;
            diff = 0.5 * fabs(f5 - of5) / (f5 + of5);
            it += 1;
        }
    }
    const double pg = pe * (1.0 + (f1 + f2 + f3 + f4 + f5 + E.ab_others) / fe);
    if (pp6) {
        pp6[0] = f1;
        pp6[1] = f2;
        pp6[2] = f5;
        pp6[3] = f3;
        pp6[4] = phtot;
        pp6[5] = fe;
    }
    return pg;
}

// witt.py:210-222
MALI_EOS_HD double init_pe_from_pg(const Tables &E, double t, double pg)
{
    const double nu = E.abund[0];
    const double sa = pow(10.0, -0.4771 + 2.5 * log10(t) - log10(pg) - (13.6 * 5040.0 / t));
    const double aaa = 1.0 + sa;
    const double bbb = -(nu - 1.0) * sa;
    const double ccc = -sa * nu;
    const double ybh = (-bbb + sqrt(bbb * bbb - 4. * aaa * ccc)) / (2. * aaa);
    return pg * ybh / (1. + ybh);
}

// witt.py:226-244
MALI_EOS_HD double pe_from_pg(const Tables &E, const PointCache &C, double t, double pg, double *fe_out = nullptr)
{
    double dif = 1.1;
    double pe = init_pe_from_pg(E, t, pg);
    double ope = pe, fe = 0.0;
    int it = 0;
    while ((fabs(dif) > E.prec) && (it < 250)) {
        pe = (ope + pe) * 0.5;
        ope = pe;
        pe = pe_pg(E, C, pe, pg, fe);
        dif = 2.0 * fabs(pe - ope) / (pe + ope);
        it += 1;
    }
    if (fe_out) *fe_out = fe;
    return pe;
}

MALI_EOS_HD double start_fraction(double t) { return t > 8000 ? 0.5 : (t > 4000 ? 0.1 : (t > 2000 ? 0.01 : 0.001)); }

// witt.py:248-279 (the reference's loop counter never advances: the loop ends on convergence only; a hard cap keeps a
// pathological input from hanging a GPU)
MALI_EOS_HD double pe_from_rho(const Tables &E, const PointCache &C, double t, double rho)
{
    const double xna = rho / E.avw;
    const double BKT = BK * t;
    const double a = start_fraction(t);
    const double xne = a * xna / (1.0 - a);
    double Pgas = (xna + xne) * BKT;
    double dif = 1.0, Pe = 0.0;
    int guard = 0;
    while (fabs(dif) > E.prec && guard < 1000) {
        Pe = pe_from_pg(E, C, t, Pgas);
        const double xna_guessed = (Pgas - Pe) / BKT;
        dif = fabs(xna - xna_guessed) / xna;
        Pgas *= xna / xna_guessed;
        ++guard;
    }
    return Pe;
}

// witt.py:312-318
MALI_EOS_HD double rho_from_pe(const Tables &E, const PointCache &C, double temp, double pe)
{
    double pp[6];
    gasc(E, C, pe, pp);
    return pe * E.rho_from_H / (pp[5] * temp);
}

// witt.py:283-308
MALI_EOS_HD double pg_from_rho(const Tables &E, const PointCache &C, double temp, double rho)
{
    const double xna = rho / E.avw;
    const double a = start_fraction(temp);
    const double xne = a * xna / (1.0 - a);
    const double pgas0 = (xna + xne) * BK * temp;
    double Pe = pe_from_pg(E, C, temp, pgas0);
    double irho = rho_from_pe(E, C, temp, Pe);
    double dif = 1.0;
    int it = 0;
    while ((dif >= E.prec) && (it < 100)) {
        Pe *= (1.0 + rho / irho) * 0.5;
        irho = rho_from_pe(E, C, temp, Pe);
        dif = fabs((irho - rho) / (rho));
        it += 1;
    }
    return gasc(E, C, Pe, nullptr);
}

// witt.py:625-667 with divide_by_u (only = 0: every stage of the element); xpa[0..nLev-1]
MALI_EOS_HD int getXparts(const Tables &E, int iatom, double t, double pg, double pe, bool divide_by_u, double *xpa)
{
    const double TBK = t * BK, xna = (pg - pe) / TBK, xne = pe / TBK;
    const double n_tot = xna * E.abund[iatom] / 1.0;      // abtot == 1.0 after witt.__init__
    double u[kMaxStage];
    const int nLev = partition_f(E, iatom, t, 0, u);
    xpa[0] = 1.0;
    for (int ii = 1; ii < nLev; ++ii) xpa[ii] = nsaha(E, t, xne, u[ii - 1], u[ii], E.eion[E.stageOff[iatom] + ii - 1]);
    for (int ii = nLev - 1; ii > 0; --ii) xpa[0] = 1.0 + xpa[0] * xpa[ii];
    xpa[0] = 1.0 / xpa[0];
    for (int ii = 1; ii < nLev; ++ii) xpa[ii] *= xpa[ii - 1];
    for (int ii = 0; ii < nLev; ++ii) xpa[ii] *= divide_by_u ? n_tot / u[ii] : n_tot;
    return nLev;
}

// witt.py:671-740 with divide_by_u = True: n[17]
MALI_EOS_HD void background_partials(const Tables &E, const PointCache &C, double t, double pg, double pe, double *n)
{
    const double tbk = t * BK;
    double x[kMaxStage];
    getXparts(E, 1, t, pg, pe, true, x);   // He / He+ / He++
    n[3] = x[0];
    n[4] = x[1];
    n[5] = x[2];
    getXparts(E, 5, t, pg, pe, true, x);   // C
    n[6] = x[0];
    getXparts(E, 12, t, pg, pe, true, x);  // Al
    n[7] = x[0];
    getXparts(E, 13, t, pg, pe, true, x);  // Si / Si+
    n[8] = x[0];
    n[9] = x[1];
    getXparts(E, 19, t, pg, pe, true, x);  // Ca / Ca+
    n[10] = x[0];
    n[11] = x[1];
    getXparts(E, 11, t, pg, pe, true, x);  // Mg / Mg+
    n[12] = x[0];
    n[13] = x[1];
    getXparts(E, 25, t, pg, pe, true, x);  // Fe
    n[14] = x[0];
    getXparts(E, 6, t, pg, pe, true, x);   // N
    n[15] = x[0];
    getXparts(E, 7, t, pg, pe, true, x);   // O
    n[16] = x[0];
    double pp[6];
    gasc(E, C, pe, pp);
    n[0] = pp[0] * pp[4] / tbk * 0.5;      // H / pf[H]
    n[1] = pp[1] * pp[4] / tbk;            // H+
    n[2] = pp[3] * pp[4] / tbk;            // H-
}

// ---------------------------------------------------------------------------------------------------------
// background opacity sources, witt.py:778-1365
#include "mali_eos_tables.inc"

MALI_EOS_HD double cube(double x) { return x * x * x; }      // numba lowers x**3 to products

MALI_EOS_HD double SEATON(double FREQ0, double XSECT, double POWER, double A, double FREQ)
{
    return XSECT * (A + (1. - A) * (FREQ0 / FREQ)) * pow(FREQ0 / FREQ, floor(2. * POWER + 0.01) * 0.5);
}

MALI_EOS_HD double COULFF(double TLOG, double FREQLG, int NZ)
{
    const double GAMLOG = 10.39638 - TLOG / 1.15129 + Z4LOG[NZ - 1];
    int IGAM = (int)(GAMLOG + 7.);
    if (IGAM > 10) IGAM = 10;
    if (IGAM < 1) IGAM = 1;
    const double HVKTLG = (FREQLG - TLOG) / 1.15129 - 20.63764;
    int IHVKT = (int)(HVKTLG + 9.);
    if (IHVKT > 11) IHVKT = 11;
    if (IHVKT < 1) IHVKT = 1;
    const double P = GAMLOG - (IGAM - 7);
    const double Q = HVKTLG - (IHVKT - 9);
    return (1. - P) * ((1. - Q) * A0[(IHVKT - 1) * 11 + IGAM - 1] + Q * A0[IHVKT * 11 + IGAM - 1]) +
           P * ((1. - Q) * A0[(IHVKT - 1) * 11 + IGAM] + Q * A0[IHVKT * 11 + IGAM]);
}

MALI_EOS_HD double COULX(int N, double freq, double Z)
{
    const double n = (N + 1.0) * (N + 1.0);
    if (freq >= (Z * Z * 3.28805e15 / n)) {
        const double FREQ1 = freq * 1.e-10;
        double CLX = 0.2815 / FREQ1 / FREQ1 / FREQ1 / n / n / (N + 1.0) * Z * Z * Z * Z;
        if (N >= 6) return CLX;
        CLX *= (A1c[N] + (B1c[N] + C1c[N] * (Z * Z / FREQ1)) * (Z * Z / FREQ1));
        return CLX;
    }
    return 0.0;
}

MALI_EOS_HD double HOP(double XNE, double XH1, double XH2, double FREQ, double FREQLG, double T, double TLOG, double TKEV,
                       double STIM, double EHVKT)
{
    double CONT[8], BOLT[8];
    const double FREQ3 = cube(FREQ * 1.E-10);
    const double CFREE = 3.6919E-22 / FREQ3;
    for (int N = 0; N < 8; ++N) {
        const double n1 = (N + 1.0) * (N + 1.0);
        BOLT[N] = exp(-13.595 * (1. - 1. / n1) / TKEV) * 2. * n1 * XH1;
    }
    const double FREET = XNE * CFREE * XH2 / sqrt(T);
    const double XR = XH1 / 13.595 * TKEV;
    double BOLTEX = exp(-13.427 / TKEV) * XR;
    const double EXLIM = exp(-13.595 / TKEV) * XR;
    for (int N = 0; N < 8; ++N) CONT[N] = COULX(N, FREQ, 1.0);
    const double C = 0.2815 / FREQ3;
    if (FREQ < 4.05933E13) BOLTEX = EXLIM / EHVKT;
    double H = (CONT[6] * BOLT[6] + CONT[7] * BOLT[7] + (BOLTEX - EXLIM) * C + COULFF(TLOG, FREQLG, 1) * FREET) * STIM;
    double s = 0.0;
    for (int N = 0; N < 6; ++N) s += CONT[N] * BOLT[N];
    H += s * (1. - EHVKT);
    return H;
}

MALI_EOS_HD double HRAYOP(double XH1, double FREQ)
{
    double WAVE = FREQ < 2.463e15 ? FREQ : 2.463e15;
    WAVE = 2.997925e18 / WAVE;
    const double WW = WAVE * WAVE;
    const double WW2 = WW * WW;
    const double SIG = (5.799e-13 + 1.422e-6 / WW + 2.784 / (WW2)) / (WW2);
    return SIG * XH1 * 2.0;
}

MALI_EOS_HD double H2PLOP(double XH1, double XH2, double FREQ, double FREQLG, double FREQ15, double TKEV, double STIM)
{
    if (FREQ > 3.28805E15) return 0.0;
    const double FR = -3.0233E3 + (3.7797E2 + (-1.82496E1 + (3.9207E-1 - 3.1672E-3 * FREQLG) * FREQLG) * FREQLG) * FREQLG;
    const double ES =
        -7.342E-3 + (-2.409 + (1.028 + (-0.4230 + (0.1224 - 0.01351 * FREQ15) * FREQ15) * FREQ15) * FREQ15) * FREQ15;
    return exp(-ES / TKEV + FR) * 2. * XH1 * XH2 * STIM;
}

MALI_EOS_HD double HMINOP(double XH1, double XHMIN, double FREQ, double T, double TKEV, double XNE, double EHVKT)
{
    const double FREQ1 = FREQ * 1.E-10;
    const double B = (1.3727E-15 + 4.3748 / FREQ) / FREQ1;
    const double C = -2.5993E-7 / (FREQ1 * FREQ1);
    double HMINBF;
    if (FREQ <= 1.8259E14)
        HMINBF = 0.;
    else if (FREQ >= 2.111E14)
        HMINBF = 6.801E-10 + (5.358E-3 + (1.481E3 + (-5.519E7 + 4.808E11 / FREQ1) / FREQ1) / FREQ1) / FREQ1;
    else
        HMINBF = 3.695E-6 + (-1.251E-1 + 1.052E3 / FREQ1) / FREQ1;
    const double HMINFF = (B + C / T) * XH1 * XNE * 2.E-20;
    double HMIN;
    if (T < 7730.)
        HMIN = XHMIN;
    else
        HMIN = exp(0.7552 / TKEV) / (2. * 2.4148E15 * T * sqrt(T)) * XH1 * XNE;
    const double H = HMINBF * (1 - EHVKT) * HMIN * 1.E-10;
    return H + HMINFF;
}

MALI_EOS_HD double HE1OP(double XHE1, double XHE2, double XNE, double FREQ, double FREQLG, double T, double TKEV,
                         double TLOG, double EHVKT, double STIM)
{
    double TRANS[10], BOLT[10];
    for (int q = 0; q < 10; ++q) {
        TRANS[q] = 0.0;
        BOLT[q] = exp(-CHI0[q] / TKEV) * G0[q] * XHE1;
    }
    const double FREET = XNE * 1.E-10 * XHE2 * 1.E-10 / sqrt(T) * 1.E-10;
    const double XRLOG = log(XHE1 * (2. / 13.595) * TKEV);
    const double BOLTEX = exp(-23.730 / TKEV + XRLOG);
    const double EXLIM = exp(-24.587 / TKEV + XRLOG);
    const double FREQ3 = cube(FREQ * 1.E-10);
    const double CFREE = 3.6919E8 / FREQ3;
    const double C = 2.815E-1 / FREQ3;
    int NMIN = 9;      // the reference's loop variable keeps its last value when no threshold is met
    for (int q = 0; q < 10; ++q)
        if (HEFREQ0[q] <= FREQ) {
            NMIN = q;
            break;
        }
    const double dum[10] = {33.32 - 2. * FREQLG,  -390.026 + (21.035 - 0.318 * FREQLG) * FREQLG,
                            26.83 - 1.91 * FREQLG, 61.21 - 2.9 * FREQLG,
                            81.35 - 3.5 * FREQLG,  12.69 - 1.54 * FREQLG,
                            23.85 - 1.86 * FREQLG, 49.30 - 2.60 * FREQLG,
                            85.20 - 3.69 * FREQLG, 58.81 - 2.89 * FREQLG};
    for (int q = NMIN; q < 10; ++q) TRANS[q] = exp(dum[q]);
    double EX = BOLTEX;
    if (FREQ < 2.055E14) EX = EXLIM / EHVKT;
    double HE1 = (EX - EXLIM) * C;
    double s = 0.0;
    for (int q = 0; q < 10; ++q) s += TRANS[q] * BOLT[q];
    HE1 += s;
    return (HE1 + COULFF(TLOG, FREQLG, 1) * FREET * CFREE) * STIM;
}

MALI_EOS_HD double HE2OP(double XHE2, double XHE3, double XNE, double FREQ, double FREQLG, double T, double TKEV,
                         double TLOG, double EHVKT, double STIM)
{
    double CONT[9], BOLT[9];
    for (int N = 0; N < 9; ++N) {
        const double N12 = (N + 1.0) * (N + 1.0);
        BOLT[N] = exp(-(54.403 - 54.403 / N12) / TKEV) * 2. * N12 * XHE2;
    }
    const double FREET = XNE * XHE3 / sqrt(T);
    const double XR = XHE2 / 13.595 * TKEV;
    const double BOLTEX = exp(-53.859 / TKEV) * XR;
    const double EXLIM = exp(-54.403 / TKEV) * XR;
    for (int N = 0; N < 9; ++N) CONT[N] = COULX(N, FREQ, 2.0);
    const double FREQ3 = cube(FREQ * 1.E-5);
    const double CFREE = 3.6919E-07 / FREQ3 * 4.0;
    const double C = 2.815E14 * 2.0 * 2.0 / FREQ3;
    double EX = BOLTEX;
    if (FREQ < 1.31522E14) EX = EXLIM / EHVKT;
    double HE2 = (EX - EXLIM) * C;
    double s = 0.0;
    for (int N = 0; N < 9; ++N) s += CONT[N] * BOLT[N];
    HE2 += s;
    HE2 = (HE2 + COULFF(TLOG, FREQLG, 2) * CFREE * FREET) * STIM;
    return HE2 >= 1.E-20 ? HE2 : 0.0;
}

MALI_EOS_HD double HEMIOP(double XHE1, double FREQ, double T, double XNE)
{
    const double A = 3.397E-26 + (-5.216E-11 + 7.039E05 / FREQ) / FREQ;
    const double B = -4.116E-22 + (1.067E-06 + 8.135E09 / FREQ) / FREQ;
    const double C = 5.081E-17 + (-8.724E-03 - 5.659E12 / FREQ) / FREQ;
    return (A * T + B + C / T) * XNE * XHE1 * 1.E-20;
}

MALI_EOS_HD double HERAOP(double XHE1, double FREQ)
{
    const double f = FREQ * 1.E-15 < 5.15 ? FREQ * 1.E-15 : 5.15;
    const double q = 2.997925E+03 / f;
    const double WW = q * q;
    const double arg = 1. + (2.44E5 + 5.94E10 / (WW - 2.90E5)) / WW;
    const double SIG = 5.484E-14 / WW / WW * arg * arg;
    return SIG * XHE1;
}

MALI_EOS_HD double Mg1OP(double FREQ, double FREQLG, double T, double TLOG)
{
    int NT = (int)floor(T / 1000.) - 3;
    if (NT > 6) NT = 6;
    if (NT < 1) NT = 1;
    const double DT = (TLOG - TLG0[NT - 1]) / (TLG0[NT] - TLG0[NT - 1]);
    int N = 6;
    for (int q = 0; q < 7; ++q)
        if (FREQ > FREQMG[q]) {
            N = q;
            break;
        }
    const double D = (FREQLG - FLOG0[N]) / (FLOG0[N + 1] - FLOG0[N]);
    if (N > 1) N = 2 * N - 1;
    const double D1 = 1.0 - D;
    const double XWL1 = PEACH0[(N + 1) * 7 + NT - 1] * D + PEACH0[N * 7 + NT - 1] * D1;
    const double XWL2 = PEACH0[(N + 1) * 7 + NT] * D + PEACH0[N * 7 + NT] * D1;
    return exp(XWL1 * (1. - DT) + XWL2 * DT);
}

MALI_EOS_HD double C1OP(double FREQ, double TKEV)
{
    const double C1240 = 5. * exp(-1.264 / TKEV);
    const double C1444 = exp(-2.683 / TKEV);
    double X1444 = 0.0, X1240 = 0.0, X1100 = 0.0;
    if (FREQ >= 2.7254E15) X1100 = SEATON(2.7254E15, 1.219E-17, 2.0E0, 3.317E0, FREQ);
    if (FREQ >= 2.4196E15) X1240 = SEATON(2.4196E15, 1.030E-17, 1.5E0, 2.789E0, FREQ);
    if (FREQ >= 2.0761E15) X1444 = SEATON(2.0761E15, 9.590E-18, 1.5E0, 3.501E0, FREQ);
    return X1100 * 9. + X1240 * C1240 + X1444 * C1444;
}

MALI_EOS_HD double Al1OP(double FREQ)
{
    if (FREQ > 1.443E15) return 2.1E-17 * (cube(1.443E15 / FREQ)) * 6.0;
    return 0.0;
}

MALI_EOS_HD double Si1OP(double FREQ, double FREQLG, double T, double TLOG)
{
    int NT = (int)floor(T / 1000.) - 3;
    if (NT > 8) NT = 8;
    if (NT < 1) NT = 1;
    const double DT = (TLOG - TLG1[NT - 1]) / (TLG1[NT] - TLG1[NT - 1]);
    int N = 8;
    for (int q = 0; q < 9; ++q)
        if (FREQ > FREQSI1[q]) {
            N = q;
            break;
        }
    const double D = (FREQLG - FLOG1[N]) / (FLOG1[N + 1] - FLOG1[N]);
    if (N > 1) N = 2 * N - 1;
    const double DD = 1. - D;
    const double XWL1 = PEACH1[(N + 1) * 9 + NT - 1] * D + PEACH1[N * 9 + NT - 1] * DD;
    const double XWL2 = PEACH1[(N + 1) * 9 + NT] * D + PEACH1[N * 9 + NT] * DD;
    return exp(-(XWL1 * (1. - DT) + XWL2 * DT)) * 9.;
}

MALI_EOS_HD double Fe1OP(double FREQ, double HKT)
{
    const double WAVENO = FREQ / 2.99792458E10;
    if (WAVENO < 21000.) return 0.0;
    double s = 0.0;
    for (int q = 0; q < 48; ++q) {
        const double BOLT = G1[q] * exp(-E1[q] * 2.99792458e10 * HKT);
        double XSECT = 0.0;
        if (WNO1[q] < WAVENO) {
            const double XXX = ((WNO1[q] + 3000. - WAVENO) / WNO1[q] / .1);
            const double x2 = XXX * XXX;
            XSECT = 3.e-18 / (1. + x2 * x2);
        }
        s += XSECT * BOLT;
    }
    return s;
}

MALI_EOS_HD double COOLOP(double XC1, double XMg1, double XAl1, double XSi1, double XFe1, double STIM, double FREQ,
                          double FREQLG, double T, double TLOG, double TKEV, double HKT)
{
    return (C1OP(FREQ, TKEV) * XC1 + Mg1OP(FREQ, FREQLG, T, TLOG) * XMg1 + Al1OP(FREQ) * XAl1 +
            Si1OP(FREQ, FREQLG, T, TLOG) * XSi1 + Fe1OP(FREQ, HKT) * XFe1) *
           STIM;
}

MALI_EOS_HD double N1OP(double FREQ, double TKEV)
{
    const double C1130 = 6. * exp(-3.575 / TKEV);
    const double C1020 = 10. * exp(-2.384 / TKEV);
    double X1130 = 0., X1020 = 0., X853 = 0.;
    if (FREQ >= 3.517915E15) X853 = SEATON(3.517915E15, 1.142E-17, 2.0E0, 4.29E0, FREQ);
    if (FREQ >= 2.941534E15) X1020 = SEATON(2.941534E15, 4.410E-18, 1.5E0, 3.85E0, FREQ);
    if (FREQ >= 2.653317E15) X1130 = SEATON(2.653317E15, 4.200E-18, 1.5E0, 4.34E0, FREQ);
    return X853 * 4. + X1020 * C1020 + X1130 * C1130;
}

MALI_EOS_HD double O1OP(double FREQ)
{
    if (FREQ >= 3.28805E15) return 9. * SEATON(3.28805E15, 2.94E-18, 1.E0, 2.66E0, FREQ);
    return 0.0;
}

MALI_EOS_HD double Mg2OP(double FREQ, double TKEV)
{
    const double C1169 = 6. * exp(-4.43 / TKEV);
    double X1169 = 0.0, X824 = 0.0;
    if (FREQ >= 3.635492E15) X824 = SEATON(3.635492E15, 1.40E-19, 4.E0, 6.7E0, FREQ);
    if (FREQ >= 2.564306E15) X1169 = 5.11E-19 * cube(2.564306E15 / FREQ);
    return X824 * 2. + X1169 * C1169;
}

MALI_EOS_HD double Si2OP(double FREQ, double FREQLG, double T, double TLOG)
{
    int NT = (int)floor(T / 2000.) - 4;
    if (NT > 5) NT = 5;
    if (NT < 1) NT = 1;
    const double DT = (TLOG - TLG2[NT - 1]) / (TLG2[NT] - TLG2[NT - 1]);
    int N = 6;
    for (int q = 0; q < 7; ++q)
        if (FREQ > FREQSI2[q]) {
            N = q;
            break;
        }
    const double D = (FREQLG - FLOG2[N]) / (FLOG2[N + 1] - FLOG2[N]);
    if (N > 1) N = 2 * N - 2;
    if (N == 13) N = 12;
    const double D1 = 1. - D;
    const double XWL1 = PEACH2[(N + 1) * 6 + NT - 1] * D + PEACH2[N * 6 + NT - 1] * D1;
    const double XWL2 = PEACH2[(N + 1) * 6 + NT] * D + PEACH2[N * 6 + NT] * D1;
    return exp(XWL1 * (1. - DT) + XWL2 * DT) * 6.;
}

MALI_EOS_HD double Ca2OP(double FREQ, double TKEV)
{
    const double C1218 = 10. * exp(-1.697 / TKEV);
    const double C1420 = 6. * exp(-3.142 / TKEV);
    double X1044 = 0., X1218 = 0., X1420 = 0.;
    if (FREQ >= 2.870454e15) {
        const double XXX = cube(2.870454e15 / FREQ);
        X1044 = 1.08e-19 * XXX;
    }
    if (FREQ >= 2.460127e15) X1218 = 1.64e-17 * sqrt(2.460127e15 / FREQ);
    if (FREQ >= 2.110779e15) X1420 = SEATON(2.110779e15, 4.13e-18, 3., 0.69, FREQ);
    return X1044 + X1218 * C1218 + X1420 * C1420;
}

MALI_EOS_HD double LUKEOP(double XN1, double XO1, double XMg2, double XSi2, double XCa2, double STIM, double FREQ,
                          double FREQLG, double T, double TLOG, double TKEV)
{
    return (N1OP(FREQ, TKEV) * XN1 + O1OP(FREQ) * XO1 + Mg2OP(FREQ, TKEV) * XMg2 + Si2OP(FREQ, FREQLG, T, TLOG) * XSi2 +
            Ca2OP(FREQ, TKEV) * XCa2) *
           STIM;
}

MALI_EOS_HD double ELECOP(double XNE) { return 0.6653E-24 * XNE; }

MALI_EOS_HD double H2RAOP(double XH1, double FREQ, double T, double TKEV, double TLOG)
{
    const double q = 2.997925E18 / (FREQ < 2.922E15 ? FREQ : 2.922E15);
    const double WW = q * q;
    const double WW2 = WW * WW;
    const double SIG = (8.14E-13 + 1.28e-6 / WW + 1.61e0 / WW2) / WW2;
    const double ARG =
        4.477 / TKEV - 4.6628E1 + (1.8031E-3 + (-5.023E-7 + (8.1424E-11 - 5.0501E-15 * T) * T) * T) * T - 1.5 * TLOG;
    const double H1 = XH1 * 2.0;
    if (ARG > -80.0) return exp(ARG) * H1 * H1 * SIG;
    return 0.0;
}

// one wavelength of cop (witt.py:1321-1362): opacity and its scattering part, per cm; n = background_partials
MALI_EOS_HD void cop_one(double T, double TKEV, double HKT, double TLOG, double XNE, double WL, const double *n,
                         double &opacity, double &scatter)
{
    const double FREQ = 2.997925E18 / WL;
    const double FREQLG = log(FREQ);
    const double FREQ15 = FREQ * 1.E-15;
    const double EHVKT = exp(-FREQ * HKT);
    const double STIM = 1.0 - EHVKT;
    double ACOOL = 0.0, ALUKE = 0.0;
    const double H1 = n[0], H2 = n[1], HMIN = n[2], HE1 = n[3], HE2 = n[4], HE3 = n[5];
    const double AHYD = HOP(XNE, H1, H2, FREQ, FREQLG, T, TLOG, TKEV, STIM, EHVKT);
    const double AH2P = H2PLOP(H1, H2, FREQ, FREQLG, FREQ15, TKEV, STIM);
    const double AHMIN = HMINOP(H1, HMIN, FREQ, T, TKEV, XNE, EHVKT);
    const double SIGH = HRAYOP(H1, FREQ);
    const double AHE1 = HE1OP(HE1, HE2, XNE, FREQ, FREQLG, T, TKEV, TLOG, EHVKT, STIM);
    const double AHE2 = HE2OP(HE2, HE3, XNE, FREQ, FREQLG, T, TKEV, TLOG, EHVKT, STIM);
    const double AHEMIN = HEMIOP(HE1, FREQ, T, XNE);
    const double SIGHE = HERAOP(HE1, FREQ);
    // cop's argument order: C1, AL1, SI1, SI2, CA1, CA2, MG1, MG2, FE1, N1, O1 = n[6..16]
    if (T < 12000.) ACOOL = COOLOP(n[6], n[12], n[7], n[8], n[14], STIM, FREQ, FREQLG, T, TLOG, TKEV, HKT);
    if (T < 30000.) ALUKE = LUKEOP(n[15], n[16], n[13], n[9], n[11], STIM, FREQ, FREQLG, T, TLOG, TKEV);
    const double AHOT = 0.0;
    const double SIGEL = ELECOP(XNE);
    const double SIGH2 = H2RAOP(H1, FREQ, T, TKEV, TLOG);
    const double A = AHYD + AHMIN + AH2P + AHE1 + AHE2 + AHEMIN + ACOOL + ALUKE + AHOT;
    const double B = SIGH + SIGHE + SIGEL + SIGH2;
    opacity = A + B;
    scatter = B;
}

}  // namespace eos
}  // namespace mali

#undef Z4LOG
#undef A0
#undef A1c
#undef B1c
#undef C1c
#undef G0
#undef HEFREQ0
#undef CHI0
#undef PEACH0
#undef FREQMG
#undef FLOG0
#undef TLG0
#undef PEACH1
#undef FREQSI1
#undef FLOG1
#undef TLG1
#undef G1
#undef E1
#undef WNO1
#undef PEACH2
#undef FREQSI2
#undef FLOG2
#undef TLG2
