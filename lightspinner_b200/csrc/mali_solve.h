// mali_solve.h -- the per-(column, atom, depth) statistical-equilibrium solve (rh_method.py:722-745), written so
// that the SAME source compiles for the device (stat_equil_kernel) and, for unit tests only, for the host
// (tests/test_solver_host.py builds a g++ shim around it; the product never runs it on the CPU).
//
// The reference calls scipy.linalg.solve (LAPACK getrf/getrs).  The systems have cond_2 ~ 1e6..3e9 (SURVEY.md
// 7.3-1), so instead of cloning LAPACK's rounding we solve to the correctly-rounded solution of the fp64 system:
// LU with partial pivoting + iterative refinement with the residual accumulated in double-double.  One
// refinement step already returns the exact solution rounded to fp64 on the CaII/FALC systems; two are run.
#pragma once
#include <cfloat>
#include <cmath>

#if defined(__CUDACC__)
#define MALI_HD __host__ __device__ __forceinline__
#else
#define MALI_HD inline
#endif

namespace mali {

MALI_HD double fma_exact(double a, double b, double c)
{
#if defined(__CUDA_ARCH__)
    return __fma_rn(a, b, c);
#else
    return std::fma(a, b, c);
#endif
}

MALI_HD void two_sum(double a, double b, double &s, double &e)
{
    s = a + b;
    const double bb = s - a;
    e = (a - (s - bb)) + (b - bb);
}

// Solves A x = b with A = Gamma[:, :] (row-major, leading dimension ldg, element stride gstride) whose row
// iEl is replaced by ones, and b = nTot * e_iEl.  Returns false for a singular / non-finite system.
template <int NLMAX>
MALI_HD bool solve_stat_equil(const double *G, long long gstride, int NL, int iEl, double nTot, double *x)
{
    double A[NLMAX * NLMAX], LU[NLMAX * NLMAX], r[NLMAX];
    int piv[NLMAX];
    for (int i = 0; i < NL; ++i)
        for (int j = 0; j < NL; ++j) {
            const double v = (i == iEl) ? 1.0 : G[(long long)(i * NL + j) * gstride];
            A[i * NLMAX + j] = v;
            LU[i * NLMAX + j] = v;
        }
    // P A = L U, rows swapped in full (LAPACK convention: earlier columns of L follow later swaps)
    for (int j = 0; j < NL; ++j) {
        int pr = j;
        double best = fabs(LU[j * NLMAX + j]);
        for (int i = j + 1; i < NL; ++i) {
            const double v = fabs(LU[i * NLMAX + j]);
            if (v > best) {
                best = v;
                pr = i;
            }
        }
        piv[j] = pr;
        if (!(best > 0.0) || !(best <= DBL_MAX)) return false;
        if (pr != j)
            for (int q = 0; q < NL; ++q) {
                const double tmp = LU[j * NLMAX + q];
                LU[j * NLMAX + q] = LU[pr * NLMAX + q];
                LU[pr * NLMAX + q] = tmp;
            }
        const double pv = LU[j * NLMAX + j];
        for (int i = j + 1; i < NL; ++i) {
            const double l = LU[i * NLMAX + j] / pv;
            LU[i * NLMAX + j] = l;
            for (int q = j + 1; q < NL; ++q) LU[i * NLMAX + q] -= l * LU[j * NLMAX + q];
        }
    }
    auto lu_solve = [&](double *b) {
        for (int j = 0; j < NL; ++j) {  // b <- P b: ALL swaps first (they were applied to whole rows of L)
            const int pr = piv[j];
            if (pr != j) {
                const double tmp = b[j];
                b[j] = b[pr];
                b[pr] = tmp;
            }
        }
        for (int j = 0; j < NL; ++j)
            for (int i = j + 1; i < NL; ++i) b[i] -= LU[i * NLMAX + j] * b[j];
        for (int i = NL - 1; i >= 0; --i) {
            double acc = b[i];
            for (int q = i + 1; q < NL; ++q) acc -= LU[i * NLMAX + q] * b[q];
            b[i] = acc / LU[i * NLMAX + i];
        }
    };
    for (int i = 0; i < NL; ++i) x[i] = 0.0;
    x[iEl] = nTot;
    lu_solve(x);
    for (int it = 0; it < 2; ++it) {
        for (int i = 0; i < NL; ++i) {  // r = b - A x, accumulated in double-double
            double hi = (i == iEl) ? nTot : 0.0, lo = 0.0;
            for (int j = 0; j < NL; ++j) {
                const double aij = A[i * NLMAX + j];
                const double ph = -(aij * x[j]);
                const double pl = -fma_exact(aij, x[j], ph);  // -(a*x) == ph + pl exactly
                double s, e;
                two_sum(hi, ph, s, e);
                hi = s;
                lo += e + pl;
            }
            r[i] = hi + lo;
        }
        lu_solve(r);
        for (int i = 0; i < NL; ++i) x[i] += r[i];
    }
    for (int i = 0; i < NL; ++i)
        if (!(fabs(x[i]) <= DBL_MAX)) return false;
    return true;
}

}  // namespace mali
