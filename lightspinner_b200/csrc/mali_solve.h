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

// The same algorithm, same operation order (hence the same bits), for a compile-time size: every loop unrolls and
// every array index is static (row exchanges are done with selects), so the matrices live in registers instead of
// local memory -- the latency of this solve matters for small launches (one column: it was 30 % of an iteration).
template <int NL>
MALI_HD bool solve_stat_equil_fixed(const double *G, long long gstride, int iEl, double nTot, double *x)
{
    double A[NL * NL], LU[NL * NL], r[NL];
    int piv[NL];
#pragma unroll
    for (int i = 0; i < NL; ++i)
#pragma unroll
        for (int j = 0; j < NL; ++j) {
            const double v = (i == iEl) ? 1.0 : G[(long long)(i * NL + j) * gstride];
            A[i * NL + j] = v;
            LU[i * NL + j] = v;
        }
    bool ok = true;
#pragma unroll
    for (int j = 0; j < NL; ++j) {
        int pr = j;
        double best = fabs(LU[j * NL + j]);
#pragma unroll
        for (int i = j + 1; i < NL; ++i) {
            const double v = fabs(LU[i * NL + j]);
            if (v > best) {
                best = v;
                pr = i;
            }
        }
        piv[j] = pr;
        if (!(best > 0.0) || !(best <= DBL_MAX)) ok = false;
#pragma unroll
        for (int i = j + 1; i < NL; ++i) {       // exchange rows j and pr (pr > j when different)
            const bool sw = (i == pr);
#pragma unroll
            for (int q = 0; q < NL; ++q) {
                const double a = LU[j * NL + q], b = LU[i * NL + q];
                LU[j * NL + q] = sw ? b : a;
                LU[i * NL + q] = sw ? a : b;
            }
        }
        const double pv = LU[j * NL + j];
#pragma unroll
        for (int i = j + 1; i < NL; ++i) {
            const double l = LU[i * NL + j] / pv;
            LU[i * NL + j] = l;
#pragma unroll
            for (int q = j + 1; q < NL; ++q) LU[i * NL + q] -= l * LU[j * NL + q];
        }
    }
    if (!ok) return false;
    auto lu_solve = [&](double *b) {
#pragma unroll
        for (int j = 0; j < NL; ++j) {
#pragma unroll
            for (int i = j + 1; i < NL; ++i) {
                const bool sw = (i == piv[j]);
                const double u = b[j], w = b[i];
                b[j] = sw ? w : u;
                b[i] = sw ? u : w;
            }
        }
#pragma unroll
        for (int j = 0; j < NL; ++j)
#pragma unroll
            for (int i = j + 1; i < NL; ++i) b[i] -= LU[i * NL + j] * b[j];
#pragma unroll
        for (int i = NL - 1; i >= 0; --i) {
            double acc = b[i];
#pragma unroll
            for (int q = i + 1; q < NL; ++q) acc -= LU[i * NL + q] * b[q];
            b[i] = acc / LU[i * NL + i];
        }
    };
#pragma unroll
    for (int i = 0; i < NL; ++i) x[i] = (i == iEl) ? nTot : 0.0;
    lu_solve(x);
#pragma unroll 1
    for (int it = 0; it < 2; ++it) {
#pragma unroll
        for (int i = 0; i < NL; ++i) {
            double hi = (i == iEl) ? nTot : 0.0, lo = 0.0;
#pragma unroll
            for (int j = 0; j < NL; ++j) {
                const double aij = A[i * NL + j];
                const double ph = -(aij * x[j]);
                const double pl = -fma_exact(aij, x[j], ph);
                double s2, e;
                two_sum(hi, ph, s2, e);
                hi = s2;
                lo += e + pl;
            }
            r[i] = hi + lo;
        }
        lu_solve(r);
#pragma unroll
        for (int i = 0; i < NL; ++i) x[i] += r[i];
    }
#pragma unroll
    for (int i = 0; i < NL; ++i)
        if (!(fabs(x[i]) <= DBL_MAX)) return false;
    return true;
}

// Dispatch: sizes 2 .. 8 take the register-resident form, anything else the general one.
template <int NLMAX>
MALI_HD bool solve_stat_equil_any(const double *G, long long gstride, int NL, int iEl, double nTot, double *x)
{
    switch (NL) {
        case 2: return solve_stat_equil_fixed<2>(G, gstride, iEl, nTot, x);
        case 3: return solve_stat_equil_fixed<3>(G, gstride, iEl, nTot, x);
        case 4: return solve_stat_equil_fixed<4>(G, gstride, iEl, nTot, x);
        case 5: return solve_stat_equil_fixed<5>(G, gstride, iEl, nTot, x);
        case 6: return solve_stat_equil_fixed<6>(G, gstride, iEl, nTot, x);
        case 7: return solve_stat_equil_fixed<7>(G, gstride, iEl, nTot, x);
        case 8: return solve_stat_equil_fixed<8>(G, gstride, iEl, nTot, x);
        default: return solve_stat_equil<NLMAX>(G, gstride, NL, iEl, nTot, x);
    }
}

}  // namespace mali
