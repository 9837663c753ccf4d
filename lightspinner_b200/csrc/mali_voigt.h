// mali_voigt.h -- Voigt function H(a, v) = Re w(v + i a) (the Faddeeva function), host/device source.
//
// Used by the device version of ComputationalTransition.compute_phi (rh_method.py:198-243; utils.py:13-15 calls
// scipy.special.wofz).  Not a port of scipy's Faddeeva package: H is evaluated from the integral representation
//     H(a, v) = (a / pi) * integral exp(-t^2) / ((v - t)^2 + a^2) dt
// with the trapezoidal rule of step h = 1/2 plus the residue of the integrand's pole at t = v + i a (the rule's only
// error term larger than exp(-pi^2/h^2) ~ 7e-18 when a < pi/h):
//     H = (h a / pi) * sum_n exp(-t_n^2) / ((v - t_n)^2 + a^2)  +  Re[ 2 exp(-z^2) / (1 -+ exp(-2 pi i z / h)) ]
// on the nodes t_n = n h ("-") or t_n = (n + 1/2) h ("+").  Both grids are exact to rounding; the one whose nearest
// node is at least h/4 away from v is used, which keeps every term and the residue well conditioned for any a > 0
// (no cancellation as a -> 0).  Beyond |z| = 16 the asymptotic series of w(z) takes over (8 terms, 4x cheaper).
// Agreement with scipy.special.wofz: <= 3e-14 relative for 1e-7 <= a <= 30,
// 0 <= |v| <= 1e4 (tests/test_voigt_host.py), i.e. at the level of wofz's own accuracy.
#pragma once
#include <cmath>

#ifdef __CUDACC__
#define MALI_VOIGT_HD __host__ __device__ __forceinline__
#else
#define MALI_VOIGT_HD inline
#endif

namespace mali {

constexpr int kVoigtTerms = 12;  // nodes n = -12 .. 12 (t up to 6.25: exp(-t^2) < 2e-17)

// (h / pi) * exp(-(n h)^2), n = 0 .. 12, and (h / pi) * exp(-((n + 1/2) h)^2), n = 0 .. 12   (h = 1/2)
MALI_VOIGT_HD double voigt_weight_int(int n)
{
    constexpr double w[kVoigtTerms + 1] = {
        1.5915494309189535e-01, 1.2394999430965298e-01, 5.8549831524319168e-02, 1.6774807587073417e-02,
        2.9150244650281935e-03, 3.0724131819283507e-04, 1.9641280346397441e-05, 7.6157508623233106e-07,
        1.7910529328280185e-08, 2.5547997977257987e-10, 2.2103349154917858e-12, 1.1598773137396176e-14,
        3.6916352404776737e-17};
    return w[n];
}
MALI_VOIGT_HD double voigt_weight_half(int n)
{
    constexpr double w[kVoigtTerms + 1] = {
        1.4951223255186186e-01, 9.0683753044789428e-02, 3.3360688393446213e-02, 7.4437757438915184e-03,
        1.0074054986493862e-03, 8.2692878970342922e-05, 4.1170360188319614e-06, 1.2432371522416446e-07,
        2.2770682733516199e-09, 2.5295943566004535e-11, 1.7044272703959557e-13, 6.9656046875934636e-16,
        1.7266007781169686e-18};
    return w[n];
}

MALI_VOIGT_HD double voigt_rcp(double d)
{
#ifdef __CUDA_ARCH__
    // 1/d to ~1 ulp: hardware seed + two Newton steps (d is a sum of squares in [1e-14, 1e9]: no special cases)
    double r = __hiloint2double(0, 0);
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
    r = fma(fma(-d, r, 1.0), r, r);
    r = fma(fma(-d, r, 1.0), r, r);
    return r;
#else
    return 1.0 / d;
#endif
}

// What depends on the damping parameter alone (one value per line and depth point: the callers evaluate H for many
// frequencies with the same a): a^2, whether the pole's residue is needed at all, and exp(2 pi a / h).
struct VoigtPre {
    double y, y2, E;
    bool res;
};
MALI_VOIGT_HD VoigtPre voigt_pre(double a)
{
    constexpr double h = 0.5, pi = 3.141592653589793;
    VoigtPre P;
    P.y = a;
    P.y2 = a * a;
    P.res = a < pi / h;
    P.E = P.res ? exp(2.0 * pi * a / h) : 0.0;
    return P;
}

// 1 / A + 1 / B with one reciprocal: (A + B) / (A B).  A, B = (v -+ t)^2 + a^2 lie in [1/64, 1e3] on the trapezoid
// branch (the grid is chosen so that no node is closer than h / 4 to v), so the product cannot leave the normal range.
MALI_VOIGT_HD double voigt_pair(double dm, double dp, double y2)
{
    const double A = fma(dm, dm, y2), B = fma(dp, dp, y2);
    return (A + B) * voigt_rcp(A * B);
}

MALI_VOIGT_HD double voigt_H_pre(const VoigtPre &P, double v)
{
    constexpr double h = 0.5, pi = 3.141592653589793;
    const double x = fabs(v), y = P.y, y2 = P.y2;
    if (fma(x, x, y2) > 256.0) {
        // |z| > 16 (a third of a line's wavelength points): asymptotic series of w(z) = i / (sqrt(pi) z) *
        // sum_k (2k-1)!! / (2 z^2)^k, 8 terms (the next one is < 5e-16), complex Horner in u = 1 / (2 z^2)
        const double zr = x * x - y2, zi = 2.0 * x * y;              // z^2
        const double n2 = voigt_rcp(fma(zr, zr, zi * zi));
        const double ur = 0.5 * zr * n2, ui = -0.5 * zi * n2;        // u = conj(z^2) / (2 |z^2|^2)
        double sr = 1.0, si = 0.0;
#pragma unroll
        for (int k = 8; k >= 1; --k) {
            const double c = 2.0 * k - 1.0;
            const double tr = fma(ur, sr, -ui * si), ti = fma(ur, si, ui * sr);
            sr = fma(c, tr, 1.0);
            si = c * ti;
        }
        // Re[i S / z] = -Im[S conj(z)] / |z|^2 = (sr y - si x) / |z|^2
        return (sr * y - si * x) * voigt_rcp(fma(x, x, y2)) * 0.5641895835477563;   // 1 / sqrt(pi)
    }
    const double r = x / h;
    const double fl = floor(r);
    const double frac = r - fl;
    const bool half = frac < 0.25 || frac > 0.75;        // v is near an integer node -> use the half-integer grid
    // One loop serves both grids -- nodes t_n = (n + off) h with off = 0 or 1/2, weights selected per lane -- so that a
    // warp whose lanes sit on different grids (they do: neighbouring frequencies alternate) does not run two 13-term
    // sums one after the other.  The two signs of a node share one reciprocal (voigt_pair); the integer grid's node 0
    // is its own mirror image, hence half its weight.
    double s = 0.0;
    const double off = half ? 0.5 * h : 0.0;
#pragma unroll
    for (int n = 0; n <= kVoigtTerms; ++n) {
        const double t = n * h + off;
        const double wi = (n == 0 ? 0.5 : 1.0) * voigt_weight_int(n);
        const double w = half ? voigt_weight_half(n) : wi;
        s += w * voigt_pair(x - t, x + t, y2);
    }
    s *= y;
    // residue of the pole at t = x + i y (inside the strip of analyticity the rule needs only when y < pi / h).  Its
    // size is at most sqrt(2) exp(y^2 - x^2) -- the grid choice keeps |1 -+ exp(-2 pi i z / h)| >= sqrt(2) -- against
    // H >= ~ y / (sqrt(pi) (x^2 + y^2)): below 1e-17 of H once x^2 - y^2 > 70 for every a >= 1e-10
    if (P.res && x * x - y2 < 70.0) {
        const double A = exp(y2 - x * x);                 // |exp(-z^2)|
        double st, ct, sp, cp;
        const double th = 2.0 * x * y;                    // exp(-z^2) = A (ct - i st)
        if (th < 0.25) {
            // chromospheric damping parameters (a ~ 1e-4 .. 1e-2) keep this angle small: Taylor series to t^13 / t^12
            // (next terms < 1e-19), a fifth of the cost of the library's sincos
            const double u = th * th;
            double ps = fma(u, 1.0 / 6227020800.0, -1.0 / 39916800.0);
            ps = fma(u, ps, 1.0 / 362880.0);
            ps = fma(u, ps, -1.0 / 5040.0);
            ps = fma(u, ps, 1.0 / 120.0);
            ps = fma(u, ps, -1.0 / 6.0);
            st = fma(th * u, ps, th);
            double pc = fma(u, 1.0 / 479001600.0, -1.0 / 3628800.0);
            pc = fma(u, pc, 1.0 / 40320.0);
            pc = fma(u, pc, -1.0 / 720.0);
            pc = fma(u, pc, 1.0 / 24.0);
            pc = fma(u, pc, -0.5);
            ct = fma(u, pc, 1.0);
        } else {
            sincos(th, &st, &ct);
        }
#ifdef __CUDA_ARCH__
        sincospi(2.0 * frac, &sp, &cp);                   // exp(-2 pi i z / h) = E (cp - i sp), x / h = fl + frac
#else
        sincos(2.0 * pi * frac, &sp, &cp);
#endif
        const double sg = half ? 1.0 : -1.0;
        const double dr = 1.0 + sg * P.E * cp, di = -sg * P.E * sp;
        s += 2.0 * A * (ct * dr - st * di) * voigt_rcp(dr * dr + di * di);     // |den|^2 >= 2 (see above)
    }
    return s;
}

MALI_VOIGT_HD double voigt_H(double a, double v) { return voigt_H_pre(voigt_pre(a), v); }

}  // namespace mali
