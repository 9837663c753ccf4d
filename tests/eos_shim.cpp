// Host shim around lightspinner_b200/csrc/mali_eos.h for tests/test_eos_host.py (unit test only: the product evaluates
// these functions inside the background kernels on the GPU).
#include "../lightspinner_b200/csrc/mali_eos.h"

using namespace mali::eos;

extern "C" {
void shim_eos(const Tables *E, int n, const double *T, const double *rho, double *pgas, double *pe)
{
    for (int k = 0; k < n; ++k) {
        PointCache C;
        point_cache(*E, T[k], C);
        pgas[k] = pg_from_rho(*E, C, T[k], rho[k]);
        pe[k] = pe_from_rho(*E, C, T[k], rho[k]);
    }
}
void shim_partials(const Tables *E, int n, const double *T, const double *pgas, const double *pe, double *out17)
{
    for (int k = 0; k < n; ++k) {
        PointCache C;
        point_cache(*E, T[k], C);
        background_partials(*E, C, T[k], pgas[k], pe[k], out17 + 17 * k);
    }
}
// contOpacity (witt.py:744-768) per cm for nw wavelengths (Angstrom) at one depth point
void shim_cont_opacity(const Tables *E, double T, double pgas, double pe, int nw, const double *w, double *opac)
{
    const double TK = T * BK, TKEV = TK / EV, HTK = HH / TK, TLOG = log(T), xne = pe / TK;
    double n[17];
    PointCache C;
    point_cache(*E, T, C);
    background_partials(*E, C, T, pgas, pe, n);
    for (int i = 0; i < nw; ++i) {
        double sc;
        cop_one(T, TKEV, HTK, TLOG, xne, w[i], n, opac[i], sc);
    }
}
}
