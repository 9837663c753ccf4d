// Host shim around lightspinner_b200/csrc/mali_solve.h for tests/test_solver_host.py (unit test only:
// the product always runs this code inside stat_equil_kernel on the GPU).
#include "../lightspinner_b200/csrc/mali_solve.h"

extern "C" int shim_solve8(const double *G, int NL, int iEl, double nTot, double *x)
{
    return mali::solve_stat_equil<8>(G, 1, NL, iEl, nTot, x) ? 0 : 1;
}
extern "C" int shim_solve16(const double *G, int NL, int iEl, double nTot, double *x)
{
    return mali::solve_stat_equil<16>(G, 1, NL, iEl, nTot, x) ? 0 : 1;
}
extern "C" int shim_solve_any(const double *G, int NL, int iEl, double nTot, double *x)
{
    return mali::solve_stat_equil_any<16>(G, 1, NL, iEl, nTot, x) ? 0 : 1;
}
