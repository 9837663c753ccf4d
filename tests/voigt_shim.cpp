// Host shim around lightspinner_b200/csrc/mali_voigt.h for tests/test_voigt_host.py (unit test only: the product
// evaluates this function inside compute_phi_kernel on the GPU).
#include "../lightspinner_b200/csrc/mali_voigt.h"

extern "C" void shim_voigt(int n, const double *a, const double *v, double *out)
{
    for (int i = 0; i < n; ++i) out[i] = mali::voigt_H(a[i], v[i]);
}
