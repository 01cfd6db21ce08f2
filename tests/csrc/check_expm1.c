/* Exhaustive check of psn_expm1_neg (py_psnode_b200/csrc/psnode_math.cuh) over EVERY float x <= 0 against
 * the correctly rounded double-precision expm1.  Prints "max_ulp <v> at <x>" and exits non-zero above 1.0 ulp... */
#include <stdio.h>
#include <stdlib.h>
#include "../../py_psnode_b200/csrc/psnode_math.cuh"

int main(void) {
    double worst = 0.0; float worst_x = 0.0f; long long n = 0, n_inexact = 0;
    /* negative floats: bit patterns 0x80000000 (-0) .. 0xff800000 (-inf) */
#pragma omp parallel
    {
        double lw = 0.0; float lx = 0.0f; long long ln = 0, li = 0;
#pragma omp for schedule(static) nowait
        for (long long b = 0x80000000LL; b <= 0xff800000LL; b++) {
            unsigned u = (unsigned)b; float x; memcpy(&x, &u, 4);
            float got = psn_expm1_neg(x);
            double ref = expm1((double)x);
            float reff = (float)ref;
            /* ulp of the reference result */
            double ulp;
            if (reff == 0.0f || fabsf(reff) < 1.17549435e-38f) ulp = 1.4012984643e-45;
            else { int e; frexpf(reff, &e); ulp = ldexp(1.0, e - 24); }
            double err = fabs((double)got - ref) / ulp;
            if (err > lw) { lw = err; lx = x; }
            if (got != reff) li++;
            ln++;
        }
#pragma omp critical
        { if (lw > worst) { worst = lw; worst_x = lx; } n += ln; n_inexact += li; }
    }
    printf("checked %lld values, max_ulp %.4f at x=%.9g, not-correctly-rounded %.4f%%\n", n, worst, worst_x, 100.0 * n_inexact / n);
    /* spot values the reference relies on: elu(-1e-8) == -1e-8 */
    if (psn_elu(-1e-8f) != -1e-8f) { printf("elu(-1e-8) wrong\n"); return 2; }
    if (psn_elu(2.5f) != 2.5f || psn_elu(0.0f) != 0.0f) { printf("elu(+) wrong\n"); return 2; }
    if (psn_elu(-100.0f) != -1.0f) { printf("elu(-100) wrong\n"); return 2; }
    return worst <= 1.0 ? 0 : 1;
}
