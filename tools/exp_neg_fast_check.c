/* CPU check of exp_neg_fast (csrc/filter.cu): same operations with libm's fma, against expl().
 * Build: gcc -O2 -o /tmp/exp_fast tools/exp_neg_fast_check.c -lm   (table generated like
 * csrc/exp2_tab512.cuh).  Prints the largest relative error over 2e7 arguments in (-700, 0]. */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static double tab[512];

static double exp_neg_fast(double x) {
    const double MAGIC = 6755399441055744.0;
    const double t = fma(x, 738.6598609351493, MAGIC);
    int64_t bits; memcpy(&bits, &t, 8);
    const int32_t n = (int32_t)(bits & 0xffffffff);
    const double nd = t - MAGIC;
    const double r = fma(nd, -0.0013538030870311431, x);
    double q = fma(r, 1.0 / 6.0, 0.5);
    q = q * r;
    const double p = fma(q, r, r);
    const double T = tab[n & 511];
    const double v = fma(T, p, T);
    int64_t vb; memcpy(&vb, &v, 8);
    vb += ((int64_t)(n >> 9)) << 52;
    double out; memcpy(&out, &vb, 8);
    return out;
}

int main(void) {
    for (int i = 0; i < 512; ++i) tab[i] = (double)exp2l((long double)i / 512.0L);
    double worst = 0.0, worst_x = 0.0;
    srand(12345);
    for (long k = 0; k < 20000000; ++k) {
        double u = (rand() + rand() / (double)RAND_MAX) / (double)RAND_MAX;
        double x = (k & 1) ? -700.0 * u : -40.0 * u * u;
        long double ref = expl((long double)x);
        double rel = (double)fabsl(((long double)exp_neg_fast(x) - ref) / ref);
        if (rel > worst) { worst = rel; worst_x = x; }
    }
    printf("max relative error %.3e at x = %.17g\n", worst, worst_x);
    printf("exp_neg_fast(0) = %.17g, (-1e-17) = %.17g, (1e-16) = %.17g\n", exp_neg_fast(0.0),
           exp_neg_fast(-1e-17), exp_neg_fast(1e-16));
    return worst < 1e-13 ? 0 : 1;
}
