/* CPU check of the error bound of the fp32 screening mean (safe_learning_b200/csrc/gp_mean_staged.cuh,
 * mean32_factor): the kernel's arithmetic restated with C floats (fmaf chains, sixteen terms per fp32
 * accumulator flushed into a double, E = 2^-Z in fp32), 2^x from exp2f with a worst-sign perturbation of
 * 2^-22 standing in for ex2.approx (PTX ISA: maximum relative error 2^-22), against the mean summed in long
 * double -- over random centres, points, training rows and gammas in several magnitude regimes.  Prints the
 * largest |error| / bound per regime (must stay below 1; typical values are a few percent, the bound is a
 * worst-case sum).      gcc -O2 -o /tmp/sbc tools/screening_bound_check.c -lm && /tmp/sbc          */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#define DIN 3
#define M 512
#define FLUSH 16

static double urand(void) { return (double)rand() / RAND_MAX; }
static double srnd(double r) { return (2.0 * urand() - 1.0) * r; }

int main(void) {
    const double S = 1.2011224087864498, u24 = 5.9604644775390625e-8;
    const double ranges[] = {1.0, 5.0, 30.0, 300.0};
    const double spans[] = {0.05, 0.7, 3.0, 7.0};
    int bad = 0;
    srand(12345);
    for (int ir = 0; ir < 4; ++ir)
        for (int is = 0; is < 4; ++is) {
            double worst = 0.0, zmax = 0.0;
            for (int trial = 0; trial < 400; ++trial) {
                double cs[DIN], xs[M][DIN], gam[M];
                for (int c = 0; c < DIN; ++c) cs[c] = srnd(ranges[ir]);
                const double gscale = pow(10.0, srnd(4.0));
                for (int j = 0; j < M; ++j) {
                    /* rows near the centre (large kernel values) and far away, gammas of mixed size */
                    const double rr = (j & 1) ? spans[is] * 1.5 : ranges[ir];
                    for (int c = 0; c < DIN; ++c) xs[j][c] = cs[c] + srnd(rr);
                    gam[j] = srnd(gscale) * ((j % 7 == 0) ? 100.0 : 1.0);
                }
                float xf[M][DIN + 1], gf[M];
                double gsum = 0.0;
                for (int j = 0; j < M; ++j) {
                    double hh = 0.0;
                    for (int c = 0; c < DIN; ++c) {
                        const double xcd = (xs[j][c] - cs[c]) * S;
                        xf[j][c] = (float)xcd;
                        hh = fma(xcd, xcd, hh);
                    }
                    xf[j][DIN] = (float)(-0.5 * hh);
                    gf[j] = (float)gam[j];
                    gsum += fabs(gam[j]);
                }
                for (int p = 0; p < 16; ++p) {
                    double zs[DIN], Z = 0.0;
                    float zc[DIN];
                    for (int c = 0; c < DIN; ++c) {
                        zs[c] = cs[c] + srnd(spans[is]);
                        const double zcd = (zs[c] - cs[c]) * S;
                        zc[c] = (float)zcd;
                        Z = fma(zcd, zcd, Z);
                    }
                    Z *= 0.5;
                    if (Z > 40.0) continue;                 /* left to the fp64 stages */
                    if (Z > zmax) zmax = Z;
                    long double exact = 0.0L;
                    for (int j = 0; j < M; ++j) {
                        long double t = 0.0L;
                        for (int c = 0; c < DIN; ++c) {
                            const long double d = (long double)zs[c] - (long double)xs[j][c];
                            t += d * d;
                        }
                        exact += (long double)gam[j] * expl(-0.5L * t);
                    }
                    float acc[4] = {0.f, 0.f, 0.f, 0.f};
                    double dot = 0.0;
                    for (int j0 = 0; j0 < M; j0 += 4) {
                        for (int u = 0; u < 4; ++u) {
                            const int j = j0 + u;
                            float arg = xf[j][DIN];
                            for (int c = 0; c < DIN; ++c) arg = fmaf(zc[c], xf[j][c], arg);
                            float k = exp2f(arg);
                            k *= (rand() & 1) ? (1.0f + 2.384185791015625e-7f) : (1.0f - 2.384185791015625e-7f);
                            acc[u] = fmaf(k, gf[j], acc[u]);
                        }
                        if ((j0 / 4 + 1) % FLUSH == 0) {
                            dot += ((double)acc[0] + (double)acc[1]) + ((double)acc[2] + (double)acc[3]);
                            acc[0] = acc[1] = acc[2] = acc[3] = 0.f;
                        }
                    }
                    dot += ((double)acc[0] + (double)acc[1]) + ((double)acc[2] + (double)acc[3]);
                    float e32 = exp2f(-(float)Z);
                    e32 *= (rand() & 1) ? (1.0f + 2.384185791015625e-7f) : (1.0f - 2.384185791015625e-7f);
                    const double mean32 = (double)e32 * dot;
                    const double epsrel = u24 * (0.7 * (DIN + 2.01) * (2.13 + 5.0 * Z) + 11.0 + FLUSH + 8.0 + 0.7 * Z);
                    const double bound = 1.05 * epsrel * gsum + 1.3e-26 * M;
                    const double err = fabs((double)((long double)mean32 - exact));
                    if (err / bound > worst) worst = err / bound;
                    if (!(err <= bound)) ++bad;
                }
            }
            printf("range %6.1f span %5.2f  max Z %6.2f  max |error| / bound %.4f\n", ranges[ir], spans[is], zmax, worst);
        }
    printf(bad ? "FAILED: %d means outside the bound\n" : "ok: every mean inside its bound (%d violations)\n", bad);
    return bad != 0;
}
