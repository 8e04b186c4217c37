// gp_mean.cuh -- mean-only GP posterior of a FunctionStack at one query point per thread
// (reinforcement_learning.py:98-99: PolicyIteration uses the mean only).  Shared by the Bellman
// kernels (light.cu) and the decision filter of the Lyapunov sweep (filter.cu).
#pragma once
#include "common.cuh"

// One factor with NO outputs on it: NO is a compile-time constant so the running dot products
// stay in registers (a runtime-bounded loop over outputs would push them to local memory).
constexpr int BCHUNK = 256;    // training rows staged per pass in the Bellman kernels

template <int DIN, int NO>
SLB_DEV void gp_mean_factor(const slb_gp_stack& gp, const slb_gp_factor& F, const int* outs,
                            const double* z, double* mu, const double* exptab, double* stage) {
    const bool general = F.kernel.num_prims > 0;     // covariance expression on the raw inputs
    double zs[DIN];
#pragma unroll
    for (int c = 0; c < DIN; ++c) zs[c] = general ? z[c] : z[c] / F.lengthscales[c];
    double dot[NO];
    const double* gam[NO];
#pragma unroll
    for (int q = 0; q < NO; ++q) { dot[q] = 0.0; gam[q] = gp.outputs[outs[q]].gamma; }
    const double* __restrict__ Xs = F.Xs;
    const int M = F.M;
    // The training inputs and gamma are staged chunk-wise in shared memory (one coalesced pass
    // per block): every thread needs every row once, and read from global the first toucher
    // of a row pays an L2 round trip inside the exp dependency chain.  All threads of the
    // block take part (callers must not exit early).
    double* xch = stage;                       // [BCHUNK][DIN]
    double* gch = stage + BCHUNK * DIN;        // [NO][BCHUNK]
    for (int c0 = 0; c0 < M; c0 += BCHUNK) {
        const int nc = min(BCHUNK, M - c0);
        __syncthreads();
        for (int i = threadIdx.x; i < nc * DIN; i += blockDim.x) xch[i] = Xs[(size_t)c0 * DIN + i];
#pragma unroll
        for (int q = 0; q < NO; ++q)
            for (int i = threadIdx.x; i < nc; i += blockDim.x) gch[q * BCHUNK + i] = gam[q][c0 + i];
        __syncthreads();
        // 4 independent exp chains per thread (the loop is bound by the fp64 pipe through exp)
        for (int j0 = 0; j0 < nc; j0 += 4) {
            double kv[4];
            if (general) {
                const double* xr[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) xr[u] = xch + min(j0 + u, nc - 1) * DIN;
                kernel_expr_cross_n<DIN, 4>(F.kernel, zs, xr, exptab, kv);
            } else {
                double t2[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const double* xr = xch + min(j0 + u, nc - 1) * DIN;
                    double acc = 0.0;
#pragma unroll
                    for (int c = 0; c < DIN; ++c) { const double df = zs[c] - xr[c]; acc = fma(df, df, acc); }
                    t2[u] = acc;
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) kv[u] = F.variance * exp_neg_tab(-0.5 * t2[u], exptab);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int j = min(j0 + u, nc - 1);
                const double k = j0 + u >= nc ? 0.0 : kv[u];
#pragma unroll
                for (int q = 0; q < NO; ++q) dot[q] = fma(k, gch[q * BCHUNK + j], dot[q]);
            }
        }
    }
    const double s2 = f64mul(F.scale, F.scale);
#pragma unroll
    for (int q = 0; q < NO; ++q) {
        const slb_gp_output& G = gp.outputs[outs[q]];
        double mx = 0.0;
        if (G.prior_mean != nullptr) {
            mx = f64mul(z[0], G.prior_mean[0]);
#pragma unroll
            for (int c = 1; c < DIN; ++c) mx = f64add(mx, f64mul(z[c], G.prior_mean[c]));
            mx = f64mul(F.scale, mx);
        }
        mu[outs[q]] = f64add(f64mul(s2, dot[q]), mx) / F.scale;
    }
}

// mean of the GP stack at z (mean only, reinforcement_learning.py:98-99):
//   mean_o = (scale^2 sum_j k_j gamma_o,j + scale m_o(z)) / scale,  gamma = L^-T alpha,
// which equals a^T alpha of functions.py:441-442 up to rounding.
template <int DIN>
SLB_DEV void gp_mean_only(const slb_gp_stack& gp, const double* z, double* mu,
                          const double* exptab, double* stage) {
    for (int f = 0; f < gp.num_factors; ++f) {
        const slb_gp_factor& F = gp.factors[f];
        int outs[SLB_MAX_OUT];
        int no = 0;
        for (int o = 0; o < gp.num_outputs; ++o)
            if (gp.outputs[o].factor == f) outs[no++] = o;
        switch (no) {
        case 1: gp_mean_factor<DIN, 1>(gp, F, outs, z, mu, exptab, stage); break;
        case 2: gp_mean_factor<DIN, 2>(gp, F, outs, z, mu, exptab, stage); break;
        case 3: gp_mean_factor<DIN, 3>(gp, F, outs, z, mu, exptab, stage); break;
        case 4: gp_mean_factor<DIN, 4>(gp, F, outs, z, mu, exptab, stage); break;
        case 5: gp_mean_factor<DIN, 5>(gp, F, outs, z, mu, exptab, stage); break;
        case 6: gp_mean_factor<DIN, 6>(gp, F, outs, z, mu, exptab, stage); break;
        default: break;
        }
    }
}

