// common.cuh -- shared device code of libslb200: error plumbing, GridWorld coordinates,
// the fused function objects (policy / V / Lipschitz / plants / reward / Triangulation),
// the Lyapunov decision formula and the order-preserving V key.
//
// Reference citations are relative to /root/reference (befelix/safe_learning @ f1aad5a).
// Bit-parity rule: the cheap element-wise pieces use __dmul_rn/__dadd_rn (never contracted
// into FMA) in the left-to-right order that oracle/reference_path.py writes out, so grid
// coordinates, linear maps, quadratic forms and barycentric weights are bit-identical to
// the CPU oracle.  Only the GP contraction (DMMA) and libm calls (exp/sin/cos/sqrt is exact)
// differ in rounding.
#pragma once
#include <stdlib.h>

#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#include "../../include/slb200.h"

// ----------------------------------------------------------------------------- host side
void slb_set_error(const char* fmt, ...);
void slb_count_launch();

#define SLB_CHECK(cond, ...)                                                     \
    do {                                                                         \
        if (!(cond)) { slb_set_error(__VA_ARGS__); return 1; }                   \
    } while (0)

#define SLB_CUDA(call)                                                           \
    do {                                                                         \
        cudaError_t e__ = (call);                                                \
        if (e__ != cudaSuccess) {                                                \
            slb_set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), \
                          __FILE__, __LINE__);                                   \
            return 2;                                                            \
        }                                                                        \
    } while (0)

// ---- programmatic dependent launch (PDL) --------------------------------------------------------
// The kernels of one sweep form a chain in which every kernel starts with work that does not depend
// on its predecessor (staging static tables by TMA, L2 prefetches, descriptor loads) -- and with
// launch latency and CTA ramp-up.  Launched with the programmatic-stream-serialization attribute a
// kernel may start as soon as every CTA of its predecessor has called pdl_launch_dependents() (or
// left); it must call pdl_wait() before it touches anything the predecessor produces (the wait
// returns once the predecessor grid has completed and its writes are visible).  Both calls are
// no-ops in a kernel launched the ordinary way.  MEASURED (C2, one B200, gpurun call 24): the step is
// 0.156 ms with the attribute against 0.152 ms without -- the dependents' early CTAs take shared memory
// and registers the predecessor's last CTAs still want, and the kernels' pre-dependency parts are short --
// so the attribute is OFF by default; SLB200_PDL=1 switches it on (all GPU tests pass either way).
#if defined(__CUDACC__)
__device__ __forceinline__ void pdl_launch_dependents() {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

inline bool slb_pdl_enabled() {
    static const bool on = [] {
        const char* e = getenv("SLB200_PDL");
        return e ? (atoi(e) != 0) : false;
    }();
    return on;
}

// kernel<<<grid, block, smem, st>>>(args...) as a programmatic dependent of the previous kernel in `st`
template <typename... KArgs, typename... Args>
inline cudaError_t slb_launch_dependent(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                                        cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = slb_pdl_enabled() ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#endif

#define SLB_LAUNCH_CHECK()                                                       \
    do {                                                                         \
        slb_count_launch();                                                      \
        SLB_CUDA(cudaGetLastError());                                            \
    } while (0)

int slb_validate_function(const slb_function* f, const char* what, int expect_in /* <=0: any */);
int slb_validate_gp(const slb_gp_stack* gp);
int slb_validate_grid(const slb_grid* g, bool need_points);

// ----------------------------------------------------------------------------- device side
#define SLB_DEV __device__ __forceinline__

SLB_DEV double f64mul(double a, double b) { return __dmul_rn(a, b); }
SLB_DEV double f64add(double a, double b) { return __dadd_rn(a, b); }
SLB_DEV double f64sub(double a, double b) { return __dsub_rn(a, b); }

// Order-preserving map double -> uint64 (ascending).  -0.0 is canonicalised to +0.0 first
// so that it ties with +0.0 like np.argsort sees it (lyapunov.py:512).
SLB_DEV uint64_t value_key(double v) {
    if (v == 0.0) v = 0.0;
    uint64_t b = (uint64_t)__double_as_longlong(v);
    return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}
SLB_DEV double key_value(uint64_t k) {
    uint64_t b = (k & 0x8000000000000000ull) ? (k & 0x7fffffffffffffffull) : ~k;
    return __longlong_as_double((long long)b);
}

// exp(x) for x <= 0, <= 1 ulp (checked against glibc on 2e7 points, tools/exp_neg_check.c):
// Cody-Waite reduction x = k ln2 + r, |r| <= ln2/2, degree-13 Taylor polynomial, 2^k by exponent
// add.  Branch-free so four evaluations interleave in the k-row generation loop; anything
// below exp(-700) flushes to 0 (it only ever multiplies finite L^-1 entries).
SLB_DEV double exp_neg(double x) {
    const double MAGIC = 6755399441055744.0;             // 1.5 * 2^52
    const double t = fma(x, 1.4426950408889634074, MAGIC);
    const int k = __double2loint(t);
    const double kd = t - MAGIC;
    double r = fma(kd, -6.93147180369123816490e-01, x);
    r = fma(kd, -1.90821492927058770002e-10, r);
    double p = 1.0 / 6227020800.0;
    p = fma(p, r, 1.0 / 479001600.0);
    p = fma(p, r, 1.0 / 39916800.0);
    p = fma(p, r, 1.0 / 3628800.0);
    p = fma(p, r, 1.0 / 362880.0);
    p = fma(p, r, 1.0 / 40320.0);
    p = fma(p, r, 1.0 / 5040.0);
    p = fma(p, r, 1.0 / 720.0);
    p = fma(p, r, 1.0 / 120.0);
    p = fma(p, r, 1.0 / 24.0);
    p = fma(p, r, 1.0 / 6.0);
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    p = __hiloint2double(__double2hiint(p) + (k << 20), __double2loint(p));
    return x < -700.0 ? 0.0 : p;
}

// Table-driven exp(x) for x <= 0 (<= 1 ulp against glibc on 2e7 points, tools/exp_neg_check.c):
// x = (64 k + j) ln2/64 + r, |r| <= ln2/128;  exp(x) = 2^k * T[j] * (1 + r + ... + r^5/120).
// 10 fp64 operations instead of 17; T (64 correctly rounded doubles) is read from a shared-memory
// copy because the index differs per lane (constant memory would serialise).
__constant__ double c_exp2_tab[64] = {
    1.0, 1.0108892860517005, 1.0218971486541166, 1.0330248790212284,
    1.0442737824274138, 1.0556451783605572, 1.0671404006768237, 1.0787607977571199,
    1.0905077326652577, 1.102382583307841, 1.1143867425958924, 1.1265216186082418,
    1.1387886347566916, 1.1511892299529827, 1.1637248587775775, 1.1763969916502812,
    1.189207115002721, 1.202156731452703, 1.215247359980469, 1.22848053610687,
    1.241857812073484, 1.255380757024691, 1.2690509571917332, 1.2828700160787783,
    1.2968395546510096, 1.3109612115247644, 1.3252366431597413, 1.339667524053303,
    1.3542555469368927, 1.3690024229745905, 1.383909881963832, 1.3989796725383112,
    1.4142135623730951, 1.42961333839197, 1.4451808069770467, 1.460917794180647,
    1.4768261459394993, 1.4929077282912648, 1.5091644275934228, 1.5255981507445384,
    1.5422108254079407, 1.559004400237837, 1.5759808451078865, 1.593142151342267,
    1.6104903319492543, 1.6280274218573478, 1.645755478153965, 1.6636765803267364,
    1.681792830507429, 1.7001063537185235, 1.718619298122478, 1.7373338352737062,
    1.7562521603732995, 1.7753764925265212, 1.7947090750031072, 1.8142521755003989,
    1.8340080864093424, 1.8539791250833855, 1.8741676341103, 1.8945759815869656,
    1.9152065613971474, 1.9360617934922943, 1.9571441241754002, 1.978456026387951
};

SLB_DEV void load_exp_table(double* tab_smem) {
    for (int i = threadIdx.x; i < 64; i += blockDim.x) tab_smem[i] = c_exp2_tab[i];
}

SLB_DEV double exp_neg_tab(double x, const double* __restrict__ tab) {
    const double MAGIC = 6755399441055744.0;             // 1.5 * 2^52
    const double t = fma(x, 92.33248261689366, MAGIC);                  // 64 / ln2
    const int n = __double2loint(t);
    const double nd = t - MAGIC;
    double r = fma(nd, -0.01083042469326756, x);                          // ln2/64, high part (32 bits)
    r = fma(nd, -2.9815858269852933e-12, r);                                 // ln2/64, low part
    double p = 1.0 / 120.0;
    p = fma(p, r, 1.0 / 24.0);
    p = fma(p, r, 1.0 / 6.0);
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = p * r;                                           // e^r - 1
    const double T = tab[n & 63];
    const double v = fma(T, p, T);
    const double s = __hiloint2double(__double2hiint(v) + ((n >> 6) << 20), __double2loint(v));
    return x < -700.0 ? 0.0 : s;
}

// ---- covariance expressions (slb_kernel): sum over terms of products of gpflow primitives ------
// cross form k(z, x) against a training row (kern.K(X, Xnew), functions.py:438)
template <int DIN>
SLB_DEV double kernel_expr_cross(const slb_kernel& K, const double* z, const double* x,
                                 const double* exptab) {
    double total = 0.0, term = 1.0;
    int cur = 0;
    for (int i = 0; i < K.num_prims; ++i) {
        const slb_kernel_prim& P = K.prims[i];
        if (P.term != cur) { total += term; term = 1.0; cur = P.term; }
        double v;
        if (P.kind == SLB_K_LINEAR) {
            v = 0.0;
#pragma unroll
            for (int c = 0; c < DIN; ++c) v = fma(P.w[c] * z[c], x[c], v);
        } else if (P.kind == SLB_K_CONSTANT) {
            v = P.variance;
        } else if (P.kind == SLB_K_WHITE) {
            v = 0.0;
        } else {
            double r2 = 0.0;
#pragma unroll
            for (int c = 0; c < DIN; ++c) {
                const double df = (z[c] - x[c]) * P.w[c];
                r2 = fma(df, df, r2);
            }
            if (P.kind == SLB_K_RBF) {
                v = P.variance * exp_neg_tab(-0.5 * r2, exptab);
            } else {
                const double r = sqrt(r2 + 1e-12);
                if (P.kind == SLB_K_MATERN12) {
                    v = P.variance * exp_neg_tab(-r, exptab);
                } else if (P.kind == SLB_K_MATERN32) {
                    const double sr = 1.7320508075688772 * r;
                    v = P.variance * (1.0 + sr) * exp_neg_tab(-sr, exptab);
                } else {
                    const double sr = 2.23606797749979 * r;
                    v = P.variance * (1.0 + sr + (5.0 / 3.0) * (r * r)) * exp_neg_tab(-sr, exptab);
                }
            }
        }
        term *= v;
    }
    return K.num_prims > 0 ? total + term : 0.0;
}

// U training rows at once: the primitive loop is outermost so its parameters are fetched once per
// batch and the U exp / sqrt chains of a primitive are independent (the one-row form above runs
// one dependent chain per primitive; measured 6x the RBF generation cost against ~2x here).
template <int DIN, int U>
SLB_DEV void kernel_expr_cross_n(const slb_kernel& K, const double* z, const double* const (&x)[U],
                                 const double* exptab, double (&out)[U]) {
    double total[U], term[U];
#pragma unroll
    for (int u = 0; u < U; ++u) { total[u] = 0.0; term[u] = 1.0; }
    int cur = 0;
    for (int i = 0; i < K.num_prims; ++i) {
        const slb_kernel_prim& P = K.prims[i];
        const int kind = P.kind;
        const double var = P.variance;
        if (P.term != cur) {
#pragma unroll
            for (int u = 0; u < U; ++u) { total[u] += term[u]; term[u] = 1.0; }
            cur = P.term;
        }
        double w[DIN];
#pragma unroll
        for (int c = 0; c < DIN; ++c) w[c] = P.w[c];
        double v[U];
        if (kind == SLB_K_LINEAR) {
#pragma unroll
            for (int c = 0; c < DIN; ++c) w[c] *= z[c];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                double a = 0.0;
#pragma unroll
                for (int c = 0; c < DIN; ++c) a = fma(w[c], x[u][c], a);
                v[u] = a;
            }
        } else if (kind == SLB_K_CONSTANT || kind == SLB_K_WHITE) {
#pragma unroll
            for (int u = 0; u < U; ++u) v[u] = kind == SLB_K_CONSTANT ? var : 0.0;
        } else {
            double r2[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                double a = 0.0;
#pragma unroll
                for (int c = 0; c < DIN; ++c) {
                    const double df = (z[c] - x[u][c]) * w[c];
                    a = fma(df, df, a);
                }
                r2[u] = a;
            }
            if (kind == SLB_K_RBF) {
#pragma unroll
                for (int u = 0; u < U; ++u) v[u] = var * exp_neg_tab(-0.5 * r2[u], exptab);
            } else {
                // s = c r with c = 1, sqrt(3), sqrt(5); polynomial 1, 1 + s, 1 + s + s^2 / 3
                const double cs = kind == SLB_K_MATERN12 ? 1.0
                                : kind == SLB_K_MATERN32 ? 1.7320508075688772 : 2.23606797749979;
                const double c1 = kind == SLB_K_MATERN12 ? 0.0 : 1.0;
                const double c2 = kind == SLB_K_MATERN52 ? 1.0 / 3.0 : 0.0;
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const double sr = cs * sqrt(r2[u] + 1e-12);
                    v[u] = var * fma(fma(c2, sr, c1), sr, 1.0) * exp_neg_tab(-sr, exptab);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) term[u] *= v[u];
    }
#pragma unroll
    for (int u = 0; u < U; ++u) out[u] = K.num_prims > 0 ? total[u] + term[u] : 0.0;
}

// diagonal form k(z, z) (kern.Kdiag(Xnew), functions.py:450)
template <int DIN>
SLB_DEV double kernel_expr_diag(const slb_kernel& K, const double* z) {
    double total = 0.0, term = 1.0;
    int cur = 0;
    for (int i = 0; i < K.num_prims; ++i) {
        const slb_kernel_prim& P = K.prims[i];
        if (P.term != cur) { total += term; term = 1.0; cur = P.term; }
        double v = P.variance;
        if (P.kind == SLB_K_LINEAR) {
            v = 0.0;
#pragma unroll
            for (int c = 0; c < DIN; ++c) v = fma(P.w[c] * z[c], z[c], v);
        }
        term *= v;
    }
    return K.num_prims > 0 ? total + term : 0.0;
}

// GridWorld.index_to_state (functions.py:714-731): ijk * unit_maxes + offset, two roundings.
SLB_DEV void grid_index_to_state(const slb_grid& g, int64_t idx, double* x) {
    if (g.nindex <= 0x7fffffffll) {        // 32-bit index arithmetic (same integers, ~5x fewer instructions)
        unsigned rest = (unsigned)idx;
#pragma unroll
        for (int c = SLB_MAX_DIM - 1; c >= 0; --c) {
            if (c < g.ndim) {
                const unsigned n = (unsigned)g.num_points[c];
                const unsigned q = rest / n;
                const unsigned i = rest - q * n;
                rest = q;
                x[c] = f64add(f64mul((double)i, g.unit_maxes[c]), g.offset[c]);
            }
        }
        return;
    }
#pragma unroll
    for (int c = SLB_MAX_DIM - 1; c >= 0; --c) {
        if (c < g.ndim) {
            const int64_t n = g.num_points[c];
            const int64_t i = idx % n;
            idx /= n;
            x[c] = f64add(f64mul((double)i, g.unit_maxes[c]), g.offset[c]);
        }
    }
}

// fmod(a, b) for a >= 0, b > 0, a / b < 2^52 -- bit-identical to fmod (whose result is exact): the
// quotient from one division (it can only come out one too large, when a / b rounds up to an
// integer), the remainder by an exact fma.  CUDA's fmod is a long software loop; the Triangulation
// lookup calls it once per dimension and dominated its cost.
SLB_DEV double fmod_exact_pos(double a, double b) {
    double q = floor(a / b);
    double r = fma(-q, b, a);
    if (r < 0.0) r = fma(-(q - 1.0), b, a);
    else if (r >= b) r = fma(-(q + 1.0), b, a);
    return r;
}

// ---- Triangulation (functions.py:1103-1158 lookup, :1473-1499 evaluation) ----------------
SLB_DEV void eval_triangulation(const slb_function& f, const double* xin, double* out) {
    const slb_grid& g = f.grid;
    const int d = g.ndim;
    const double eps = 2.220446049250313e-16;
    double unit[SLB_MAX_DIM];
    int64_t corner = 0;
    int poff = 0;
    int pattern = 0;
    bool all_clipped = true;
    for (int c = 0; c < d; ++c) {
        const double* pts = g.discrete_points + poff;
        const int n = (int)g.num_points[c];
        poff += n;
        const double xc = xin[c];
        // np.digitize(x, pts) - 1 clipped to [0, n-2]   (functions.py:771-773)
        int k = (int)floor((xc - g.offset[c]) / g.unit_maxes[c]);
        k = k < 0 ? 0 : (k > n - 1 ? n - 1 : k);
        while (k + 1 <= n - 1 && pts[k + 1] <= xc) ++k;
        while (k >= 0 && pts[k] > xc) --k;            // k = -1 when x < pts[0]
        k = k < 0 ? 0 : (k > n - 2 ? n - 2 : k);
        corner = corner * g.num_points[c] + k;        // rectangle_corner_index (:800-817)
        // _center_states(clip=True) % unit_maxes          (:691-712, :1120-1123)
        double cen = f64sub(xc, g.offset[c]);
        const double lo = 2.0 * eps;
        const double hi = f64sub(f64sub(g.upper[c], g.offset[c]), 2.0 * eps);
        if (cen < lo) cen = lo;
        else if (cen > hi) { cen = hi; pattern |= 1 << c; }
        else all_clipped = false;
        unit[c] = fmod_exact_pos(cen, g.unit_maxes[c]);
    }
    // simplex inside the unit cell: first simplex whose barycentric weights are all >= -tol
    int best = 0;
    double best_min = -1e300;
    const bool tabled = all_clipped && f.corner_simplex != nullptr;
    if (tabled) best = f.corner_simplex[pattern];
    for (int s = 0; s < f.nsimplex && !tabled; ++s) {
        const int64_t v0 = f.unit_simplices[s * (d + 1)];
        double o[SLB_MAX_DIM];
        if (g.nindex <= 0x7fffffffll) {        // 32-bit index arithmetic (same integers)
            unsigned t = (unsigned)v0;
            for (int c = d - 1; c >= 0; --c) {
                const unsigned n = (unsigned)g.num_points[c];
                const unsigned qq = t / n;
                o[c] = (double)(t - qq * n) * g.unit_maxes[c];
                t = qq;
            }
        } else {
            int64_t t = v0;
            for (int c = d - 1; c >= 0; --c) {
                o[c] = (double)(t % g.num_points[c]) * g.unit_maxes[c];
                t /= g.num_points[c];
            }
        }
        const double* H = f.hyperplanes + (size_t)s * d * d;
        double wsum = 0.0, wmin = 1e300;
        for (int c = 0; c < d; ++c) {
            double w = 0.0;
            for (int k = 0; k < d; ++k) w += (unit[k] - o[k]) * H[k * d + c];
            wsum += w;
            wmin = fmin(wmin, w);
        }
        wmin = fmin(wmin, 1.0 - wsum);
        if (wmin > best_min) { best_min = wmin; best = s; }
        if (wmin >= -1e-12) break;
    }
    // weights with the ORIGINAL (optionally projected) point   (:1479-1491)
    const int64_t* simp = f.unit_simplices + (size_t)best * (d + 1);
    const double* H = f.hyperplanes + (size_t)best * d * d;
    double origin[SLB_MAX_DIM], off[SLB_MAX_DIM];
    grid_index_to_state(g, simp[0] + corner, origin);
    for (int c = 0; c < d; ++c) {
        double xc = xin[c];
        if (f.flags & SLB_FLAG_PROJECT) xc = fmin(fmax(xc, g.offset[c]), g.upper[c]);
        off[c] = f64sub(xc, origin[c]);
    }
    double w[SLB_MAX_DIM + 1];
    for (int c = 0; c < d; ++c) {
        double acc = f64mul(off[0], H[c]);
        for (int k = 1; k < d; ++k) acc = f64add(acc, f64mul(off[k], H[k * d + c]));
        w[c + 1] = acc;
    }
    double acc = w[1];
    for (int c = 2; c <= d; ++c) acc = f64add(acc, w[c]);
    w[0] = f64sub(1.0, acc);
    if (f.flags & SLB_FLAG_GRADIENT) {
        // Triangulation.gradient (:1260-1326): weights[k][0] = -sum_c H[k][c], weights[k][1 + c] =
        // H[k][c]; d/dx_k = sum_v weights[k][v] * value[vertex v]   (one output column)
        for (int k = 0; k < d; ++k) {
            double hs = H[k * d];
            for (int c = 1; c < d; ++c) hs = f64add(hs, H[k * d + c]);
            double v = f64mul(-hs, f.matrix[simp[0] + corner]);
            for (int c = 0; c < d; ++c)
                v = f64add(v, f64mul(H[k * d + c], f.matrix[simp[c + 1] + corner]));
            out[k] = v;
        }
        return;
    }
    // gather vertex values and combine  (:1494-1499)
    const int od = f.out_dim;
    for (int o = 0; o < od; ++o) {
        double v = f64mul(w[0], f.matrix[(simp[0] + corner) * od + o]);
        for (int k = 1; k <= d; ++k)
            v = f64add(v, f64mul(w[k], f.matrix[(simp[k] + corner) * od + o]));
        out[o] = v;
    }
}

// ---- plants (examples/utilities.py:242-289 pendulum, :387-437 cart-pole) -----------------
// cparams layout is written by safe_learning_b200/functions.py (InvertedPendulum/CartPole).
SLB_DEV void eval_pendulum(const slb_function& f, const double* in, double* out) {
    const double* p = f.cparams;
    const double g_l = p[0], inertia = p[1], fric_i = p[2], dt = p[3];
    const bool has_norm = p[9] != 0.0, has_fric = p[10] != 0.0;
    double th = in[0], om = in[1], u = in[2];
    if (has_norm) { th = f64mul(th, p[4]); om = f64mul(om, p[5]); u = f64mul(u, p[6]); }
    const double ui = u / inertia;
    for (int i = 0; i < 10; ++i) {
        double acc = f64add(f64mul(g_l, sin(th)), ui);
        if (has_fric) acc = f64sub(acc, f64mul(fric_i, om));
        const double th_n = f64add(th, f64mul(dt, om));
        const double om_n = f64add(om, f64mul(dt, acc));
        th = th_n; om = om_n;
    }
    if (has_norm) { th = f64mul(th, p[7]); om = f64mul(om, p[8]); }
    out[0] = th; out[1] = om;
}

SLB_DEV void eval_cartpole(const slb_function& f, const double* in, double* out) {
    const double* p = f.cparams;
    const double m = p[0], M = p[1], L = p[2], b = p[3], g = p[4], dt = p[5];
    const bool has_norm = p[15] != 0.0;
    double s[4] = {in[0], in[1], in[2], in[3]};
    double u = in[4];
    if (has_norm) { for (int c = 0; c < 4; ++c) s[c] = f64mul(s[c], p[6 + c]); u = f64mul(u, p[10]); }
    for (int i = 0; i < 10; ++i) {
        const double th = s[1], v = s[2], om = s[3];
        const double st = sin(th), ct = cos(th), s2t = sin(2.0 * th);
        const double det = L * (M + m * (st * st));
        const double v_dot = (u - m * L * (om * om) * st - b * om * ct + 0.5 * m * g * L * s2t) * L / det;
        const double om_dot = (u * ct - 0.5 * m * L * (om * om) * s2t - b * (m + M) * om / (m * L)
                               + (m + M) * g * st) / det;
        s[0] = f64add(s[0], f64mul(dt, v));
        s[1] = f64add(s[1], f64mul(dt, om));
        s[2] = f64add(s[2], f64mul(dt, v_dot));
        s[3] = f64add(s[3], f64mul(dt, om_dot));
    }
    if (has_norm) for (int c = 0; c < 4; ++c) s[c] = f64mul(s[c], p[11 + c]);
    for (int c = 0; c < 4; ++c) out[c] = s[c];
}

// ---- LyapunovNetwork (examples/utilities.py:85-104): net <- act(net . kernel_i^T), V = |net|^2.
// cparams: [0] number of layers, [1 + i] output width of layer i, [9 + i] activation of layer i
// (0 tanh, 1 relu, 2 identity); `matrix` holds the layer kernels [out_i, in_i] back to back
// (kernel_i = [W^T W + eps I; W_extra], built on the host).  Widths up to SLB_NN_MAX_WIDTH.
#define SLB_NN_MAX_WIDTH 64
SLB_DEV void eval_lyapunov_nn(const slb_function& f, const double* in, double* out) {
    double h[SLB_NN_MAX_WIDTH], g[SLB_NN_MAX_WIDTH];
    int width = f.in_dim;
    for (int k = 0; k < width; ++k) h[k] = in[k];
    const double* K = f.matrix;
    const int layers = (int)f.cparams[0];
    for (int l = 0; l < layers; ++l) {
        const int od = (int)f.cparams[1 + l];
        const int act = (int)f.cparams[9 + l];
        for (int o = 0; o < od; ++o) {
            const double* row = K + (size_t)o * width;
            double acc = f64mul(h[0], row[0]);
            for (int k = 1; k < width; ++k) acc = f64add(acc, f64mul(h[k], row[k]));
            g[o] = act == 0 ? tanh(acc) : (act == 1 ? fmax(acc, 0.0) : acc);
        }
        K += (size_t)od * width;
        width = od;
        for (int k = 0; k < width; ++k) h[k] = g[k];
    }
    double v = f64mul(h[0], h[0]);
    for (int k = 1; k < width; ++k) v = f64add(v, f64mul(h[k], h[k]));
    out[0] = v;
}

// ---- NeuralNetwork inference (functions.py:1702-1729): dense layers, bias in the hidden layers
// only, output layer without bias, result scaled by output_scale.  matrix = per layer the weight
// [out_i, in_i] (already transposed to row-per-output) followed, for hidden layers with bias, by
// the bias [out_i].
SLB_DEV void eval_mlp(const slb_function& f, const double* in, double* out) {
    double h[SLB_NN_MAX_WIDTH], g[SLB_NN_MAX_WIDTH];
    int width = f.in_dim;
    for (int k = 0; k < width; ++k) h[k] = in[k];
    const double* P = f.matrix;
    const int layers = (int)f.cparams[0];
    const bool use_bias = f.cparams[18] != 0.0;
    for (int l = 0; l < layers; ++l) {
        const int od = (int)f.cparams[1 + l];
        const int act = (int)f.cparams[9 + l];
        const bool bias = use_bias && (l + 1 < layers);
        const double* b = P + (size_t)od * width;
        for (int o = 0; o < od; ++o) {
            const double* row = P + (size_t)o * width;
            double acc = f64mul(h[0], row[0]);
            for (int k = 1; k < width; ++k) acc = f64add(acc, f64mul(h[k], row[k]));
            if (bias) acc = f64add(acc, b[o]);
            g[o] = act == 0 ? tanh(acc) : (act == 1 ? fmax(acc, 0.0) : acc);
        }
        P += (size_t)od * width + (bias ? od : 0);
        width = od;
        for (int k = 0; k < width; ++k) h[k] = g[k];
    }
    for (int k = 0; k < width; ++k) out[k] = f64mul(h[k], f.cparams[17]);
}

// Evaluate a fused function object. `in` has f.in_dim entries, `out` receives the result
// columns; returns the number of columns (1 after NORM1).  In gp_sweep.cu (SLB_EVAL_NOINLINE) it
// is deliberately NOT inlined: the tile kernel calls it five times per point (policy, V twice,
// L_V twice) in cold prologue / epilogue code, and one shared copy keeps the kernel text small
// (27 k -> ~10 k instructions: -1.3% back-to-back, -3.3% after an L2 flush).  The thread-per-point
// kernels of light.cu inline it (the call overhead costs them ~20%).
#ifdef SLB_EVAL_NOINLINE
#define SLB_EVAL_ATTR static __device__ __noinline__
#else
#define SLB_EVAL_ATTR static __device__ __forceinline__
#endif
SLB_EVAL_ATTR int eval_fn(const slb_function& f, const double* in, double* out) {
    int od = f.out_dim;
    switch (f.kind) {
    case SLB_FN_CONSTANT:
        for (int o = 0; o < od; ++o) out[o] = f.cparams[o];
        break;
    case SLB_FN_LINEAR:       // functions.py:1583   y_o = sum_k x_k A[o,k]
        for (int o = 0; o < od; ++o) {
            const double* row = f.matrix + o * f.in_dim;
            double acc = f64mul(in[0], row[0]);
            for (int k = 1; k < f.in_dim; ++k) acc = f64add(acc, f64mul(in[k], row[k]));
            out[o] = acc;
        }
        break;
    case SLB_FN_QUADRATIC: {  // functions.py:1537-1539   sum_c (sum_r x_r P[r,c]) * x_c
        const int n = f.in_dim;
        double total = 0.0;
        for (int c = 0; c < n; ++c) {
            double lin = f64mul(in[0], f.matrix[c]);
            for (int r = 1; r < n; ++r) lin = f64add(lin, f64mul(in[r], f.matrix[r * n + c]));
            const double prod = f64mul(lin, in[c]);
            total = (c == 0) ? prod : f64add(total, prod);
        }
        out[0] = total;
        od = 1;
        break;
    }
    case SLB_FN_TRIANGULATION:
        eval_triangulation(f, in, out);
        break;
    case SLB_FN_PENDULUM:
        eval_pendulum(f, in, out); od = 2;
        break;
    case SLB_FN_CARTPOLE:
        eval_cartpole(f, in, out); od = 4;
        break;
    case SLB_FN_LYAPUNOV_NN:
        eval_lyapunov_nn(f, in, out); od = 1;
        break;
    case SLB_FN_MLP:
        eval_mlp(f, in, out);
        break;
    default:
        for (int o = 0; o < od; ++o) out[o] = __longlong_as_double(0x7ff8000000000000ll);
        break;
    }
    if (f.flags & SLB_FLAG_SATURATE)
        for (int o = 0; o < od; ++o) out[o] = fmin(fmax(out[o], f.lower), f.upper);
    if (f.flags & (SLB_FLAG_ABS | SLB_FLAG_NORM1))
        for (int o = 0; o < od; ++o) out[o] = fabs(out[o]);
    if (f.flags & SLB_FLAG_NORM1) {
        double acc = out[0];
        for (int o = 1; o < od; ++o) acc = f64add(acc, out[o]);
        out[0] = acc;
        od = 1;
    }
    if (f.flags & SLB_FLAG_MAXABS) {
        double acc = fabs(out[0]);
        for (int o = 1; o < od; ++o) acc = fmax(acc, fabs(out[o]));
        out[0] = acc;
        od = 1;
    }
    if (f.flags & SLB_FLAG_SCALE)
        for (int o = 0; o < od; ++o) out[o] = f64mul(out[o], f.out_scale);
    return od;
}

// Fast path in front of eval_fn for the small linear-algebra objects that make up the per-point
// prologue / epilogue of every sweep of the reference's experiments (policy = Saturation(LinearSystem),
// V = QuadraticFunction, L_V = abs(LinearSystem)): in_dim <= 4, out_dim <= 2, operands in registers,
// eval_fn's arithmetic operation for operation (bit-identical).  The generic interpreter costs a
// call, local-memory operand arrays and runtime-bounded loops per evaluation: five of them made up
// most of the 24 us floor of the filter's mean stage (M = 0: tools/mean_floor_probe.py).
SLB_DEV int eval_fn_small(const slb_function& f, const double* in, double* out) {
    const int n = f.in_dim;
    if (f.kind == SLB_FN_LINEAR && n <= 4 && f.out_dim <= 2 &&
        !(f.flags & (SLB_FLAG_MAXABS | SLB_FLAG_GRADIENT))) {
        int od = f.out_dim;
        double x[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) x[k] = k < n ? in[k] : 0.0;
        double y[2];
#pragma unroll
        for (int o = 0; o < 2; ++o) {
            if (o < od) {
                const double* row = f.matrix + o * n;
                double acc = f64mul(x[0], __ldg(row));
#pragma unroll
                for (int k = 1; k < 4; ++k)
                    if (k < n) acc = f64add(acc, f64mul(x[k], __ldg(row + k)));
                y[o] = acc;
            } else {
                y[o] = 0.0;
            }
        }
        if (f.flags & SLB_FLAG_SATURATE) {
            y[0] = fmin(fmax(y[0], f.lower), f.upper);
            y[1] = fmin(fmax(y[1], f.lower), f.upper);
        }
        if (f.flags & (SLB_FLAG_ABS | SLB_FLAG_NORM1)) { y[0] = fabs(y[0]); y[1] = fabs(y[1]); }
        if (f.flags & SLB_FLAG_NORM1) {
            if (od == 2) y[0] = f64add(y[0], y[1]);
            od = 1;
        }
        if (f.flags & SLB_FLAG_SCALE) { y[0] = f64mul(y[0], f.out_scale); y[1] = f64mul(y[1], f.out_scale); }
        out[0] = y[0];
        if (od == 2) out[1] = y[1];
        return od;
    }
    if (f.kind == SLB_FN_QUADRATIC && n <= 4 && !(f.flags & ~SLB_FLAG_SCALE)) {
        double x[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) x[k] = k < n ? in[k] : 0.0;
        double total = 0.0;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            if (c < n) {
                double lin = f64mul(x[0], __ldg(f.matrix + c));
#pragma unroll
                for (int r = 1; r < 4; ++r)
                    if (r < n) lin = f64add(lin, f64mul(x[r], __ldg(f.matrix + r * n + c)));
                const double prod = f64mul(lin, x[c]);
                total = (c == 0) ? prod : f64add(total, prod);
            }
        }
        if (f.flags & SLB_FLAG_SCALE) total = f64mul(total, f.out_scale);
        out[0] = total;
        return 1;
    }
    return eval_fn(f, in, out);
}

// The per-point prologue / epilogue of a sweep kernel reads a handful of small operands through
// pointers of the descriptor (policy / V / L_V / L_f matrices, prior-mean rows), one dependent global
// load after the other; after an L2 flush (or any eviction) each is an HBM round trip in every CTA's
// serial chain.  Touch them all at once when the kernel starts.
SLB_DEV void prefetch_descriptor_operands(const slb_sweep& cfg) {
    const int i = threadIdx.x;
    const void* p = nullptr;
    if (i == 0) p = cfg.policy.matrix;
    else if (i == 1) p = cfg.lyapunov.matrix;
    else if (i == 2) p = cfg.lipschitz_v.matrix;
    else if (i == 3) p = cfg.lipschitz_f.matrix;
    else if (i < 4 + cfg.gp.num_outputs && i < 4 + SLB_MAX_OUT) p = cfg.gp.outputs[i - 4].prior_mean;
    if (p != nullptr) asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
}

// The Lyapunov decision for one state (lyapunov.py:265-288, 324-376, 441).
//   x [d]; mu [d] predicted mean; err [d] error bounds (beta*sigma) or nullptr when the
//   dynamics are deterministic.  Returns negative = decrease < threshold (false on NaN).
struct slb_decision { double vx, decrease, threshold; bool negative; };

// The decision of lyapunov.py:436-441 in three independent pieces (the tile kernel evaluates the
// x-only piece while the GP runs, and the two mean-dependent pieces on different warps):
// (1) V(x) and threshold(x) = -|L_V(x)|_1 (1 + L_f) tau                      :284-288
// `flat_index` is the point's flat grid index (for cfg.lf_values), or -1 for an explicit state.
SLB_DEV void lyapunov_state_terms(const slb_sweep& cfg, const double* x, int64_t flat_index,
                                  double* vx_out, double* threshold_out) {
    double tmp[SLB_MAX_OUT], vx[1];
    eval_fn_small(cfg.lyapunov, x, vx);
    *vx_out = vx[0];
    double lvx;
    if (cfg.lipschitz_v.kind != SLB_FN_NONE) {               // :284-286 (1-norm of a vector lv)
        const int nl = eval_fn_small(cfg.lipschitz_v, x, tmp);
        lvx = tmp[0];
        if (nl > 1) {
            lvx = fabs(tmp[0]);
            for (int j = 1; j < nl; ++j) lvx = f64add(lvx, fabs(tmp[j]));
        }
    } else {
        lvx = cfg.lv_const;
    }
    double lf = cfg.lf_const;                                // lyapunov.py:227-244, :287
    if (cfg.lf_values != nullptr && flat_index >= 0) {
        lf = cfg.lf_values[flat_index - cfg.lf_index_base];
    } else if (cfg.lipschitz_f.kind != SLB_FN_NONE) {
        eval_fn_small(cfg.lipschitz_f, x, tmp);
        lf = tmp[0];
    }
    *threshold_out = f64mul(f64mul(-lvx, f64add(1.0, lf)), cfg.tau);   // :288
}

// (2) sum_j L_V(mu)_j err_j, L_V evaluated at the predicted MEAN                :344-347
SLB_DEV double lyapunov_error_bound(const slb_sweep& cfg, const double* mu, const double* err) {
    const int d = cfg.grid.ndim;
    double tmp[SLB_MAX_OUT];
    double bound;
    if (cfg.lipschitz_v.kind != SLB_FN_NONE) {
        const int nl = eval_fn_small(cfg.lipschitz_v, mu, tmp);
        if (nl == 1) {
            bound = f64mul(tmp[0], err[0]);
            for (int j = 1; j < d; ++j) bound = f64add(bound, f64mul(tmp[0], err[j]));
        } else {
            bound = f64mul(tmp[0], err[0]);
            for (int j = 1; j < d; ++j) bound = f64add(bound, f64mul(tmp[j], err[j]));
        }
    } else {
        bound = f64mul(cfg.lv_const, err[0]);
        for (int j = 1; j < d; ++j) bound = f64add(bound, f64mul(cfg.lv_const, err[j]));
    }
    return bound;
}

// (3) V(mu); then decrease = (V(mu) - V(x)) + bound  and  negative = decrease < threshold
SLB_DEV slb_decision lyapunov_combine(double vx, double threshold, double vm, double bound) {
    slb_decision r;
    r.vx = vx;
    r.threshold = threshold;
    r.decrease = f64add(f64sub(vm, vx), bound);                // :351-352, :376
    r.negative = r.decrease < r.threshold;                   // :441 strict, NaN -> false
    return r;
}

SLB_DEV slb_decision lyapunov_decide(const slb_sweep& cfg, const double* x, int64_t flat_index,
                                     const double* mu, const double* err) {
    double vx, threshold, vm[1];
    lyapunov_state_terms(cfg, x, flat_index, &vx, &threshold);
    eval_fn_small(cfg.lyapunov, mu, vm);
    const double bound = err != nullptr ? lyapunov_error_bound(cfg, mu, err) : 0.0;
    return lyapunov_combine(vx, threshold, vm[0], bound);
}
