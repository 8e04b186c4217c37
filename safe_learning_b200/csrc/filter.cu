// filter.cu -- certified decision filter in front of the O(M^2) GP posterior of the Lyapunov sweep.
//
// The reference evaluates, for every grid point (lyapunov.py:436-441, functions.py:417-458, 507-515)
//     negative = V(mu) - V(x) + sum_j L_V(mu)_j beta_j sigma_j  <  -L_V(x) (1 + L_f) tau
// and all of its cost is sigma_j = sqrt(k** - |L^-1 k|^2): M^2 flops per point and Cholesky factor,
// against M for the mean mu = k . (L^-T alpha).  But the comparison is monotone in every sigma_j, and
//     0 <= sigma_j <= sigma_j given ANY subset of the training set <= prior sigma_j,
// where the posterior given the first R rows of the training set is just the first R rows of the
// same triangular solve (L is lower triangular): sum_{i<R} a_i^2 <= sum_{i<M} a_i^2.  So:
//   stage 1  (filter_mean_kernel) exact mean (all M kernel values, one exp each -- the cost of a
//            Bellman sweep), V(mu), L_V(mu); decide every point whose outcome is the same for
//            sigma = 0 and the prior sigma; the rest is compacted into list A with its terms;
//   stage 2  (filter_head_kernel, over list A) the head rows a_i = sum_j L^-1[i,j] k_j,
//            i < R = SLB_HEAD_RANK, thread per point in registers; decide with the tighter bound;
//   rest     compacted into list B for the full fp64 posterior (gp_tile_kernel, gp_sweep.cu).
// A point is only decided when the outcome holds with a guard band of 1e-6 relative to the
// magnitudes involved (five orders above the rounding differences between two fp64 evaluation
// orders of the posterior; the GP tolerance of the parity contract is 1e-5), anything with a NaN
// goes to the full path, so the flags equal those of slb_lyapunov_sweep bit for bit.
// On the C2 workload (256 x 256, M = 500) 90.7% of the points are decided by the mean alone, 8%
// by the head-rank bound and 1.3% are refined (tools/filter_probe.py reproduces this on the CPU).
#define SLB_EVAL_NOINLINE 1
#include "common.cuh"
#include "gp_mean.cuh"

namespace {

constexpr int FT = 64;                 // threads per CTA = points per CTA (1024 CTAs at 256 x 256,
                                       // 8 resident per SM: single wave, 98.8% balanced)
constexpr int HR = SLB_HEAD_RANK;
constexpr int HB = 16;                 // column block of the head-row update (static row ranges)
constexpr int HH = HR / 2;             // rows per pass of the head solve (accumulators in registers)
constexpr int64_t CHUNK = 1 << 22;     // points per pass of the three stages (bounds the workspace)

// terms of one undecided point, carried from stage 1 to stage 2
struct filter_side { double dec0, thr, guard, coef[SLB_MAX_OUT]; };

struct filter_args {
    const double* points;              // explicit states [n, d] or nullptr (grid index range)
    int64_t n;
    int64_t idx_begin;
    uint8_t* negative;
    double* values;
    int64_t* list_a;                   // undecided after stage 1 (index relative to the range)
    filter_side* side_a;               // their terms, same order
    int64_t* list_b;                   // undecided after stage 2 -> full posterior
    unsigned long long* counts;        // [0] entries of list_a, [1] entries of list_b
    unsigned long long* stats;         // nullptr or [4], see slb200.h
};

// outcome for err_j = beta_j sigma_j with sigma_j in [0, shi_j]:  +1 decided negative (True),
// 0 decided not negative (False), -1 undecided.  term_j = L_V(mu)_j beta_j sigma_j lies between 0 and
// coef_j shi_j; anything non-finite stays undecided (NaN compares false on both sides).
SLB_DEV int decide(const filter_side& t, const double* shi, int d) {
    double ub = 0.0, lb = 0.0;
    for (int j = 0; j < d; ++j) {
        const double e = t.coef[j] * shi[j];
        ub += fmax(e, 0.0);
        lb += fmin(e, 0.0);
        if (!(e == e)) { ub = e; lb = e; break; }        // NaN: poison both sums
    }
    const double slack = t.guard + 1e-6 * (fabs(ub) + fabs(lb));
    if (t.dec0 + ub + slack < t.thr) return 1;
    if (t.dec0 + lb - slack >= t.thr) return 0;
    return -1;
}

// warp-aggregated append of the lanes with `take` to a device list; returns the slot (or -1)
SLB_DEV long long list_append(bool take, unsigned long long* counter) {
    const unsigned ballot = __ballot_sync(0xffffffffu, take);
    if (ballot == 0) return -1;
    const int lane = threadIdx.x & 31;
    unsigned long long base = 0;
    if (lane == __ffs(ballot) - 1) base = atomicAdd(counter, (unsigned long long)__popc(ballot));
    base = __shfl_sync(0xffffffffu, base, __ffs(ballot) - 1);
    return take ? (long long)(base + __popc(ballot & ((1u << lane) - 1))) : -1;
}

SLB_DEV void count_stat(bool hit, unsigned long long* slot) {
    const unsigned ballot = __ballot_sync(0xffffffffu, hit);
    if (ballot != 0 && (threadIdx.x & 31) == 0) atomicAdd(slot, (unsigned long long)__popc(ballot));
}

SLB_DEV void load_state(const slb_sweep& cfg, const filter_args& a, int64_t rel, double* z) {
    const int d = cfg.grid.ndim;
    if (a.points != nullptr) {
        for (int c = 0; c < d; ++c) z[c] = a.points[rel * d + c];
    } else {
        grid_index_to_state(cfg.grid, a.idx_begin + rel, z);
    }
}

// ---- stage 1: exact mean, prior bound ---------------------------------------------------------
template <int DIN>
__global__ void __launch_bounds__(FT, 8)
filter_mean_kernel(const __grid_constant__ slb_sweep cfg, const filter_args a) {
    __shared__ double exptab[64];
    __shared__ double stage[BCHUNK * (DIN + SLB_MAX_OUT)];
    load_exp_table(exptab);
    __syncthreads();
    const int64_t rel0 = (int64_t)blockIdx.x * FT + threadIdx.x;
    const bool valid = rel0 < a.n;
    const int64_t rel = valid ? rel0 : a.n - 1;   // every thread stays for the block barriers
    const int d = cfg.grid.ndim;
    const int D = cfg.gp.num_outputs;

    // ---- x, V(x), threshold(x), u = policy(x)           (lyapunov.py:436, 284-288)
    double z[SLB_MAX_IN];
    load_state(cfg, a, rel, z);
    filter_side t;
    double vx;
    lyapunov_state_terms(cfg, z, a.points != nullptr ? -1 : a.idx_begin + rel, &vx, &t.thr);
    {
        double u[SLB_MAX_OUT];
        const int m = eval_fn(cfg.policy, z, u);
        for (int c = 0; c < m; ++c) z[d + c] = u[c];
    }

    // ---- exact posterior mean of every output (functions.py:439-442 as k . L^-T alpha)
    double mu[SLB_MAX_OUT];
    gp_mean_only<DIN>(cfg.gp, z, mu, exptab, stage);

    // ---- V(mu), L_V(mu) and the coefficient of every sigma_j          (lyapunov.py:344-352)
    double vm[1];
    eval_fn(cfg.lyapunov, mu, vm);
    t.dec0 = f64sub(vm[0], vx);
    double lvmu = 0.0;                      // sum_j |L_V(mu)_j mu_j|: scale of V's sensitivity to mu
    {
        double lv[SLB_MAX_OUT];
        int nl = 1;
        if (cfg.lipschitz_v.kind != SLB_FN_NONE) nl = eval_fn(cfg.lipschitz_v, mu, lv);
        else lv[0] = cfg.lv_const;
        for (int j = 0; j < SLB_MAX_OUT; ++j) {
            const double l = j < D ? (nl == 1 ? lv[0] : lv[j]) : 0.0;
            t.coef[j] = j < D ? l * cfg.gp.outputs[j].beta : 0.0;
            if (j < D) lvmu += fabs(l * mu[j]);
        }
    }
    t.guard = 1e-6 * (fabs(vm[0]) + fabs(vx) + fabs(t.thr) + lvmu) + 1e-300;

    // ---- sigma_j <= prior sigma_j     (functions.py:450 without data)
    double shi[SLB_MAX_OUT];
    for (int j = 0; j < D; ++j) {
        const slb_gp_factor& F = cfg.gp.factors[cfg.gp.outputs[j].factor];
        shi[j] = sqrt(F.kernel.num_prims > 0 ? kernel_expr_diag<DIN>(F.kernel, z) : F.variance);
    }
    const int outcome = decide(t, shi, D);
    const bool undecided = valid && outcome < 0;
    if (valid) {
        a.negative[rel] = outcome > 0 ? 1 : 0;
        if (a.values != nullptr) a.values[rel] = vx;
    }
    const long long slot = list_append(undecided, a.counts + 0);
    if (undecided) { a.list_a[slot] = rel; a.side_a[slot] = t; }
    if (a.stats != nullptr) {
        count_stat(valid && !undecided, a.stats + 0);
        count_stat(valid, a.stats + 3);
    }
}

// ---- stage 2: head rows of the triangular solve, thread per point -----------------------------
// acc[i - ROW0] += W[i, j] k_j for rows ROW0 .. ROW0 + HH - 1 and the columns of block B, rows
// below the block's first column only (W is lower triangular).  W is the zero-padded column-major
// head block: column j holds rows contiguously, 2 doubles per LDG.128, same address in all lanes.
template <int DIN, int ROW0, int B>
SLB_DEV void head_cols(double (&acc)[HH], const slb_gp_factor& F, bool general, const double* zs,
                       const double* xhead, int rows, double s2, const double* exptab) {
    constexpr int I0 = (HB * B > ROW0) ? HB * B : ROW0;      // first row this block touches
    const double* __restrict__ Wt = F.Whead;
#pragma unroll 1
    for (int jj = 0; jj < HB; jj += 4) {
        const int j0 = HB * B + jj;
        if (j0 >= rows) break;
        double kv[4];
        if (general) {
            const double* xr[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) xr[u] = xhead + min(j0 + u, rows - 1) * DIN;
            kernel_expr_cross_n<DIN, 4>(F.kernel, zs, xr, exptab, kv);
        } else {
            double t2[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const double* xr = xhead + min(j0 + u, rows - 1) * DIN;
                double a2 = 0.0;
#pragma unroll
                for (int c = 0; c < DIN; ++c) { const double df = zs[c] - xr[c]; a2 = fma(df, df, a2); }
                t2[u] = a2;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) kv[u] = F.variance * exp_neg_tab(-0.5 * t2[u], exptab);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const double k = j0 + u < rows ? s2 * kv[u] : 0.0;       // functions.py:438 (scale^2 K)
            const double2* col = reinterpret_cast<const double2*>(Wt + (size_t)(j0 + u) * HR);
#pragma unroll
            for (int i = I0; i < ROW0 + HH; i += 2) {
                const double2 w = __ldg(col + (i >> 1));
                acc[i - ROW0] = fma(w.x, k, acc[i - ROW0]);
                acc[i + 1 - ROW0] = fma(w.y, k, acc[i + 1 - ROW0]);
            }
        }
    }
}

// upper bound of the latent variance of factor F at z from its first min(M, HR) training rows:
// (scale^2 k** - sum_{i < R} a_i^2) / scale^2   (functions.py:450-451, 456 restricted to R rows).
// Two passes of HH rows each keep the accumulators in registers at 8 CTAs per SM.
template <int DIN>
SLB_DEV double head_variance(const slb_gp_factor& F, const double* z, const double* exptab,
                             const double* xhead) {
    const bool general = F.kernel.num_prims > 0;
    const int rows = min(F.M, HR);
    double zs[DIN];
#pragma unroll
    for (int c = 0; c < DIN; ++c) zs[c] = general ? z[c] : z[c] / F.lengthscales[c];
    const double s2 = f64mul(F.scale, F.scale);
    static_assert(HR == 4 * HB && HH == 2 * HB, "head_cols instantiations below cover HR rows");
    double ss = 0.0;
    {
        double acc[HH];
#pragma unroll
        for (int i = 0; i < HH; ++i) acc[i] = 0.0;
        head_cols<DIN, 0, 0>(acc, F, general, zs, xhead, rows, s2, exptab);
        head_cols<DIN, 0, 1>(acc, F, general, zs, xhead, rows, s2, exptab);
#pragma unroll
        for (int i = 0; i < HH; ++i) ss = fma(acc[i], acc[i], ss);
    }
    if (rows > HH) {
        double acc[HH];
#pragma unroll
        for (int i = 0; i < HH; ++i) acc[i] = 0.0;
        head_cols<DIN, HH, 0>(acc, F, general, zs, xhead, rows, s2, exptab);
        head_cols<DIN, HH, 1>(acc, F, general, zs, xhead, rows, s2, exptab);
        head_cols<DIN, HH, 2>(acc, F, general, zs, xhead, rows, s2, exptab);
        head_cols<DIN, HH, 3>(acc, F, general, zs, xhead, rows, s2, exptab);
#pragma unroll
        for (int i = 0; i < HH; ++i) ss = fma(acc[i], acc[i], ss);
    }
    double kss = F.kss;
    if (general) kss = s2 * kernel_expr_diag<DIN>(F.kernel, z);
    return f64sub(kss, ss) / s2;
}

template <int DIN>
__global__ void __launch_bounds__(FT, 8)
filter_head_kernel(const __grid_constant__ slb_sweep cfg, const filter_args a) {
    __shared__ double exptab[64];
    __shared__ double xhead[SLB_MAX_OUT][HR * DIN];
    const int64_t count = (int64_t)a.counts[0];
    const int64_t k0 = (int64_t)blockIdx.x * FT;
    if (k0 >= count) return;                      // the grid covers the worst case
    load_exp_table(exptab);
    const int nf = cfg.gp.num_factors;
    for (int f = 0; f < nf; ++f) {
        const slb_gp_factor& F = cfg.gp.factors[f];
        const int rows = min(F.M, HR);
        for (int i = threadIdx.x; i < rows * DIN; i += FT) xhead[f][i] = F.Xs[i];
    }
    __syncthreads();
    const bool valid = k0 + threadIdx.x < count;
    const int64_t k = valid ? k0 + threadIdx.x : count - 1;
    const int64_t rel = a.list_a[k];
    const filter_side t = a.side_a[k];
    const int d = cfg.grid.ndim;
    const int D = cfg.gp.num_outputs;
    double z[SLB_MAX_IN];
    load_state(cfg, a, rel, z);
    {
        double u[SLB_MAX_OUT];
        const int m = eval_fn(cfg.policy, z, u);
        for (int c = 0; c < m; ++c) z[d + c] = u[c];
    }
    double shi[SLB_MAX_OUT];
    for (int j = 0; j < D; ++j) {
        const slb_gp_factor& F = cfg.gp.factors[cfg.gp.outputs[j].factor];
        shi[j] = sqrt(F.kernel.num_prims > 0 ? kernel_expr_diag<DIN>(F.kernel, z) : F.variance);
    }
    for (int f = 0; f < nf; ++f) {
        const slb_gp_factor& F = cfg.gp.factors[f];
        if (F.M == 0 || F.Whead == nullptr) continue;
        const double s = sqrt(head_variance<DIN>(F, z, exptab, xhead[f]));   // NaN if negative
        for (int j = 0; j < D; ++j)
            if (cfg.gp.outputs[j].factor == f) shi[j] = s;
    }
    const int outcome = decide(t, shi, D);
    const bool undecided = valid && outcome < 0;
    if (valid && outcome >= 0) a.negative[rel] = outcome > 0 ? 1 : 0;
    const long long slot = list_append(undecided, a.counts + 1);
    if (undecided) a.list_b[slot] = rel;
    if (a.stats != nullptr) {
        count_stat(valid && !undecided, a.stats + 1);
        count_stat(undecided, a.stats + 2);
    }
}

template <int DIN>
int launch_filter(cudaStream_t st, const slb_sweep& cfg, const filter_args& a) {
    const int64_t blocks = (a.n + FT - 1) / FT;
    filter_mean_kernel<DIN><<<(unsigned)blocks, FT, 0, st>>>(cfg, a);
    SLB_LAUNCH_CHECK();
    filter_head_kernel<DIN><<<(unsigned)blocks, FT, 0, st>>>(cfg, a);
    SLB_LAUNCH_CHECK();
    return 0;
}

}  // namespace

// gp_sweep.cu: the full posterior on the compacted list (count read on the device)
int slb_launch_refine(cudaStream_t st, const slb_sweep& cfg, int64_t n_max, int64_t idx_begin,
                      const int64_t* list, const unsigned long long* count, uint8_t* negative,
                      double* values);

extern "C" {

int64_t slb_filter_workspace(int64_t n) {
    if (n < 0) n = 0;
    if (n > CHUNK) n = CHUNK;     // longer ranges are swept in passes of CHUNK points
    // [0] |list A|, [1] |list B| (uint64, 64 bytes reserved), list A, list B, terms of list A
    return 64 + n * (int64_t)(2 * sizeof(int64_t) + sizeof(filter_side));
}

int slb_lyapunov_sweep_filtered(void* stream, const slb_sweep* cfg, int64_t idx_begin,
                                int64_t idx_end, uint8_t* negative_dev, double* values_dev,
                                void* workspace_dev, int64_t* stats_dev) {
    SLB_CHECK(cfg != nullptr, "slb_lyapunov_sweep_filtered: null config");
    SLB_CHECK(idx_begin >= 0 && idx_end >= idx_begin && idx_end <= cfg->grid.nindex,
              "slb_lyapunov_sweep_filtered: index range [%lld, %lld) outside the grid (nindex %lld)",
              (long long)idx_begin, (long long)idx_end, (long long)cfg->grid.nindex);
    const int64_t n_all = idx_end - idx_begin;
    if (n_all == 0) return 0;
    SLB_CHECK(negative_dev != nullptr && workspace_dev != nullptr,
              "slb_lyapunov_sweep_filtered: negative_dev and workspace_dev are required");
    if (slb_validate_grid(&cfg->grid, false)) return 1;
    const int d = cfg->grid.ndim;
    if (slb_validate_function(&cfg->policy, "policy", d)) return 1;
    SLB_CHECK(cfg->policy.kind != SLB_FN_NONE, "lyapunov sweep: a policy is required");
    if (slb_validate_function(&cfg->lyapunov, "lyapunov_function", d)) return 1;
    SLB_CHECK(cfg->lyapunov.kind != SLB_FN_NONE, "lyapunov sweep: a Lyapunov function is required");
    if (slb_validate_function(&cfg->lipschitz_v, "lipschitz_lyapunov", d)) return 1;
    if (slb_validate_function(&cfg->lipschitz_f, "lipschitz_dynamics", d)) return 1;
    const int m = (cfg->policy.flags & SLB_FLAG_NORM1) ? 1 : cfg->policy.out_dim;
    SLB_CHECK(m >= 1 && m <= SLB_MAX_ACT, "policy output dim %d unsupported", m);
    SLB_CHECK(cfg->gp.num_outputs > 0, "slb_lyapunov_sweep_filtered needs GP dynamics "
              "(deterministic dynamics have nothing to filter: use slb_lyapunov_sweep)");
    if (slb_validate_gp(&cfg->gp)) return 1;
    SLB_CHECK(cfg->gp.num_outputs == d, "GP stack has %d outputs but the state has %d dims",
              cfg->gp.num_outputs, d);
    SLB_CHECK(cfg->gp.input_dim == d + m, "GP input_dim %d != state %d + action %d",
              cfg->gp.input_dim, d, m);
    for (int o = 0; o < cfg->gp.num_outputs; ++o)
        SLB_CHECK(cfg->gp.outputs[o].gamma != nullptr, "filtered sweep: GP output %d has no gamma", o);
    for (int f = 0; f < cfg->gp.num_factors; ++f)
        SLB_CHECK(cfg->gp.factors[f].M == 0 || cfg->gp.factors[f].Whead != nullptr,
                  "filtered sweep: GP factor %d has no head block (Whead)", f);
    cudaStream_t st = (cudaStream_t)stream;
    char* ws = static_cast<char*>(workspace_dev);
    const int64_t cap = n_all < CHUNK ? n_all : CHUNK;
    filter_args a;
    a.points = nullptr;
    a.counts = reinterpret_cast<unsigned long long*>(ws);
    a.list_a = reinterpret_cast<int64_t*>(ws + 64);
    a.list_b = a.list_a + cap;
    a.side_a = reinterpret_cast<filter_side*>(a.list_b + cap);
    a.stats = reinterpret_cast<unsigned long long*>(stats_dev);
    for (int64_t off = 0; off < n_all; off += CHUNK) {
        const int64_t n = n_all - off < CHUNK ? n_all - off : CHUNK;
        SLB_CUDA(cudaMemsetAsync(a.counts, 0, 64, st));
        a.n = n; a.idx_begin = idx_begin + off;
        a.negative = negative_dev + off;
        a.values = values_dev ? values_dev + off : nullptr;
        int rc;
        switch (cfg->gp.input_dim) {
        case 1: rc = launch_filter<1>(st, *cfg, a); break;
        case 2: rc = launch_filter<2>(st, *cfg, a); break;
        case 3: rc = launch_filter<3>(st, *cfg, a); break;
        case 4: rc = launch_filter<4>(st, *cfg, a); break;
        case 5: rc = launch_filter<5>(st, *cfg, a); break;
        case 6: rc = launch_filter<6>(st, *cfg, a); break;
        default:
            slb_set_error("GP input_dim %d not compiled (1..6)", cfg->gp.input_dim);
            return 1;
        }
        if (rc) return rc;
        rc = slb_launch_refine(st, *cfg, n, idx_begin + off, a.list_b, a.counts + 1,
                               negative_dev + off, values_dev ? values_dev + off : nullptr);
        if (rc) return rc;
    }
    return 0;
}

}  // extern "C"
