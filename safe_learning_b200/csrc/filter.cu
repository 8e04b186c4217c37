// filter.cu -- certified decision filter in front of the O(M^2) GP posterior of the Lyapunov sweep.
//
// The reference evaluates, for every grid point (lyapunov.py:436-441, functions.py:417-458, 507-515)
//     negative = V(mu) - V(x) + sum_j L_V(mu)_j beta_j sigma_j  <  -L_V(x) (1 + L_f) tau
// and all of its cost is sigma_j = sqrt(k** - |L^-1 k|^2): M^2 flops per point and Cholesky factor,
// against M for the mean mu = k . (L^-T alpha).  But the comparison is monotone in every sigma_j, and
//     0 <= sigma_j <= sigma_j given ANY subset of the training set <= prior sigma_j.  So:
//   stage 1  (filter_mean_kernel, thread per point) the posterior mean from all M kernel values
//            (one exp each -- the cost of a Bellman sweep), V(mu), L_V(mu); decide every point
//            whose outcome is the same for sigma = 0 and the prior sigma; the rest is compacted
//            into list A with its terms;
//   stage 2  (filter_head_kernel, warp per point of list A) the posterior variance given a HEAD
//            SUBSET of at most SLB_HEAD_RANK training points (chosen by the host in pivoted-
//            Cholesky order, its own small factor: slb_gp_factor.Whead / Xhead); decide with
//            the tighter bound;
//   rest     compacted into list B for the full fp64 posterior (gp_tile_kernel, gp_sweep.cu).
// Certification.  A point is only decided when the outcome holds with a guard band that covers
// (a) 1e-6 relative to the magnitudes involved (five orders above the rounding differences between
// two fp64 evaluation orders of the posterior; the GP tolerance of the parity contract is 1e-5) and
// (b) a computed bound of the error of stage 1's mean: it is summed as k . gamma with a kernel
// value accurate to EPS_K relative (expanded squared distance, table-driven exp with a cubic
// remainder -- tools/exp_neg_fast_check.c), so |mean error| <= (EPS_K + M 2^-52) sum_j |k_j
// gamma_j| <= (...) |gamma|_1 (slb_gp_output.gamma_l1, k <= 1 after folding the variances into
// gamma), entering the comparison through L_V(mu).  Anything non-finite goes to the full path, so
// the flags equal those of slb_lyapunov_sweep bit for bit.
// The training inputs / gamma slices are block-wide shared-memory landings: they are brought in
// by TMA bulk copies (cp.async.bulk + mbarrier, bulk_copy.cuh), double buffered, issued by one
// thread while the block works on the previous slice.
#define SLB_EVAL_NOINLINE 1
#include "common.cuh"

#include <atomic>
#include <string.h>
#include <stdlib.h>

#include "gp_mean_staged.cuh"
#include "gp_args.h"

namespace {

#ifndef SLB_FT
#define SLB_FT 64
#endif
#ifndef SLB_MEAN_MINB
#define SLB_MEAN_MINB 7
#endif
constexpr int FT = SLB_FT;             // stage 1: threads per CTA = points per CTA (1024 CTAs at
                                       // 256 x 256, 7 resident per SM: single wave, 98.8% balanced)
constexpr int HR = SLB_HEAD_RANK;
constexpr int HT = 512;                // stage 2: threads per CTA (16 warps, 8 or 2 list entries each)
constexpr int HEAD_CTAS = 148;         // one CTA per SM (it stages the head factors in shared memory)
constexpr int64_t CHUNK = 1 << 22;     // points per pass of the three stages (bounds the workspace)
constexpr int64_t WS_HEAD = 64 + SLB_SPLIT_TICKET_BYTES + (int64_t)SLB_SPLIT_PARTIAL_BYTES;   // bytes before the lists


// terms of one undecided point, carried from stage 1 to stage 2
struct filter_side { double dec0, thr, guard, coef[SLB_MAX_OUT], z[SLB_MAX_IN], dm[SLB_MAX_OUT]; };
// fp32 screening stage: dec0 = V(x), coef[j] = screened mean of output j, dm[j] = its certified bound
// (the head stage rebuilds the mean-dependent terms from them, and from an fp64 mean where needed)

struct filter_args {
    int64_t n;
    int64_t idx_begin;
    uint8_t* negative;
    double* values;
    int64_t* list_a;                   // undecided after stage 1 (index relative to the range)
    filter_side* side_a;               // their terms, same order
    int64_t* list_b;                   // undecided after stage 2 -> full posterior
    unsigned long long* counts;        // [0] entries of list_a, [1] entries of list_b
    unsigned long long* stats;         // nullptr or [4], see slb200.h
    int chunk_rows;                    // training rows per staged slice (multiple of 8)
    int max_outputs_per_factor;
    int head_factors_staged;           // head stage: factors whose tables fit in shared memory (the
                                       // others are read from global memory)
    int screened;                      // stage 1 was the fp32 screening kernel: list A entries carry only
                                       // z, threshold and V(x) (in dec0); the head stage computes the
                                       // fp64 mean and the terms that depend on it
    int mean_off[SLB_MAX_OUT];         // screened: offset (doubles) of factor f's [Xf | gamma_f ...] block
                                       // in the head stage's shared-memory copy
    int mean_doubles;                  // screened: size of that copy
    int prefetch_factors;              // head stage: warm L2 with the packed factors for the refine pass
    double* probe_mu;                  // slb_debug_screening_probe: nullptr or [n, D] screened means ...
    double* probe_dm;                  // ... and their certified error bounds (inf: point left to fp64)
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

// V(mu), L_V(mu) -> the terms of the comparison that depend on the mean: dec0 = V(mu) - V(x), the
// coefficient of every sigma_j, and the guard band (1e-6 relative + the mean's own error bound
// through L_V(mu), x 4)                                                      (lyapunov.py:344-352)
SLB_DEV void mean_decision_terms(const slb_sweep& cfg, filter_side& t, double vx, const double* mu,
                                 const double* mean_err) {
    const int D = cfg.gp.num_outputs;
    double vm[1];
    eval_fn_small(cfg.lyapunov, mu, vm);
    t.dec0 = f64sub(vm[0], vx);
    double lvmu = 0.0;                      // sum_j |L_V(mu)_j mu_j|: scale of V's sensitivity to mu
    double lverr = 0.0;                     // sum_j |L_V(mu)_j| mean_err_j
    {
        double lv[SLB_MAX_OUT];
        int nl = 1;
        if (cfg.lipschitz_v.kind != SLB_FN_NONE) nl = eval_fn_small(cfg.lipschitz_v, mu, lv);
        else lv[0] = cfg.lv_const;
        for (int j = 0; j < SLB_MAX_OUT; ++j) {
            const double l = j < D ? (nl == 1 ? lv[0] : lv[j]) : 0.0;
            t.coef[j] = j < D ? l * cfg.gp.outputs[j].beta : 0.0;
            if (j < D) { lvmu += fabs(l * mu[j]); lverr += fabs(l) * mean_err[j]; }
        }
    }
    t.guard = 1e-6 * (fabs(vm[0]) + fabs(vx) + fabs(t.thr) + lvmu) + 4.0 * lverr + 1e-300;
}

// The screening stage knows the mean only to within dm_j (certified, gp_mean_staged.cuh).  For the
// function kinds it is enabled for -- V = x^T P x (QUADRATIC, optional scale), L_V constant or a
// LINEAR map with abs / 1-norm / scale -- the change of the comparison over the box mu +- dm is
// bounded in closed form:  |V(mu + e) - V(mu)| <= sum_i |((P + P^T) mu)_i| dm_i + sum_ij |P_ij| dm_i
// dm_j;  |L_V(mu + e)_j - L_V(mu)_j| <= |scale| sum_i |A_ji| dm_i (all rows for the 1-norm), which
// enters with beta_j sigma_j <= beta_j shi_j.  Returns the amount to add to the guard band.
SLB_DEV double screening_slack(const slb_sweep& cfg, const double* mu, const double* dm,
                               const double* shi) {
    // straight-line for up to 4 outputs (screening_applicable): operands in registers, all matrix
    // entries loaded at once -- this runs once per grid point in the screening kernel's epilogue
    constexpr int NS = 4;
    const int D = cfg.gp.num_outputs;
    const slb_function& V = cfg.lyapunov;
    const int n = V.in_dim;
    double m[NS], d[NS], bs[NS], P[NS][NS];
#pragma unroll
    for (int i = 0; i < NS; ++i) {
        m[i] = i < n ? mu[i] : 0.0;
        d[i] = i < n ? dm[i] : 0.0;
        bs[i] = i < D ? fabs(cfg.gp.outputs[i].beta) * shi[i] : 0.0;
#pragma unroll
        for (int j = 0; j < NS; ++j) P[i][j] = (i < n && j < n) ? __ldg(V.matrix + i * n + j) : 0.0;
    }
    double dv = 0.0;
#pragma unroll
    for (int i = 0; i < NS; ++i) {
        double gi = 0.0;
#pragma unroll
        for (int r = 0; r < NS; ++r) gi += m[r] * (P[r][i] + P[i][r]);
        dv += fabs(gi) * d[i];
#pragma unroll
        for (int j = 0; j < NS; ++j) dv += fabs(P[i][j]) * d[i] * d[j];
    }
    if (V.flags & SLB_FLAG_SCALE) dv *= fabs(V.out_scale);
    double dl = 0.0;
    const slb_function& L = cfg.lipschitz_v;
    if (L.kind == SLB_FN_LINEAR) {
        const double sc = (L.flags & SLB_FLAG_SCALE) ? fabs(L.out_scale) : 1.0;
        const int mi = L.in_dim, mo = L.out_dim;
        double row[NS];                           // row[o] = sum_i |A_oi| dm_i
#pragma unroll
        for (int o = 0; o < NS; ++o) {
            row[o] = 0.0;
#pragma unroll
            for (int i = 0; i < NS; ++i)
                if (o < mo && i < mi) row[o] += fabs(__ldg(L.matrix + o * mi + i)) * d[i];
        }
        if ((L.flags & SLB_FLAG_NORM1) || mo == 1) {
            dl = sc * (row[0] + row[1] + row[2] + row[3]) * (bs[0] + bs[1] + bs[2] + bs[3]);
        } else {
#pragma unroll
            for (int j = 0; j < NS; ++j) dl += sc * row[j] * bs[j];
        }
    }
    return 1.000001 * (dv + dl);
}

template <int DIN>
__global__ void __launch_bounds__(FT, SLB_MEAN_MINB)
filter_mean_kernel(const __grid_constant__ slb_sweep cfg, const filter_args a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    pdl_launch_dependents();                      // the head stage may start staging its tables
    prefetch_descriptor_operands(cfg);
    mean_pipe P;
    double *tab512, *tab64;
    mean_pipe_setup(P, smem_raw, DIN, a.chunk_rows, a.max_outputs_per_factor, cfg.gp, &tab512, &tab64);
    uint64_t* bar = P.bar;
    if (threadIdx.x == 0) mean_pipe_init(P, tab512);
    mean_pipe_start<DIN>(cfg.gp, P);                                   // slice 0 -> buffer 0
    __syncthreads();

    const int64_t rel0 = (int64_t)blockIdx.x * FT + threadIdx.x;
    const bool valid = rel0 < a.n;
    const int64_t rel = valid ? rel0 : a.n - 1;   // every thread stays for the block barriers
    const int d = cfg.grid.ndim;
    const int D = cfg.gp.num_outputs;

    // ---- x, V(x), threshold(x), u = policy(x)           (lyapunov.py:436, 284-288)
    filter_side t;
    grid_index_to_state(cfg.grid, a.idx_begin + rel, t.z);
    double vx;
    lyapunov_state_terms(cfg, t.z, a.idx_begin + rel, &vx, &t.thr);
    {
        double u[SLB_MAX_OUT];
        const int m = eval_fn_small(cfg.policy, t.z, u);
        for (int c = 0; c < m; ++c) t.z[d + c] = u[c];
    }
    // the expanded squared distance needs moderate magnitudes; NaN / huge inputs -> full path
    bool sane = true;
#pragma unroll
    for (int c = 0; c < DIN; ++c) sane &= fabs(t.z[c]) < 1e100;

    // ---- posterior mean of every output (functions.py:439-442 as k . L^-T alpha)
    slb_bulk::mbar_wait(bar + 2, 0);              // exp tables have landed
    double mu[SLB_MAX_OUT];
    double mean_err[SLB_MAX_OUT];
    gp_mean_staged<DIN, true>(cfg.gp, t.z, mu, mean_err, tab512, tab64, P);

    mean_decision_terms(cfg, t, vx, mu, mean_err);

    // ---- sigma_j <= prior sigma_j     (functions.py:450 without data)
    double shi[SLB_MAX_OUT];
    for (int j = 0; j < D; ++j) {
        const slb_gp_factor& F = cfg.gp.factors[cfg.gp.outputs[j].factor];
        shi[j] = sqrt(F.kernel.num_prims > 0 ? kernel_expr_diag<DIN>(F.kernel, t.z) : F.variance);
    }
    const int outcome = sane ? decide(t, shi, D) : -1;
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

// ---- stage 1, fp32 screening variant ---------------------------------------------------------------
// Same role as filter_mean_kernel with the mean from the fp32 scheme of gp_mean_staged.cuh (three FFMA
// and one MUFU.EX2 per kernel value instead of twelve fp64 operations) and its certified error bound
// dm: a point is decided when the comparison has the same outcome for every mean in mu +- dm and every
// sigma between 0 and the prior's (screening_slack).  The undecided points go to list A with z,
// threshold and V(x) only; the head stage recomputes their mean in fp64 (warp-cooperatively, on ~8% of
// the grid at C2), so everything downstream of this kernel is the fp64 arithmetic of the other path.
template <int DIN>
__global__ void __launch_bounds__(FT, SLB_MEAN_MINB)
filter_mean32_kernel(const __grid_constant__ slb_sweep cfg, const filter_args a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ double s_cen[SLB_MAX_IN];
    pdl_launch_dependents();                      // the head stage may start staging its tables
    prefetch_descriptor_operands(cfg);
    constexpr int W32 = row32<DIN>::W;
    mean_pipe P;
    P.bar = reinterpret_cast<uint64_t*>(smem_raw);
    P.C = a.chunk_rows;
    P.xstride = P.C * (DIN + 1);
    P.gstride = P.C * a.max_outputs_per_factor;
    P.xbuf = reinterpret_cast<double*>(smem_raw + 32);
    P.gbuf = P.xbuf + 2 * P.xstride;
    P.t = 0;
    mean32_bufs B;
    B.red = P.gbuf + 2 * P.gstride;
    B.xf = reinterpret_cast<float*>(B.red + 8 * (FT / 32) + 8);
    B.g = B.xf + P.C * W32;
    if (threadIdx.x == 0) {
        slb_bulk::mbar_init(P.bar + 0, 1);
        slb_bulk::mbar_init(P.bar + 1, 1);
        slb_bulk::fence_barrier_init();
        slb_bulk::fence_proxy_async();
    }
    mean_pipe_start<DIN>(cfg.gp, P);                                   // slice 0 -> buffer 0

    const int64_t rel0 = (int64_t)blockIdx.x * FT + threadIdx.x;
    const bool valid = rel0 < a.n;
    const int64_t rel = valid ? rel0 : a.n - 1;   // every thread stays for the block barriers
    const int d = cfg.grid.ndim;
    const int D = cfg.gp.num_outputs;

    // ---- x, V(x), threshold(x), u = policy(x)           (lyapunov.py:436, 284-288)
    filter_side t;
    grid_index_to_state(cfg.grid, a.idx_begin + rel, t.z);
    double vx;
    lyapunov_state_terms(cfg, t.z, a.idx_begin + rel, &vx, &t.thr);
    {
        double u[SLB_MAX_OUT];
        const int m = eval_fn_small(cfg.policy, t.z, u);
        for (int c = 0; c < m; ++c) t.z[d + c] = u[c];
    }
    bool sane = true;
#pragma unroll
    for (int c = 0; c < DIN; ++c) sane &= fabs(t.z[c]) < 1e100;
    // the CTA's centre: the query point of its middle thread
    if (threadIdx.x == FT / 2) {
#pragma unroll
        for (int c = 0; c < DIN; ++c) s_cen[c] = t.z[c];
    }
    __syncthreads();                              // centre visible; barriers initialised
    double zcen[DIN];
#pragma unroll
    for (int c = 0; c < DIN; ++c) zcen[c] = s_cen[c];

    double mu[SLB_MAX_OUT], dm[SLB_MAX_OUT];
    gp_mean32_staged<DIN>(cfg.gp, t.z, zcen, mu, dm, sane, P, B);
    if (a.probe_mu != nullptr && valid) {
        for (int o = 0; o < D; ++o) {
            a.probe_mu[rel * D + o] = mu[o];
            a.probe_dm[rel * D + o] = sane ? dm[o] : __longlong_as_double(0x7ff0000000000000ll);
        }
    }

    // ---- the comparison over mu +- dm and sigma_j in [0, prior sigma_j]
    double zero[SLB_MAX_OUT];
    for (int j = 0; j < SLB_MAX_OUT; ++j) zero[j] = 0.0;
    mean_decision_terms(cfg, t, vx, mu, zero);
    double shi[SLB_MAX_OUT];
    for (int j = 0; j < D; ++j) shi[j] = sqrt(cfg.gp.factors[cfg.gp.outputs[j].factor].variance);
    t.guard += screening_slack(cfg, mu, dm, shi);
    const int outcome = sane ? decide(t, shi, D) : -1;
    const bool undecided = valid && outcome < 0;
    if (valid) {
        a.negative[rel] = outcome > 0 ? 1 : 0;
        if (a.values != nullptr) a.values[rel] = vx;
    }
    const long long slot = list_append(undecided, a.counts + 0);
    if (undecided) {
        filter_side* dst = a.side_a + slot;
        dst->dec0 = vx;                           // the head stage rebuilds the mean-dependent terms
        dst->thr = t.thr;
#pragma unroll
        for (int c = 0; c < DIN; ++c) dst->z[c] = t.z[c];
        for (int j = 0; j < D; ++j) {
            dst->coef[j] = mu[j];
            dst->dm[j] = sane ? dm[j] : __longlong_as_double(0x7ff0000000000000ll);
        }
        a.list_a[slot] = rel;
    }
    if (a.stats != nullptr) {
        count_stat(valid && !undecided, a.stats + 0);
        count_stat(valid, a.stats + 3);
    }
}

// ---- stage 2: variance given the head subset, one warp per HP undecided points ---------------------
// One CTA per SM, 8 warps.  The head factors W = L_S^-1 (column-major, zero padded, 32 KB each) and
// the subset's inputs are staged ONCE per CTA in shared memory by TMA bulk copies (read from
// global memory per point they cost an L2/HBM round trip per column: measured 47 us for 5000
// points); then every warp walks the list in groups of HP = 8 points.  Lane l owns rows l and l + 32
// of a = W k for all HP points (16 independent FMA chains): a column of W read from shared memory
// serves 8 points -- one point per warp made the stage shared-memory-bandwidth bound (6.4 ms for
// the 1.9 M list entries of a 2048 x 2048 grid).  The kernel values k_j of the HR subset points are
// computed two per lane and point and exchanged through shared memory ([row][point]: one row's
// HP values are four broadcast 128-bit loads); lane p < HP makes the decision of point p.
constexpr int HW = HT / 32;            // warps per CTA
constexpr int HP = 8;                  // list entries per warp iteration (long lists)
constexpr int HP_SHORT = 2;            // ... when the list has fewer than HP entries per warp of the grid:
                                       // the latency of one group is the whole stage then

// screened lists: fp64 mean of one factor's NO outputs at the lane's point.  L lanes (a power of two,
// 4..32, consecutive in the warp) share a point: lane r of them takes rows r, r + L, ... of the
// factor's shared-memory copy [Xf | gamma_f ...] and the parts are summed by log2(L) shuffles.  The
// arithmetic of mean_factor<.., FAST = true>: expanded distance, exp_neg_fast, two partial sums.
template <int DIN, int NO>
SLB_DEV void head_mean_factor(const double* __restrict__ xf, int Mp, const double* zs, double zz, int r,
                              int L, const double* __restrict__ tab512, double* dot) {
    constexpr int W = DIN + 1;
    const double* __restrict__ gm = xf + (size_t)Mp * W;
    double d0[NO], d1[NO];
#pragma unroll
    for (int q = 0; q < NO; ++q) { d0[q] = 0.0; d1[q] = 0.0; }
    for (int j = r; j < Mp; j += 4 * L) {
        double arg[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int jc = min(j + L * u, Mp - 1);
            double row[W];
            load_row<W>(xf + jc * W, row);
            double acc = row[DIN] + zz;
#pragma unroll
            for (int c = 0; c < DIN; ++c) acc = fma(zs[c], row[c], acc);
            arg[u] = acc;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            bool far;
            double k = exp_neg_fast(arg[u], tab512, far);
            k = (far || j + L * u >= Mp) ? 0.0 : k;
            const int jc = min(j + L * u, Mp - 1);
#pragma unroll
            for (int q = 0; q < NO; ++q) {
                if (u & 1) d1[q] = fma(k, gm[q * Mp + jc], d1[q]);
                else d0[q] = fma(k, gm[q * Mp + jc], d0[q]);
            }
        }
    }
#pragma unroll
    for (int q = 0; q < NO; ++q) {
        double v = d0[q] + d1[q];
        for (int off = L >> 1; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        dot[q] = v;
    }
}

// screened lists, once per round of the head kernel: the fp64 means of the entries the round could not
// decide from their screened means (`slots`: s = 8 w + p is entry p of warp w's group), by ALL warps of the
// CTA -- L = 512 / (number of entries) lanes per point, so a CTA with few of them (short lists: the stage's
// duration is the latency of one group) still spreads the M exps per point and factor over its threads.
// Results: mu_s / merr_s [slot][SLB_MAX_OUT] in shared memory.
template <int DIN>
SLB_DEV void head_round_means(const slb_sweep& cfg, const filter_args& a, const int* slots, int nneed,
                              int64_t grp0, int64_t count, const double* mbuf, const double* tab512,
                              double* mu_s, double* merr_s) {
    const int nf = cfg.gp.num_factors, D = cfg.gp.num_outputs;
    const int L = nneed <= 16 ? 32 : nneed <= 32 ? 16 : nneed <= 64 ? 8 : 4;
    const int r = threadIdx.x & (L - 1);
    const int per_warp = 32 / L;
    const int nloop = (nneed + per_warp - 1) / per_warp * per_warp;   // whole warps take part in the shuffles
    for (int i = threadIdx.x / L; i < nloop; i += HT / L) {
        const bool live = i < nneed;
        const int slot = live ? slots[i] : 0;
        const int64_t k = (grp0 + (int64_t)(slot / HP) * gridDim.x) * HP + (slot % HP);
        double z[DIN];
#pragma unroll
        for (int c = 0; c < DIN; ++c) z[c] = (live && k < count) ? a.side_a[k].z[c] : 0.0;
        for (int f = 0; f < nf; ++f) {
            const slb_gp_factor& F = cfg.gp.factors[f];
            int outs[SLB_MAX_OUT];
            int no = 0;
            for (int o = 0; o < D; ++o)
                if (cfg.gp.outputs[o].factor == f) outs[no++] = o;
            double zs[DIN];
            double zz = 0.0;
#pragma unroll
            for (int c = 0; c < DIN; ++c) {
                zs[c] = z[c] / F.lengthscales[c];
                zz = fma(zs[c], zs[c], zz);
            }
            zz *= -0.5;
            const int Mp = padded_rows(F.M);
            const double* xf = mbuf + a.mean_off[f];
            double dot[SLB_MAX_OUT];
            switch (no) {
            case 1: head_mean_factor<DIN, 1>(xf, Mp, zs, zz, r, L, tab512, dot); break;
            case 2: head_mean_factor<DIN, 2>(xf, Mp, zs, zz, r, L, tab512, dot); break;
            case 3: head_mean_factor<DIN, 3>(xf, Mp, zs, zz, r, L, tab512, dot); break;
            case 4: head_mean_factor<DIN, 4>(xf, Mp, zs, zz, r, L, tab512, dot); break;
            case 5: head_mean_factor<DIN, 5>(xf, Mp, zs, zz, r, L, tab512, dot); break;
            case 6: head_mean_factor<DIN, 6>(xf, Mp, zs, zz, r, L, tab512, dot); break;
            default: break;
            }
            if (r == 0 && live) {
                for (int q = 0; q < no; ++q)
                    mean_output_finish<DIN>(F, cfg.gp.outputs[outs[q]], z, dot[q], zz, 1.0, false,
                                            &mu_s[slot * SLB_MAX_OUT + outs[q]],
                                            &merr_s[slot * SLB_MAX_OUT + outs[q]]);
            }
        }
    }
}

// one group of P list entries [g P, g P + P) on one warp
template <int DIN, int P, bool ALL_STAGED>
SLB_DEV void head_group_bound(const slb_sweep& cfg, const filter_args& a, int64_t grp, int64_t count,
                              const double* exptab, double* kw, const double* wbuf, const double* xbuf,
                              filter_side& t, int64_t& rel, bool& mine, double* shi) {
    const int lane = threadIdx.x & 31;
    const int nf = cfg.gp.num_factors;
    const int D = cfg.gp.num_outputs;
    // lane p < P owns list entry grp * P + p: its terms, its index and finally its decision
    const int64_t k = grp * P + min(lane, P - 1);
    mine = lane < P && k < count;
    t = filter_side{};
    rel = 0;
    if (mine) { t = a.side_a[k]; rel = a.list_a[k]; }
    for (int j = 0; j < D; ++j) {
        const slb_gp_factor& F = cfg.gp.factors[cfg.gp.outputs[j].factor];
        shi[j] = mine ? sqrt(F.kernel.num_prims > 0 ? kernel_expr_diag<DIN>(F.kernel, t.z) : F.variance)
                      : 0.0;
    }
    for (int f = 0; f < nf; ++f) {
        const slb_gp_factor& F = cfg.gp.factors[f];
        const int rows = F.head_rows;
        if (rows <= 0) continue;
        const bool general = F.kernel.num_prims > 0;
        const double s2 = f64mul(F.scale, F.scale);
        // ALL_STAGED: the tables are known to be in shared memory (LDS instead of generic loads)
        const bool staged = ALL_STAGED || f < a.head_factors_staged;
        const double* xh = xbuf + (size_t)f * HR * DIN;
        if (!ALL_STAGED && !staged) xh = F.Xhead;
        // kernel values of every point of the group against subset points lane and lane + 32
        // (functions.py:438); entries beyond the list carry zeros (never decided)
        double zown[DIN];                       // this lane's point in the factor's units (one division
#pragma unroll                                  // per lane and dimension instead of one per point)
        for (int c = 0; c < DIN; ++c) zown[c] = general ? t.z[c] : t.z[c] / F.lengthscales[c];
#pragma unroll
        for (int p = 0; p < P; ++p) {
            double zs[DIN];
#pragma unroll
            for (int c = 0; c < DIN; ++c) zs[c] = __shfl_sync(0xffffffffu, zown[c], p);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int j = lane + 32 * h;
                double kv = 0.0;
                if (j < rows) {
                    const double* xr = xh + j * DIN;
                    if (general) {
                        kv = kernel_expr_cross<DIN>(F.kernel, zs, xr, exptab);
                    } else {
                        double a2 = 0.0;
#pragma unroll
                        for (int c = 0; c < DIN; ++c) { const double df = zs[c] - xr[c]; a2 = fma(df, df, a2); }
                        kv = F.variance * exp_neg_tab(-0.5 * a2, exptab);
                    }
                    kv = s2 * kv;
                }
                kw[j * P + p] = kv;
            }
        }
        __syncwarp();
        double ssp = 0.0;                       // lane p: sum_i a_i^2 of point p
        if constexpr (P == HP) {
            // a = W k on the fp64 tensor pipe: W (64 x 64, lower triangular) pre-packed in DMMA
            // A-fragment order (row block b, k-step s: slb_gp_factor.Wheadp), the HP = 8 points are
            // the n dimension, k values [row][point] in shared memory are the B fragments as they
            // lie.  Only the blocks on or below the diagonal (s <= 2 b + 1) are multiplied: 72 DMMAs.
            const double* __restrict__ Wp = wbuf + (size_t)f * HR * HR;
            if (!ALL_STAGED && !staged) Wp = F.Wheadp;
            double acc[8][2];
#pragma unroll
            for (int b = 0; b < 8; ++b) { acc[b][0] = 0.0; acc[b][1] = 0.0; }
#pragma unroll
            for (int sk = 0; sk < 16; ++sk) {
                const double bf = kw[(4 * sk + (lane & 3)) * HP + (lane >> 2)];
#pragma unroll
                for (int b = sk / 2; b < 8; ++b) {
                    const double af = Wp[(b * 16 + sk) * 32 + lane];
                    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                                 : "+d"(acc[b][0]), "+d"(acc[b][1]) : "d"(af), "d"(bf));
                }
            }
            __syncwarp();
            // lane T holds rows 8 b + T/4 of points 2 (T%4), 2 (T%4) + 1: square, sum over b, then over
            // the 8 lanes that share T%4; lane p fetches point p's sum
            double v0 = 0.0, v1 = 0.0;
#pragma unroll
            for (int b = 0; b < 8; ++b) { v0 = fma(acc[b][0], acc[b][0], v0); v1 = fma(acc[b][1], acc[b][1], v1); }
#pragma unroll
            for (int off = 4; off < 32; off <<= 1) {
                v0 += __shfl_xor_sync(0xffffffffu, v0, off);
                v1 += __shfl_xor_sync(0xffffffffu, v1, off);
            }
            const double s0 = __shfl_sync(0xffffffffu, v0, (lane >> 1) & 3);
            const double s1 = __shfl_sync(0xffffffffu, v1, (lane >> 1) & 3);
            ssp = (lane & 1) ? s1 : s0;
        } else {
        // short groups: two partial sums per row and point (even / odd columns) halve the dependent
        // FMA chain; 8-point groups already carry 16 independent chains
        constexpr int NS = P <= 2 ? 2 : 1;
        double al[2][P], ah[2][P];
#pragma unroll
        for (int p = 0; p < P; ++p) { al[0][p] = al[1][p] = 0.0; ah[0][p] = ah[1][p] = 0.0; }
        const double* Wt = F.Whead;            // column-major table (global / L2): reference path
#pragma unroll 4
        for (int j = 0; j < HR; ++j) {
            const double wl = Wt[j * HR + lane], wh = Wt[j * HR + 32 + lane];
            double kj[P];
#pragma unroll
            for (int p = 0; p < P; p += 2) {
                const double2 v = *reinterpret_cast<const double2*>(kw + j * P + p);
                kj[p] = v.x; kj[p + 1] = v.y;
            }
#pragma unroll
            for (int p = 0; p < P; ++p) {
                al[j & (NS - 1)][p] = fma(wl, kj[p], al[j & (NS - 1)][p]);
                ah[j & (NS - 1)][p] = fma(wh, kj[p], ah[j & (NS - 1)][p]);
            }
        }
        __syncwarp();
        // sum a^2 per point over the 64 rows: lane p ends up with point p's
#pragma unroll
        for (int p = 0; p < P; ++p) {
            const double lo = al[0][p] + al[1][p], hi = ah[0][p] + ah[1][p];
            double ss = fma(lo, lo, hi * hi);
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, off);
            if (lane == p) ssp = ss;
        }
        }
        double kss = F.kss;
        if (general && mine) kss = s2 * kernel_expr_diag<DIN>(F.kernel, t.z);
        const double sdev = sqrt(f64sub(kss, ssp) / s2);          // NaN if negative
        for (int j = 0; j < D; ++j)
            if (cfg.gp.outputs[j].factor == f) shi[j] = sdev;
    }
}

// a lane's final outcome: the flag of a decided point, list B for an undecided one (warp-collective)
SLB_DEV void head_group_finish(const filter_args& a, bool mine, int outcome, int64_t rel, unsigned* s_stat) {
    const int lane = threadIdx.x & 31;
    const bool undecided = mine && outcome < 0;
    if (mine && outcome >= 0) a.negative[rel] = outcome > 0 ? 1 : 0;
    const long long slot = list_append(undecided, a.counts + 1);
    if (undecided) a.list_b[slot] = rel;
    const unsigned dec = __ballot_sync(0xffffffffu, mine && !undecided);
    const unsigned und = __ballot_sync(0xffffffffu, undecided);
    if (lane == 0) {
        if (dec) atomicAdd(s_stat + 0, (unsigned)__popc(dec));
        if (und) atomicAdd(s_stat + 1, (unsigned)__popc(und));
    }
}

template <int DIN>
__global__ void __launch_bounds__(HT, 1)
filter_head_kernel(const __grid_constant__ slb_sweep cfg, const filter_args a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    pdl_launch_dependents();                      // the refine launch may get its CTAs ready
    prefetch_descriptor_operands(cfg);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);             // [1]
    unsigned* s_stat = reinterpret_cast<unsigned*>(smem_raw + 8);      // decided / undecided by this CTA
    double* tab512 = reinterpret_cast<double*>(smem_raw + 16);         // [512] (screened lists: exp_neg_fast)
    double* exptab = tab512 + 512;                                     // [64]
    double* kbuf = exptab + 64;                                        // [HW][HR][HP]
    double* wbuf = kbuf + HW * HR * HP;                                // [staged][HR * HR]
    const int nf = cfg.gp.num_factors;
    double* xbuf = wbuf + (size_t)a.head_factors_staged * HR * HR;     // [staged][HR * DIN]
    double* mbuf = xbuf + (size_t)a.head_factors_staged * HR * DIN;    // screened: [Xf | gamma_f ...] per factor
    // ---- everything that does not depend on stage 1 first (the kernel is a programmatic dependent of
    // it: this part overlaps stage 1's tail).  The refine pass that follows streams every factor's
    // packed L^-1 (1 MB at M = 500); if it is not L2-resident by then (first sweep after a cache
    // update, or evicted in between) its CTAs start with HBM round trips in lockstep: prefetch it.
    if (a.prefetch_factors) {
        for (int f = 0; f < nf; ++f) {
            const slb_gp_factor& F = cfg.gp.factors[f];
            const char* base = reinterpret_cast<const char*>(F.Wpack);
            const size_t nbytes = (size_t)F.nrb * (F.nrb + 1) * 32 * sizeof(double);
            for (size_t off = ((size_t)blockIdx.x * HT + threadIdx.x) * 128; off < nbytes;
                 off += (size_t)gridDim.x * HT * 128)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(base + off));
            // ... and the training inputs its generation phases stage panel by panel
            const char* xs = reinterpret_cast<const char*>(F.Xs);
            const size_t xbytes = (size_t)F.M * DIN * sizeof(double);
            if (blockIdx.x == (unsigned)f)
                for (size_t off = (size_t)threadIdx.x * 128; off < xbytes; off += (size_t)HT * 128)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(xs + off));
        }
        if (blockIdx.x == gridDim.x - 1 && threadIdx.x < cfg.gp.num_outputs) {
            const slb_gp_output& G = cfg.gp.outputs[threadIdx.x];
            const size_t abytes = (size_t)cfg.gp.factors[G.factor].M * sizeof(double);
            for (size_t off = 0; off < abytes; off += 128)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(G.alpha) + off));
        }
    }
    if (threadIdx.x == 0) {
        slb_bulk::mbar_init(bar, 1);
        slb_bulk::fence_barrier_init();
        slb_bulk::fence_proxy_async();
        unsigned bytes = 576 * sizeof(double);
        for (int f = 0; f < a.head_factors_staged; ++f)
            if (cfg.gp.factors[f].head_rows > 0)
                bytes += (unsigned)(HR * HR + HR * DIN) * sizeof(double);
        if (a.screened) bytes += (unsigned)a.mean_doubles * sizeof(double);
        slb_bulk::mbar_arrive_expect_tx(bar, bytes);
        slb_bulk::copy_g2s(tab512, g_exp_tables, 576 * sizeof(double), bar);
        for (int f = 0; f < a.head_factors_staged; ++f) {
            const slb_gp_factor& F = cfg.gp.factors[f];
            if (F.head_rows <= 0) continue;
            slb_bulk::copy_g2s(wbuf + (size_t)f * HR * HR, F.Wheadp, HR * HR * sizeof(double), bar);
            slb_bulk::copy_g2s(xbuf + (size_t)f * HR * DIN, F.Xhead, HR * DIN * sizeof(double), bar);
        }
        if (a.screened) {
            for (int f = 0; f < nf; ++f) {
                const slb_gp_factor& F = cfg.gp.factors[f];
                const int Mp = padded_rows(F.M);
                if (Mp == 0) continue;
                double* dst = mbuf + a.mean_off[f];
                slb_bulk::copy_g2s(dst, F.Xf, (unsigned)(Mp * (DIN + 1)) * sizeof(double), bar);
                dst += (size_t)Mp * (DIN + 1);
                for (int o = 0; o < cfg.gp.num_outputs; ++o) {
                    if (cfg.gp.outputs[o].factor != f) continue;
                    slb_bulk::copy_g2s(dst, cfg.gp.outputs[o].gamma_f, (unsigned)Mp * sizeof(double), bar);
                    dst += Mp;
                }
            }
        }
    }
    if (threadIdx.x < 2) s_stat[threadIdx.x] = 0;
    if (a.screened && threadIdx.x < 2)
        reinterpret_cast<int*>(mbuf + a.mean_doubles + 2 * HW * HP * SLB_MAX_OUT)[threadIdx.x] = 0;
    __syncthreads();
    slb_bulk::mbar_wait(bar, 0);                  // also before leaving: the copies land in this CTA's memory
    pdl_wait();                                   // ---- stage 1 has completed: its lists are visible
    const int64_t count = (int64_t)a.counts[0];
    const int64_t nwarps = (int64_t)gridDim.x * HW;
    const int64_t ngroups = (count + HP - 1) / HP;
    // groups are dealt round-robin over the CTAs (group g: CTA g % gridDim, warp g / gridDim): a short
    // list spreads over all SMs instead of filling the 16 warps of the first few
    if ((int64_t)blockIdx.x >= ngroups) return;                        // no group for this CTA
    const int warp = threadIdx.x >> 5;
    double* kw = kbuf + warp * HR * HP;
    double* mu_s = mbuf + a.mean_doubles;      // screened: [HW * HP][SLB_MAX_OUT] fp64 means, then their error bounds
    double* merr_s = mu_s + HW * HP * SLB_MAX_OUT;
    int* need_s = reinterpret_cast<int*>(merr_s + HW * HP * SLB_MAX_OUT);   // [2] counters (round parity), slots
    // round: warp w takes group grp0 + w gridDim (every warp of the CTA makes the same number of rounds)
    int round = 0;
    for (int64_t grp0 = blockIdx.x; grp0 < ngroups; grp0 += nwarps, ++round) {
        const int64_t grp = grp0 + (int64_t)warp * gridDim.x;
        const bool active = grp < ngroups;
        const int lane = threadIdx.x & 31;
        filter_side t;
        int64_t rel = 0;
        bool mine = false, need = false;
        double shi[SLB_MAX_OUT];
        double vx = 0.0;
        int outcome = 0;
        if (active) {
            if (a.head_factors_staged == nf)
                head_group_bound<DIN, HP, true>(cfg, a, grp, count, exptab, kw, wbuf, xbuf, t, rel, mine, shi);
            else
                head_group_bound<DIN, HP, false>(cfg, a, grp, count, exptab, kw, wbuf, xbuf, t, rel, mine, shi);
            if (a.screened) {
                // first with the screened mean and its error box (stage 1 left them in the entry): most
                // entries are decided by the tighter variance bound alone
                double mu[SLB_MAX_OUT], dm[SLB_MAX_OUT], zero[SLB_MAX_OUT];
                for (int j = 0; j < SLB_MAX_OUT; ++j) { mu[j] = t.coef[j]; dm[j] = t.dm[j]; zero[j] = 0.0; }
                vx = t.dec0;
                mean_decision_terms(cfg, t, vx, mu, zero);
                t.guard += screening_slack(cfg, mu, dm, shi);
                outcome = mine ? decide(t, shi, cfg.gp.num_outputs) : 0;
                need = mine && outcome < 0;
                if (need) need_s[2 + atomicAdd(need_s + (round & 1), 1)] = warp * HP + lane;
            } else {
                outcome = mine ? decide(t, shi, cfg.gp.num_outputs) : 0;
            }
        }
        if (a.screened) {
            if (threadIdx.x == 0) need_s[(round + 1) & 1] = 0;       // the next round's counter
            __syncthreads();
            const int nneed = need_s[round & 1];
            if (nneed > 0) {
                // the rest gets its mean in fp64 (all warps), then the same comparison as the fp64 path
                head_round_means<DIN>(cfg, a, need_s + 2, nneed, grp0, count, mbuf, tab512, mu_s, merr_s);
                __syncthreads();
                if (need) {
                    double mu[SLB_MAX_OUT], merr[SLB_MAX_OUT];
                    const int slot = warp * HP + lane;
                    for (int j = 0; j < SLB_MAX_OUT; ++j) {
                        mu[j] = mu_s[slot * SLB_MAX_OUT + j];
                        merr[j] = merr_s[slot * SLB_MAX_OUT + j];
                    }
                    mean_decision_terms(cfg, t, vx, mu, merr);
                    outcome = decide(t, shi, cfg.gp.num_outputs);
                }
            }
        }
        if (active) head_group_finish(a, mine, outcome, rel, s_stat);
    }
    // one pair of global atomics per CTA (one per point serialised on the counter's L2 line)
    __syncthreads();
    if (a.stats != nullptr && threadIdx.x < 2 && s_stat[threadIdx.x] != 0)
        atomicAdd(a.stats + 1 + threadIdx.x, (unsigned long long)s_stat[threadIdx.x]);
}

double* g_probe_mu = nullptr;          // slb_debug_screening_probe
double* g_probe_dm = nullptr;
int g_filter_stages = 3;               // slb_debug_filter_stages: bit 0 head stage, bit 1 refine pass,
                                       // bit 2 forces the fp64 mean stage (no fp32 screening)

// The fp32 screening kernel needs closed-form bounds of V and L_V over a box of means
// (screening_slack): plain RBF factors, V = QUADRATIC (optional scale) on the GP outputs, L_V constant
// or LINEAR with abs / 1-norm / scale only.  Everything else keeps the fp64 mean stage.
bool screening_applicable(const slb_sweep& cfg) {
    if (g_filter_stages & 4) return false;
    const int D = cfg.gp.num_outputs;
    for (int f = 0; f < cfg.gp.num_factors; ++f)
        if (cfg.gp.factors[f].kernel.num_prims > 0) return false;
    const slb_function& V = cfg.lyapunov;
    if (D > 4) return false;                   // screening_slack is written out for up to 4 outputs
    if (V.kind != SLB_FN_QUADRATIC || V.in_dim != D || (V.flags & ~(uint32_t)SLB_FLAG_SCALE)) return false;
    const slb_function& L = cfg.lipschitz_v;
    if (L.kind == SLB_FN_NONE) return true;
    if (L.kind != SLB_FN_LINEAR || L.in_dim != D) return false;
    if (L.flags & ~(uint32_t)(SLB_FLAG_ABS | SLB_FLAG_NORM1 | SLB_FLAG_SCALE)) return false;
    if (L.out_dim > 4) return false;
    return (L.flags & SLB_FLAG_NORM1) || L.out_dim == 1 || L.out_dim == D;
}

// Shared-memory plan of the head stage (one CTA per SM): which factors' head tables are staged, and --
// when the fp32 screening kernel is stage 1 -- every factor's [Xf | gamma_f] next to them.  Screening
// is only used when all of it fits (otherwise the fp64 mean stage runs, whose list entries are complete).
void head_layout(const slb_sweep& cfg, int din, filter_args& ah, size_t& head_smem) {
    const size_t head_fixed = 16 + (576 + HW * HR * HP) * sizeof(double);
    const size_t per_factor = (size_t)(HR * HR + HR * din) * sizeof(double);
    ah.head_factors_staged = cfg.gp.num_factors;
    while (ah.head_factors_staged > 0 && head_fixed + ah.head_factors_staged * per_factor > 226 * 1024)
        --ah.head_factors_staged;
    head_smem = head_fixed + ah.head_factors_staged * per_factor;
    ah.screened = 0;
    ah.mean_doubles = 0;
    static const int head_prefetch = [] {                  // SLB200_HEAD_PREFETCH=0: A/B timing knob
        const char* e = getenv("SLB200_HEAD_PREFETCH");
        return e ? (atoi(e) != 0) : 1;
    }();
    ah.prefetch_factors = ((g_filter_stages & 2) && head_prefetch) ? 1 : 0;
    if (screening_applicable(cfg) && ah.head_factors_staged == cfg.gp.num_factors) {
        int off = 0;
        for (int f = 0; f < cfg.gp.num_factors; ++f) {
            int no = 0;
            for (int o = 0; o < cfg.gp.num_outputs; ++o) no += cfg.gp.outputs[o].factor == f;
            ah.mean_off[f] = off;
            off += ((cfg.gp.factors[f].M + 7) & ~7) * (din + 1 + no);
        }
        const size_t extra = ((size_t)off + 2 * HW * HP * SLB_MAX_OUT) * sizeof(double) +
                             (size_t)(HW * HP + 4) * sizeof(int);
        if (head_smem + extra <= 226 * 1024) {
            ah.screened = 1;
            ah.mean_doubles = off;
            head_smem += extra;
        }
    }
}

template <int DIN>
int launch_filter(cudaStream_t st, const slb_sweep& cfg, const filter_args& a, size_t smem) {
    static std::atomic<bool> configured[64];
    int device = 0;
    SLB_CUDA(cudaGetDevice(&device));
    if (device < 0 || device >= 64 || !configured[device].load(std::memory_order_acquire)) {
        SLB_CUDA(cudaFuncSetAttribute(filter_mean_kernel<DIN>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
        SLB_CUDA(cudaFuncSetAttribute(filter_mean32_kernel<DIN>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
        SLB_CUDA(cudaFuncSetAttribute(filter_head_kernel<DIN>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        if (device >= 0 && device < 64) configured[device].store(true, std::memory_order_release);
    }
    const int64_t blocks = (a.n + FT - 1) / FT;
    filter_args ah = a;
    size_t head_smem = 0;
    head_layout(cfg, DIN, ah, head_smem);
    ah.probe_mu = g_probe_mu;
    ah.probe_dm = g_probe_dm;
    if (ah.screened) {
        const size_t smem32 = mean32_smem_bytes(DIN, a.max_outputs_per_factor, a.chunk_rows, FT / 32);
        filter_mean32_kernel<DIN><<<(unsigned)blocks, FT, smem32, st>>>(cfg, ah);
    } else {
        filter_mean_kernel<DIN><<<(unsigned)blocks, FT, smem, st>>>(cfg, a);
    }
    SLB_LAUNCH_CHECK();
    if (!(g_filter_stages & 1)) return 0;
    SLB_CUDA(slb_launch_dependent(filter_head_kernel<DIN>, dim3(HEAD_CTAS), dim3(HT), head_smem, st, cfg, ah));
    SLB_LAUNCH_CHECK();
    return 0;
}

}  // namespace

// gp_sweep.cu: the full posterior on the compacted list (count read on the device)
int slb_launch_refine(cudaStream_t st, const slb_sweep& cfg, int64_t n_max, int64_t idx_begin,
                      const int64_t* list, const unsigned long long* count, uint8_t* negative,
                      double* values, double* split_partial, int* split_ticket);

extern "C" {

int slb_debug_filter_stages(int32_t mask) {
    g_filter_stages = mask;
    return 0;
}

int slb_filter_stage1(const slb_sweep* cfg) {
    if (cfg == nullptr || cfg->gp.num_outputs <= 0) return 0;
    filter_args a;
    memset(&a, 0, sizeof(a));
    size_t smem = 0;
    head_layout(*cfg, cfg->gp.input_dim, a, smem);
    return a.screened ? 32 : 64;
}

int slb_debug_screening_probe(double* mu_dev, double* dm_dev) {
    g_probe_mu = (mu_dev != nullptr && dm_dev != nullptr) ? mu_dev : nullptr;
    g_probe_dm = g_probe_mu != nullptr ? dm_dev : nullptr;
    return 0;
}

int64_t slb_filter_workspace(int64_t n) {
    if (n < 0) n = 0;
    if (n > CHUNK) n = CHUNK;     // longer ranges are swept in passes of CHUNK points
    // [0] |list A|, [1] |list B| (uint64, 64 bytes reserved), tile tickets and partial sums of the
    // row-split refine pass, list A, list B, terms of list A
    return WS_HEAD + n * (int64_t)(2 * sizeof(int64_t) + sizeof(filter_side));
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
    int nomax = 1;
    for (int f = 0; f < cfg->gp.num_factors; ++f) {
        const slb_gp_factor& F = cfg->gp.factors[f];
        SLB_CHECK(F.M == 0 || (F.Xf != nullptr && F.Whead != nullptr && F.Wheadp != nullptr &&
                               F.Xhead != nullptr),
                  "filtered sweep: GP factor %d lacks the filter tables (Xf / Whead / Wheadp / Xhead)", f);
        SLB_CHECK(F.M == 0 || ((reinterpret_cast<uintptr_t>(F.Wheadp) & 15) == 0 &&
                               (reinterpret_cast<uintptr_t>(F.Xhead) & 15) == 0),
                  "filtered sweep: GP factor %d: Wheadp / Xhead must be 16-byte aligned", f);
        SLB_CHECK(F.head_rows >= 0 && F.head_rows <= SLB_HEAD_RANK && F.head_rows <= F.M,
                  "filtered sweep: GP factor %d has %d head rows (0..min(M, %d))", f, F.head_rows,
                  SLB_HEAD_RANK);
        SLB_CHECK((reinterpret_cast<uintptr_t>(F.Xf) & 15) == 0,
                  "filtered sweep: GP factor %d: Xf must be 16-byte aligned", f);
        int no = 0;
        for (int o = 0; o < cfg->gp.num_outputs; ++o) no += cfg->gp.outputs[o].factor == f;
        if (no > nomax) nomax = no;
    }
    for (int o = 0; o < cfg->gp.num_outputs; ++o) {
        const slb_gp_output& G = cfg->gp.outputs[o];
        SLB_CHECK(cfg->gp.factors[G.factor].M == 0 ||
                  (G.gamma_f != nullptr && (reinterpret_cast<uintptr_t>(G.gamma_f) & 15) == 0),
                  "filtered sweep: GP output %d has no (16-byte aligned) gamma_f", o);
        SLB_CHECK(G.gamma_l1 >= 0.0, "filtered sweep: GP output %d has no gamma_l1", o);
    }
    cudaStream_t st = (cudaStream_t)stream;
    char* ws = static_cast<char*>(workspace_dev);
    const int64_t cap = n_all < CHUNK ? n_all : CHUNK;
    filter_args a;
    memset(&a, 0, sizeof(a));
    a.counts = reinterpret_cast<unsigned long long*>(ws);
    int* tickets = reinterpret_cast<int*>(ws + 64);
    double* partial = reinterpret_cast<double*>(ws + 64 + SLB_SPLIT_TICKET_BYTES);
    a.list_a = reinterpret_cast<int64_t*>(ws + WS_HEAD);
    a.list_b = a.list_a + cap;
    a.side_a = reinterpret_cast<filter_side*>(a.list_b + cap);
    a.stats = reinterpret_cast<unsigned long long*>(stats_dev);
    // rows per staged slice: two buffers of (d_in + 1 + outputs per factor) doubles per row within
    // ~24 KB, so that 7 CTAs stay resident per SM
    const int din = cfg->gp.input_dim;
#ifndef SLB_MEAN_SMEM_KB
#define SLB_MEAN_SMEM_KB 24
#endif
    a.chunk_rows = mean_chunk_rows(din, nomax, SLB_MEAN_SMEM_KB);
    a.max_outputs_per_factor = nomax;
    const size_t smem = mean_smem_bytes(din, nomax, a.chunk_rows);
    for (int64_t off = 0; off < n_all; off += CHUNK) {
        const int64_t n = n_all - off < CHUNK ? n_all - off : CHUNK;
        SLB_CUDA(cudaMemsetAsync(a.counts, 0, 64 + SLB_SPLIT_TICKET_BYTES, st));
        a.n = n; a.idx_begin = idx_begin + off;
        a.negative = negative_dev + off;
        a.values = values_dev ? values_dev + off : nullptr;
        int rc;
        switch (din) {
        case 1: rc = launch_filter<1>(st, *cfg, a, smem); break;
        case 2: rc = launch_filter<2>(st, *cfg, a, smem); break;
        case 3: rc = launch_filter<3>(st, *cfg, a, smem); break;
        case 4: rc = launch_filter<4>(st, *cfg, a, smem); break;
        case 5: rc = launch_filter<5>(st, *cfg, a, smem); break;
        case 6: rc = launch_filter<6>(st, *cfg, a, smem); break;
        default:
            slb_set_error("GP input_dim %d not compiled (1..6)", din);
            return 1;
        }
        if (rc) return rc;
        if (!(g_filter_stages & 2)) continue;
        rc = slb_launch_refine(st, *cfg, n, idx_begin + off, a.list_b, a.counts + 1,
                               negative_dev + off, values_dev ? values_dev + off : nullptr,
                               partial, tickets);
        if (rc) return rc;
    }
    return 0;
}

}  // extern "C"
