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
struct filter_side { double dec0, thr, guard, coef[SLB_MAX_OUT], z[SLB_MAX_IN]; };

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

template <int DIN>
__global__ void __launch_bounds__(FT, SLB_MEAN_MINB)
filter_mean_kernel(const __grid_constant__ slb_sweep cfg, const filter_args a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
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

    // ---- V(mu), L_V(mu) and the coefficient of every sigma_j          (lyapunov.py:344-352)
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

// one group of P list entries [g P, g P + P) on one warp
template <int DIN, int P, bool ALL_STAGED>
SLB_DEV void head_group(const slb_sweep& cfg, const filter_args& a, int64_t grp, int64_t count,
                        const double* exptab, double* kw, const double* wbuf, const double* xbuf,
                        unsigned* s_stat) {
    const int lane = threadIdx.x & 31;
    const int nf = cfg.gp.num_factors;
    const int D = cfg.gp.num_outputs;
    // lane p < P owns list entry grp * P + p: its terms, its index and finally its decision
    const int64_t k = grp * P + min(lane, P - 1);
    const bool mine = lane < P && k < count;
    filter_side t = {};
    int64_t rel = 0;
    if (mine) { t = a.side_a[k]; rel = a.list_a[k]; }
    double shi[SLB_MAX_OUT];
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
#pragma unroll
        for (int p = 0; p < P; ++p) {
            double zs[DIN];
#pragma unroll
            for (int c = 0; c < DIN; ++c) {
                const double zc = __shfl_sync(0xffffffffu, t.z[c], p);
                zs[c] = general ? zc : zc / F.lengthscales[c];
            }
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
    const int outcome = mine ? decide(t, shi, D) : 0;
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
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);             // [1]
    unsigned* s_stat = reinterpret_cast<unsigned*>(smem_raw + 8);      // decided / undecided by this CTA
    double* exptab = reinterpret_cast<double*>(smem_raw + 16);         // [64]
    double* kbuf = exptab + 64;                                        // [HW][HR][HP]
    double* wbuf = kbuf + HW * HR * HP;                                // [staged][HR * HR]
    const int nf = cfg.gp.num_factors;
    double* xbuf = wbuf + (size_t)a.head_factors_staged * HR * HR;     // [staged][HR * DIN]
    const int64_t count = (int64_t)a.counts[0];
    const int64_t nwarps = (int64_t)gridDim.x * HW;
    const bool short_list = false;     // every list takes 8-point DMMA groups (the 2-point FMA form of
                                       // head_group is kept for reference / A-B timing)
    const int64_t ngroups = (count + HP - 1) / HP;
    (void)nwarps;
    if ((int64_t)blockIdx.x * HW >= ngroups) return;                   // no group for this CTA
    if (threadIdx.x == 0) {
        slb_bulk::mbar_init(bar, 1);
        slb_bulk::fence_barrier_init();
        slb_bulk::fence_proxy_async();
        unsigned bytes = 64 * sizeof(double);
        for (int f = 0; f < a.head_factors_staged; ++f)
            if (cfg.gp.factors[f].head_rows > 0)
                bytes += (unsigned)(HR * HR + HR * DIN) * sizeof(double);
        slb_bulk::mbar_arrive_expect_tx(bar, bytes);
        slb_bulk::copy_g2s(exptab, g_exp_tables + 512, 64 * sizeof(double), bar);
        for (int f = 0; f < a.head_factors_staged; ++f) {
            const slb_gp_factor& F = cfg.gp.factors[f];
            if (F.head_rows <= 0) continue;
            slb_bulk::copy_g2s(wbuf + (size_t)f * HR * HR, F.Wheadp, HR * HR * sizeof(double), bar);
            slb_bulk::copy_g2s(xbuf + (size_t)f * HR * DIN, F.Xhead, HR * DIN * sizeof(double), bar);
        }
    }
    if (threadIdx.x < 2) s_stat[threadIdx.x] = 0;
    __syncthreads();
    slb_bulk::mbar_wait(bar, 0);
    const int warp = threadIdx.x >> 5;
    double* kw = kbuf + warp * HR * HP;
    for (int64_t grp = (int64_t)blockIdx.x * HW + warp; grp < ngroups; grp += nwarps) {
        if (a.head_factors_staged == nf) {
            if (short_list) head_group<DIN, HP_SHORT, true>(cfg, a, grp, count, exptab, kw, wbuf, xbuf, s_stat);
            else head_group<DIN, HP, true>(cfg, a, grp, count, exptab, kw, wbuf, xbuf, s_stat);
        } else {
            if (short_list) head_group<DIN, HP_SHORT, false>(cfg, a, grp, count, exptab, kw, wbuf, xbuf, s_stat);
            else head_group<DIN, HP, false>(cfg, a, grp, count, exptab, kw, wbuf, xbuf, s_stat);
        }
    }
    // one pair of global atomics per CTA (one per point serialised on the counter's L2 line)
    __syncthreads();
    if (a.stats != nullptr && threadIdx.x < 2 && s_stat[threadIdx.x] != 0)
        atomicAdd(a.stats + 1 + threadIdx.x, (unsigned long long)s_stat[threadIdx.x]);
}

int g_filter_stages = 3;               // slb_debug_filter_stages: bit 0 head stage, bit 1 refine pass

template <int DIN>
int launch_filter(cudaStream_t st, const slb_sweep& cfg, const filter_args& a, size_t smem) {
    static std::atomic<bool> configured[64];
    int device = 0;
    SLB_CUDA(cudaGetDevice(&device));
    if (device < 0 || device >= 64 || !configured[device].load(std::memory_order_acquire)) {
        SLB_CUDA(cudaFuncSetAttribute(filter_mean_kernel<DIN>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
        SLB_CUDA(cudaFuncSetAttribute(filter_head_kernel<DIN>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        if (device >= 0 && device < 64) configured[device].store(true, std::memory_order_release);
    }
    const int64_t blocks = (a.n + FT - 1) / FT;
    filter_mean_kernel<DIN><<<(unsigned)blocks, FT, smem, st>>>(cfg, a);
    SLB_LAUNCH_CHECK();
    if (!(g_filter_stages & 1)) return 0;
    const size_t head_fixed = 16 + (64 + HW * HR * HP) * sizeof(double);
    const size_t per_factor = (size_t)(HR * HR + HR * DIN) * sizeof(double);
    filter_args ah = a;
    ah.head_factors_staged = cfg.gp.num_factors;
    while (ah.head_factors_staged > 0 && head_fixed + ah.head_factors_staged * per_factor > 226 * 1024)
        --ah.head_factors_staged;
    const size_t head_smem = head_fixed + ah.head_factors_staged * per_factor;
    filter_head_kernel<DIN><<<HEAD_CTAS, HT, head_smem, st>>>(cfg, ah);
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
