/*
 * slb200.h -- C ABI of libslb200.so: the B200 (sm_100a) implementation of the
 * safe_learning region-of-attraction hot path.
 *
 * The reference (befelix/safe_learning @ f1aad5a) has no FFI: the path sits behind
 * Python classes that build a TF1 graph and call Session.run once per 10 000-point
 * batch.  The entry points below are what a binding for that path would bind; each
 * cites the reference code it replaces (paths relative to /root/reference).  The
 * Python host side (safe_learning_b200/) loads this library with ctypes -- see
 * INTEGRATION.md for the stub a maintainer of the reference would add.
 *
 * Conventions
 *   - plain C: pointers, sizes, POD structs passed by pointer; no torch/C++ types.
 *   - every `*_dev` / `const double*` inside a struct is a DEVICE pointer (fp64,
 *     row-major, contiguous) owned by the caller; the library never allocates or frees.
 *   - `stream` is a cudaStream_t (CUstream) cast to void*; calls enqueue and return.
 *   - return 0 on success, non-zero on error; slb_last_error() gives the message.
 *     There is NO CPU fallback: without a CUDA device every compute call fails.
 *   - fp64 everywhere (safe_learning/configuration.py:16); flags are uint8; flat grid
 *     indices int64 with the last dimension fastest (functions.py:622-638).
 */
#ifndef SLB200_H
#define SLB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SLB_ABI_VERSION 5   /* 2: slb_gp_factor.kernel (covariance expressions), GRADIENT / MAXABS flags
                               3: decision filter (slb_gp_factor.Whead, slb_lyapunov_sweep_filtered),
                                  state-dependent lipschitz_dynamics, peer-memory key exchange
                                  (slb_exchange), fixed-action Bellman tables
                               4: filter tables staged by TMA bulk copies (slb_gp_factor.Xf / Xhead /
                                  head_rows / hmax, slb_gp_output.gamma_f / gamma_l1): pivoted head
                                  subset, computed error bound of the filter's mean
                               5: fp32 screening stage in front of the filter (slb_debug_screening_probe,
                                  bit 2 of slb_debug_filter_stages); no structure changed           */
#define SLB_MAX_DIM 6   /* state dimension d                         */
#define SLB_MAX_IN  8   /* GP input dimension d_in = d + m           */
#define SLB_MAX_OUT 6   /* stacked one-output GPs (FunctionStack)    */
#define SLB_MAX_ACT 2   /* action dimension m                        */
#define SLB_TILE_POINTS 64  /* grid points per CTA tile of the GP kernels */
#define SLB_HEAD_RANK 64    /* size of the training subset behind the decision filter's variance bound */
#define SLB_MAX_RANKS 16    /* ranks of one peer-memory key exchange (one NVSwitch domain)       */

/* ---- GridWorld (functions.py:579-817) ------------------------------------------- */
typedef struct slb_grid {
    int32_t ndim;
    int32_t _pad;
    int64_t nindex;                    /* prod(num_points)                              */
    int64_t num_points[SLB_MAX_DIM];
    double  offset[SLB_MAX_DIM];       /* limits[:, 0]                  (:604)          */
    double  unit_maxes[SLB_MAX_DIM];   /* (hi - lo) / (n - 1)           (:605-606)      */
    double  upper[SLB_MAX_DIM];        /* limits[:, 1]                                  */
    const double* discrete_points;     /* device, concatenated np.linspace values per dim
                                          (:612-614); needed only by SLB_FN_TRIANGULATION */
} slb_grid;

/* ---- function objects fused into the kernels (functions.py, examples/utilities.py) -- */
enum slb_fn_kind {
    SLB_FN_NONE = 0,
    SLB_FN_CONSTANT = 1,       /* ConstantFunction            functions.py:241-251          */
    SLB_FN_LINEAR = 2,         /* LinearSystem  y = x A^T      functions.py:1546-1583        */
    SLB_FN_QUADRATIC = 3,      /* QuadraticFunction sum((xP)*x) functions.py:1513-1539       */
    SLB_FN_TRIANGULATION = 4,  /* Triangulation               functions.py:1103-1158,1442-1499 */
    SLB_FN_PENDULUM = 5,       /* InvertedPendulum            examples/utilities.py:144-289 */
    SLB_FN_CARTPOLE = 6,       /* CartPole                    examples/utilities.py:292-437 */
    SLB_FN_LYAPUNOV_NN = 7,    /* LyapunovNetwork             examples/utilities.py:48-104  */
    SLB_FN_MLP = 8             /* NeuralNetwork (inference)   functions.py:1702-1729        */
};
/* post-ops, applied in this order: saturate -> abs -> norm1 | maxabs -> out_scale */
#define SLB_FLAG_SATURATE 1u   /* Saturation  functions.py:349-354                     */
#define SLB_FLAG_ABS      2u   /* tf.abs(fun(x))     (notebook Lipschitz lambdas)      */
#define SLB_FLAG_NORM1    4u   /* tf.norm(., ord=1, axis=1, keepdims=True)             */
#define SLB_FLAG_PROJECT  8u   /* Triangulation(project=True) functions.py:1479-1485  */
#define SLB_FLAG_SCALE   16u   /* multiply by out_scale (MultipliedFunction / __neg__) */
#define SLB_FLAG_GRADIENT 32u  /* TRIANGULATION with one output column: return the d partial
                                  derivatives of the piecewise-linear interpolant instead of its
                                  value (Triangulation.gradient, functions.py:1260-1326, 1506-1510);
                                  out_dim = d                                              */
#define SLB_FLAG_MAXABS  64u   /* tf.reduce_max(tf.abs(.), axis=1, keepdims=True): the
                                  Lipschitz lambda of examples/inverted_pendulum.ipynb cell 14;
                                  applied after saturate, reduces to 1 column              */

typedef struct slb_function {
    int32_t kind;
    int32_t in_dim;
    int32_t out_dim;            /* before NORM1 / MAXABS (which reduce to 1 column)    */
    uint32_t flags;
    double  out_scale;
    double  lower, upper;       /* saturation bounds                                   */
    double  cparams[24];        /* CONSTANT: value; PENDULUM/CARTPOLE: plant constants
                                   (see safe_learning_b200/functions.py); LYAPUNOV_NN / MLP:
                                   [0] layers, [1+i] width of layer i (<= 64), [9+i] activation
                                   (0 tanh, 1 relu, 2 identity); MLP: [17] output_scale, [18] 1 if
                                   hidden layers carry a bias (matrix = [W_i (out x in), b_i] ...) */
    const double*  matrix;      /* LINEAR [out,in]; QUADRATIC [in,in]; TRIANGULATION
                                   vertex values [nindex,out]; LYAPUNOV_NN packed kernels */
    const double*  hyperplanes; /* TRIANGULATION [nsimplex, d, d]  (functions.py:1090-1101) */
    const int64_t* unit_simplices; /* TRIANGULATION [nsimplex, d+1] (functions.py:1064-1088) */
    const int32_t* corner_simplex; /* TRIANGULATION [2^d] or NULL: Qhull's find_simplex answer for
                                      a query clipped in EVERY dimension (bit c set = clipped to
                                      the upper limit); such points sit on a unit-cell corner
                                      where several simplices meet and extrapolation differs  */
    int32_t nsimplex;
    int32_t _pad;
    slb_grid grid;              /* TRIANGULATION discretization                        */
} slb_function;

/* ---- covariance functions ------------------------------------------------------------------
 * The reference hands any gpflow kernel to GPRCached (functions.py:370-393) and only ever calls
 * kern.K(X), kern.K(X, Xnew) and kern.Kdiag(Xnew) on it (functions.py:399, 438, 450).  Its
 * experiments use sums of products of gpflow==0.4.0 primitives with `active_dims`
 * (examples/inverted_pendulum.ipynb cell 6, adaptive_safety_verification.ipynb cell 9,
 * 1d_region_of_attraction_estimate.ipynb cell 5):
 *     k(x, x') = sum over terms t of  prod over primitives p with p.term == t of  k_p(x, x')
 * gpflow 0.4.0 kernels.py arithmetic of the primitives (r^2 = sum_c ((x_c - x'_c) w_c)^2 with
 * w_c = 1 / lengthscale_c on active dimensions and 0 elsewhere; r = sqrt(r^2 + 1e-12)):          */
enum slb_kernel_kind {
    SLB_K_RBF = 0,        /* variance exp(-r^2 / 2)                                          */
    SLB_K_MATERN12 = 1,   /* variance exp(-r)                                                */
    SLB_K_MATERN32 = 2,   /* variance (1 + sqrt3 r) exp(-sqrt3 r)                            */
    SLB_K_MATERN52 = 3,   /* variance (1 + sqrt5 r + 5/3 r^2) exp(-sqrt5 r)                  */
    SLB_K_LINEAR = 4,     /* sum_c w_c x_c x'_c, w_c = variance_c on active dims (ARD or not) */
    SLB_K_CONSTANT = 5,   /* variance (gpflow Constant / Bias)                               */
    SLB_K_WHITE = 6       /* variance on the diagonal (Kdiag, K(X)); 0 against new points    */
};
#define SLB_MAX_KPRIM 6

typedef struct slb_kernel_prim {
    int32_t kind;               /* slb_kernel_kind                                      */
    int32_t term;               /* product term; primitives are listed in term order    */
    double  variance;           /* stationary / constant / white primitives             */
    double  w[SLB_MAX_IN];      /* per input column, see above; 0 = inactive dimension  */
} slb_kernel_prim;

typedef struct slb_kernel {
    int32_t num_prims;          /* 0 => the factor is the plain full-dimensional RBF described
                                   by slb_gp_factor.lengthscales / variance / kss (fast path)  */
    int32_t _pad;
    slb_kernel_prim prims[SLB_MAX_KPRIM];
} slb_kernel;

/* ---- GP stack: FunctionStack of GaussianProcess(GPRCached) (functions.py:254-546) ---- */
typedef struct slb_gp_factor {
    int32_t M;                  /* training points; 0 => prior only (empty data set of the
                                   notebooks before the first add_data_point)           */
    int32_t nrb;                /* ceil(M / 8) row blocks of the packed factor          */
    const double* Xs;           /* device [M, d_in]: X / lengthscales when kernel.num_prims == 0
                                   (gpflow RBF), the raw X otherwise                    */
    const double* Wpack;        /* device: L^-1 in DMMA fragment order, k-steps paired
                                   (slb_pack_factor); L = chol(scale^2 (K + noise I))
                                   functions.py:399-408 */
    double lengthscales[SLB_MAX_IN];
    double variance;            /* RBF variance (Kdiag)                                 */
    double scale;               /* GPRCached _scale                functions.py:392     */
    double kss;                 /* (scale**2) * variance           functions.py:450     */
    slb_kernel kernel;          /* general covariance expression (num_prims > 0)        */
    /* ---- tables of the decision filter (slb_lyapunov_sweep_filtered); may be NULL / 0 if that
       call is not used.  The posterior variance given ANY subset S of the training set is an upper
       bound of the full posterior variance; the host picks up to SLB_HEAD_RANK points in pivoted-
       Cholesky (greedy max-variance) order and factors them on their own:                        */
    const double* Whead;        /* device [SLB_HEAD_RANK, SLB_HEAD_RANK], COLUMN-major, zero padded:
                                   Whead[j * SLB_HEAD_RANK + i] = L_S^-1[i, j], L_S = chol(scale^2
                                   (K(X_S) + noise I))                                   */
    const double* Wheadp;       /* device, 16-byte aligned: the same L_S^-1 (zero padded to SLB_HEAD_RANK^2) in
                                   DMMA.8x8x4 A-fragment order: for row block b (8 rows) and k-step s
                                   (4 columns) 32 doubles, lane T <-> L_S^-1[8b + T/4, 4s + T%4], block
                                   offset (b * (SLB_HEAD_RANK / 4) + s) * 32                */
    const double* Xhead;        /* device [SLB_HEAD_RANK, d_in], zero padded: the subset's inputs,
                                   scaled like Xs                                        */
    int32_t head_rows;          /* |S| = min(M, SLB_HEAD_RANK)                           */
    int32_t _pad2;
    const double* Xf;           /* device [8 ceil(M/8), w], zero padded, 16-byte aligned (TMA bulk
                                   copies): kernel.num_prims == 0: w = d_in + 1, row j =
                                   (Xs[j, :], -|Xs[j, :]|^2 / 2); otherwise w = d_in, the raw X  */
    double hmax;                /* max_j |Xs[j, :]|^2 / 2 (rounding bound of the expanded distance) */
} slb_gp_factor;

typedef struct slb_gp_output {
    int32_t factor;             /* index into factors[] (outputs sharing X, kernel and
                                   noise share one factor)                              */
    int32_t _pad;
    double  beta;               /* GaussianProcess.beta            functions.py:487,514 */
    const double* alpha;        /* device [8*nrb]: L^-1 scale (Y - m(X)), zero padded
                                                                    functions.py:405-409 */
    const double* gamma;        /* device [M]: L^-T alpha (mean-only Bellman path)      */
    const double* prior_mean;   /* device [d_in] linear prior-mean row, or NULL         */
    const double* gamma_f;      /* device [8 ceil(M/8)], zero padded, 16-byte aligned: the filter's
                                   mean weights, scale^2 gamma (times the RBF variance when
                                   kernel.num_prims == 0, whose kernel values are then <= 1)   */
    double gamma_l1;            /* >= sum_i (|L^-1|^T |alpha|)_i in the units of gamma_f: bounds the
                                   rounding error of every way of summing the mean (filter)    */
} slb_gp_output;

typedef struct slb_gp_stack {
    int32_t num_outputs;        /* D; 0 => dynamics are deterministic                   */
    int32_t num_factors;        /* D' distinct Cholesky factors                         */
    int32_t input_dim;          /* d_in                                                 */
    int32_t _pad;
    slb_gp_factor factors[SLB_MAX_OUT];
    slb_gp_output outputs[SLB_MAX_OUT];
} slb_gp_stack;

/* ---- one Lyapunov sweep: the graph of lyapunov.py:433-441 ----------------------------- */
typedef struct slb_sweep {
    slb_grid     grid;          /* discretization                                       */
    slb_function policy;        /* x -> u                       lyapunov.py:436         */
    slb_function dynamics;      /* deterministic [x,u] -> x+ when gp.num_outputs == 0   */
    slb_gp_stack gp;            /* uncertain dynamics           lyapunov.py:437         */
    slb_function lyapunov;      /* V                            lyapunov.py:351-352     */
    slb_function lipschitz_v;   /* L_V(.) as a function, or kind NONE => lv_const       */
    double lv_const;            /* scalar lipschitz_lyapunov    lyapunov.py:246-263     */
    double lf_const;            /* scalar lipschitz_dynamics    lyapunov.py:227-244     */
    double tau;                 /* discretization constant      lyapunov.py:195         */
    slb_function lipschitz_f;   /* state-dependent L_f(x) as a fused function (first column), or
                                   kind NONE                    lyapunov.py:227-244, 287 */
    const double* lf_values;    /* device, or NULL: L_f tabulated per flat grid index (an arbitrary
                                   Python callable evaluated once by the host); entry for grid
                                   index i is lf_values[i - lf_index_base].  Index-range sweeps only.
                                   Precedence: lf_values, lipschitz_f, lf_const            */
    int64_t lf_index_base;
} slb_sweep;

/* ---- one Bellman sweep: PolicyIteration.future_values (reinforcement_learning.py:65-114) */
typedef struct slb_bellman {
    slb_grid     grid;          /* value_function.discretization (state space, :58-59)  */
    slb_function policy;        /* ignored when fixed_action != 0                       */
    slb_function dynamics;      /* deterministic dynamics when gp.num_outputs == 0      */
    slb_gp_stack gp;            /* GP dynamics: mean only               (:98-99)        */
    slb_function reward;        /* r([x,u])                             (:95)           */
    slb_function value;         /* V as Triangulation (vertex table = `matrix`) (:101)  */
    double gamma;               /*                                      (:104)          */
    int32_t fixed_action;       /* 1 => use `action` for every state (:266-270)         */
    int32_t _pad;
    double action[SLB_MAX_ACT];
} slb_bellman;


/* ---- result of the first-fail reduction (sort-free form of lyapunov.py:512-587) ------- */
typedef struct slb_fail_key {
    uint64_t key_value;   /* order-preserving bits of V at the first failing point in stable
                             V-order, or UINT64_MAX if no point fails                     */
    int64_t  key_index;   /* its flat grid index, or INT64_MAX                            */
    int64_t  n_ok;        /* number of points with negative | initial                     */
    int64_t  _pad;
} slb_fail_key;

typedef struct slb_prefix_stats {
    int64_t  n_safe;        /* |{key < k*} U initial|                                     */
    int64_t  n_below;       /* |{key < k*}| = sorted position p* of the first failure     */
    uint64_t max_below;     /* order-preserving bits of max{V_i : key_i < k*} (0 if none) */
    uint64_t max_all;       /* order-preserving bits of max V over the range              */
} slb_prefix_stats;

/* ---- per-sweep key exchange between the ranks of one NVLink domain without a collective call:
 *      every rank owns `slots` = slb_fail_key[2][world] in peer-mapped (symmetric) memory; after its
 *      first-fail reduction a rank STORES its key into slot [parity][rank] of every peer (P2P
 *      writes through NVLink) with the sweep's sequence number as the release flag, and the prefix
 *      kernel of every rank waits for the `world` flags of the current sweep in its own copy.
 *      Ranks must issue the same sequence of sweeps (like any collective).                        */
typedef struct slb_exchange {
    int32_t world;
    int32_t rank;
    slb_fail_key* slots[SLB_MAX_RANKS];  /* slots[r]: rank r's slot array as mapped in THIS process */
    int64_t* seq_dev;                    /* local device int64 (zero-initialised): sweeps issued  */
} slb_exchange;

/* ---- library ---------------------------------------------------------------------------- */
int         slb_abi_version(void);
const char* slb_last_error(void);
/* number of CUDA devices visible, or <0 with slb_last_error set */
int         slb_device_count(void);
/* sizeof() of the ABI structs in declaration order (grid, function, gp_factor, gp_output,
   gp_stack, sweep, bellman, fail_key, prefix_stats, exchange); returns how many there are */
int         slb_struct_sizes(int64_t* out, int32_t n);
/* kernels this library has launched since load (bench.py's gpu_launches) */
int64_t     slb_launch_count(void);
/* a caller that replays a CUDA graph captured around `kernels` launches of this library reports
   each replay here so that slb_launch_count stays truthful */
void        slb_note_graph_replay(int64_t kernels);

/* diagnostics: when buffer_dev != NULL, every later Lyapunov sweep writes int64 cycle counts
 * [tile][warp(8)][8] = {k-row generation, DMMA contraction, panel epilogues, whole tile}, the
 * %globaltimer (ns) at tile start / end, cycles spent waiting at block barriers, 0;
 * pass NULL to switch it off (default) */
int         slb_debug_phase_timing(void* buffer_dev);
/* Records an event (owned by the library, one per device) on `stream` that every later launch reading
 * the packed factors (slb_gp_factor.Wpack: the full posterior of slb_gp_predict / slb_lyapunov_sweep /
 * slb_lyapunov_points and the refine pass of slb_lyapunov_sweep_filtered) waits for on ITS stream.  A
 * caller that restores the GP tables from the host copies the packed factors -- 90% of the bytes -- on a
 * second stream, calls this behind the copy, and enqueues the sweep at once: the filter stages do not
 * read the packed factors and overlap the copy.  Under stream capture the wait is an external-event
 * node (re-evaluated at every replay).  Single-threaded use like the other setters. */
int         slb_record_factor_dependency(void* stream);
/* Restore of a packed table arena (safe_learning_b200.functions.PackedCache: every cached GP table of a
 * stack in one device buffer, mirrored by one page-locked host buffer, the packed factors last) in one
 * call: bytes [0, split_bytes) are copied host -> device on `stream`, bytes [split_bytes, total_bytes)
 * -- the packed factors -- on `side_stream` behind everything enqueued on `stream` so far, followed by
 * slb_record_factor_dependency(side_stream).  side_stream NULL (or == stream): one stream, no event. */
int         slb_restore_tables(void* dst_dev, const void* src_host, int64_t split_bytes, int64_t total_bytes,
                               void* stream, void* side_stream);
/* diagnostics (timing of the individual stages of slb_lyapunov_sweep_filtered; the flags are only
 * complete with bits 0 and 1 set -- 3, the default, or 7): bit 0 runs the head stage, bit 1 the
 * refine pass; bit 2 forces the fp64 mean stage where the fp32 screening stage would run */
int         slb_debug_filter_stages(int32_t mask);
/* diagnostics: while both pointers are non-NULL, the fp32 screening stage of
 * slb_lyapunov_sweep_filtered also writes, for every point of the swept range (row = index relative to
 * idx_begin of the LAST pass), its screened mean [n, D] and the certified bound of its error [n, D]
 * (+inf where the point is left to the fp64 stages); tests hold |mean - fp64 mean| <= bound */
int         slb_debug_screening_probe(double* mu_dev, double* dm_dev);
/* diagnostics: deterministic-dynamics sweeps of the LQR composition (d = 2, saturated linear policy,
 * linear dynamics, quadratic V, constant / abs-linear L_V) take a specialised register-resident
 * kernel; 0 switches back to the generic interpreter (A/B timing, parity tests). */
int         slb_debug_det_fast(int32_t enable);
/* diagnostics / tuning: the refine pass of slb_lyapunov_sweep_filtered uses 16-point tiles for
 * lists of up to `upto16` points, 32-point tiles up to `upto32`, 64-point tiles beyond (defaults
 * 0 -- no 16-point launch -- and 32 * 148); with 16- and 32-point tiles the rows of a tile are additionally split
 * over the CTAs a one-per-SM grid has to spare (up to 8 per tile) */
int         slb_debug_refine_split(int64_t upto16, int64_t upto32);

/* ---- GP factor packing (after GPRCached.update_cache, functions.py:395-415) ------------ */
/* doubles needed for the packed L^-1 of an M-point GP */
int64_t slb_packed_len(int32_t M);
/* Linv_dev [M,M] row-major lower-triangular -> Wpack_dev (slb_packed_len(M) doubles) */
int slb_pack_factor(void* stream, const double* Linv_dev, int32_t M, double* Wpack_dev);

/* head subset of the decision filter: the first r <= min(M, SLB_HEAD_RANK) pivots of the pivoted
 * Cholesky factorisation of the symmetric positive semi-definite kernel_dev [M, M] (greedy: the
 * training point with the largest variance given those already chosen; ties: lowest index)
 * -> picks_dev [r] (int64).  scratch_dev: M * (r + 1) doubles. */
int slb_pivoted_subset(void* stream, const double* kernel_dev, int32_t M, int32_t r,
                       int64_t* picks_dev, double* scratch_dev);

/* ---- GP posterior on an explicit point list: GaussianProcess.__call__ / FunctionStack
 *      (functions.py:278-291, 417-458, 507-515).  points_dev [n, d_in] (already [x,u]
 *      concatenated, utilities.py:143); mean_dev, err_dev [n, D];
 *      err = beta*sqrt(var) (want_var == 0) or the latent variance (want_var != 0). ------- */
int slb_gp_predict(void* stream, const slb_gp_stack* gp, const double* points_dev, int64_t n,
                   double* mean_dev, double* err_dev, int32_t want_var);

/* ---- the fused Lyapunov sweep over flat grid indices [idx_begin, idx_end):
 *      index -> x (functions.py:714-731) -> u = policy(x) -> [x,u] -> GP mean / beta*sigma
 *      (or deterministic dynamics) -> V(x), V(mu), L_V(mu) . e -> decrease < threshold
 *      (lyapunov.py:265-288, 324-376, 436-441).  Outputs are arrays of length
 *      idx_end - idx_begin; any of values/decrease/threshold/mean/err may be NULL. -------- */
int slb_lyapunov_sweep(void* stream, const slb_sweep* cfg, int64_t idx_begin, int64_t idx_end,
                       uint8_t* negative_dev, double* values_dev, double* decrease_dev,
                       double* threshold_dev, double* mean_dev, double* err_dev);
/* The same decision flags (and V) with a certified filter in front of the O(M^2) posterior:
 * a thread-per-point kernel computes the GP mean (k . L^-T alpha, with a computed error bound),
 * V(mu), L_V(mu) and an UPPER bound of every output's standard deviation -- first the prior's,
 * then (one warp per remaining point) the posterior given only a head subset of at most
 * SLB_HEAD_RANK training points (factor.Whead / Xhead) -- and decides every point whose
 * comparison `decrease < threshold` has the same outcome for all sigma in [0, bound] (guard band:
 * 1e-6 relative + the mean's error bound); the remaining points are compacted and go through the
 * full fp64 posterior (the kernel of slb_lyapunov_sweep).  Flags are identical to
 * slb_lyapunov_sweep's.
 * workspace_dev: >= slb_filter_workspace(n) bytes.  stats_dev: NULL or 4 int64 (device):
 * {decided by mean + prior bound, decided by the head-rank bound, refined by the full posterior,
 * points} accumulated over calls (the caller zeroes it). */
int64_t slb_filter_workspace(int64_t n);
/* which first stage slb_lyapunov_sweep_filtered runs for this configuration: 32 = the fp32 screening
 * kernel (plain RBF factors, quadratic V on at most four outputs, constant / abs-linear L_V, tables fit
 * the head stage's shared memory) with an fp64 mean only for the points its error box leaves open;
 * 64 = the fp64 mean kernel; 0 = no GP */
int slb_filter_stage1(const slb_sweep* cfg);
int slb_lyapunov_sweep_filtered(void* stream, const slb_sweep* cfg, int64_t idx_begin,
                                int64_t idx_end, uint8_t* negative_dev, double* values_dev,
                                void* workspace_dev, int64_t* stats_dev);
/* same on an explicit state list states_dev [n, d] (get_safe_sample-style callers) */
int slb_lyapunov_points(void* stream, const slb_sweep* cfg, const double* states_dev, int64_t n,
                        uint8_t* negative_dev, double* values_dev, double* decrease_dev,
                        double* threshold_dev, double* mean_dev, double* err_dev);

/* ---- prefix rule without sorting (lyapunov.py:500-606):
 *      k* = min over failing points of key (V_i, i); safe_i = key_i < k* | initial_i.
 *      workspace_dev: >= slb_first_fail_workspace(n) bytes.  result_dev: one slb_fail_key
 *      in device memory (the caller all-reduces it across ranks, then applies).           */
int64_t slb_first_fail_workspace(int64_t n);
int slb_first_fail(void* stream, const double* values_dev, const uint8_t* negative_dev,
                   const uint8_t* initial_dev /* may be NULL */, int64_t n, int64_t idx_begin,
                   void* workspace_dev, slb_fail_key* result_dev);
/* sharded form with the exchange fused into the two kernels: slb_first_fail_x pushes this rank's
 * key to every peer, slb_apply_prefix_x waits for all keys of the sweep, reduces them (writing the
 * winner to key_out_dev) and applies the prefix rule -- no collective call, no host round trip */
int slb_first_fail_x(void* stream, const double* values_dev, const uint8_t* negative_dev,
                     const uint8_t* initial_dev, int64_t n, int64_t idx_begin, void* workspace_dev,
                     slb_fail_key* result_dev, const slb_exchange* xchg);
int slb_apply_prefix_x(void* stream, const double* values_dev, const uint8_t* initial_dev,
                       int64_t n, int64_t idx_begin, slb_fail_key* key_out_dev, uint8_t* safe_dev,
                       void* workspace_dev, slb_prefix_stats* stats_dev, const slb_exchange* xchg);
/* multi-GPU: lexicographic min (and n_ok sum) over `world` keys all-gathered by the caller
 * (the one collective of a sweep) -> out_dev; may alias gathered_dev[0] */
int slb_combine_fail_keys(void* stream, const slb_fail_key* gathered_dev, int32_t world,
                          slb_fail_key* out_dev);
int slb_apply_prefix(void* stream, const double* values_dev, const uint8_t* initial_dev,
                     int64_t n, int64_t idx_begin, const slb_fail_key* key_dev,
                     uint8_t* safe_dev, void* workspace_dev, slb_prefix_stats* stats_dev);

/* ---- generic evaluation of a fused function object on explicit points:
 *      DeterministicFunction.__call__, Lyapunov.update_values (lyapunov.py:305-322).
 *      points_dev [n, fn->in_dim] -> out_dev [n, out columns]. ---------------------------- */
int slb_eval_function(void* stream, const slb_function* fn, const double* points_dev, int64_t n,
                      double* out_dev);
/* grid coordinates for flat indices [idx_begin, idx_end): GridWorld.index_to_state */
int slb_index_to_state(void* stream, const slb_grid* grid, int64_t idx_begin, int64_t idx_end,
                       double* states_dev);

/* ---- Bellman sweep (reinforcement_learning.py:65-114, 135-140, 213-279) ---------------- */
/* out_dev[i] = r(x_i,u_i) + gamma * V(mean f(x_i,u_i)) for flat indices [idx_begin, idx_end) */
int slb_bellman_sweep(void* stream, const slb_bellman* cfg, int64_t idx_begin, int64_t idx_end,
                      double* out_dev);
/* discrete_policy_optimization: actions_dev [n_actions, m]; constraint_dev [n_actions, n] or
 * NULL (value < 0 => -inf, :272-275); best_dev[i] = first argmax over actions (:278, NaN counts as
 * the maximum like np.argmax).  workspace_dev: NULL, or >= slb_bellman_argmax_workspace bytes: with
 * GP dynamics on plain RBF factors the action is then factored out of the exponent (one kernel
 * row per STATE, an [n_actions x M] x [M x states] fp64 tensor-core contraction per output)
 * instead of n_actions sweeps; the workspace size is 0 when that path does not apply. */
int64_t slb_bellman_argmax_workspace(const slb_bellman* cfg, int32_t n_actions);
int slb_bellman_argmax(void* stream, const slb_bellman* cfg, int64_t idx_begin, int64_t idx_end,
                       const double* actions_dev, int32_t n_actions, const double* constraint_dev,
                       int32_t* best_dev, double* best_value_dev, void* workspace_dev);
/* max_i |a_i - b_i| into result_dev[0] (value-iteration convergence test, test_rl.py:66-69) */
int slb_max_abs_diff(void* stream, const double* a_dev, const double* b_dev, int64_t n,
                     double* result_dev);

#ifdef __cplusplus
}
#endif
#endif /* SLB200_H */
