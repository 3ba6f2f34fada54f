/*
 * qw_b200.h -- C ABI of the B200-native RK4 heatmap path.
 *
 * The reference (mgarbellini/Quantum-Walks-Time-Dependent-Hamiltonians) is plain Python
 * with no FFI of its own; its boundary is a set of Python call signatures plus module
 * globals.  This header is the C ABI placed directly beneath those
 * signatures: plain pointers and sizes, no torch types.  Each entry point names the
 * reference code it replaces (paths relative to the reference's Code/ directory).
 * INTEGRATION.md shows the ctypes stub a maintainer of the reference would add.
 *
 * Conventions
 *   - every function returns int: 0 ok, <0 invalid argument (QW_E_*), >0 a cudaError_t.
 *     qw_last_error() returns a human-readable message for the calling thread.
 *   - "_dev" functions take DEVICE pointers, are stream-ordered (cudaStream_t passed as
 *     void*; NULL = legacy default stream) and allocate nothing persistent, except the
 *     large-N paths which take scratch from the stream-ordered pool (cudaMallocAsync).
 *     Outputs are valid once the stream has been synchronised.
 *   - "_host" functions take HOST pointers, do their own H2D/D2H copies on the given
 *     device and return when the results are in the host buffers.
 *   - the library keeps no global mutable state besides counters and is safe from one
 *     host thread per GPU.
 *   - complex128 arrays are interleaved (re, im) doubles, as numpy lays them out.
 */
#ifndef QW_B200_H
#define QW_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QW_ABI_VERSION 1

/* topology -- QW_Search.py:70-81 (0 complete, 1 cycle); BCB.py:30-67 'complete'/'circular' */
enum { QW_TOPO_COMPLETE = 0, QW_TOPO_CYCLE = 1 };

/* schedule g_T(t/T) -- QW_Search.py:104-109 fixes 0..2; BCB.py:75-87 the rest;
 * QW_SCHED_STATIC is the time-independent comparator H = L - gamma|w><w| (BCB.py:101-135) */
enum {
    QW_SCHED_CERF3 = 0, QW_SCHED_LIN = 1, QW_SCHED_SQRT = 2, QW_SCHED_CBRT = 3,
    QW_SCHED_CERF5 = 4, QW_SCHED_CERF7 = 5, QW_SCHED_STATIC = 6
};

/* where gamma sits: H = (1-g) L - g*gamma |w><w|   (all reference code, BCB.py:93-97)
 *               or  H = (1-g) gamma L - g |w><w|   (thesis Eq. 2.14 p.47) */
enum { QW_GAMMA_ON_ORACLE = 0, QW_GAMMA_ON_LAPLACIAN = 1 };

enum {
    QW_E_NULL = -1,        /* required pointer is NULL */
    QW_E_DIMENSION = -2,   /* n < 1, cycle graph with n < 3, target out of range */
    QW_E_ENUM = -3,        /* unknown topology / schedule / gamma_on */
    QW_E_STEPS = -4,       /* neither step_size > 0 nor n_steps >= 0 usable */
    QW_E_SIZE = -5,        /* negative or inconsistent sizes */
    QW_E_UNSUPPORTED = -6  /* configuration not supported by this build */
};

/* The module globals the reference reads at call time (BCB.py:20-24: dimension, topology,
 * step_function) plus the fixed-step rule of QW_Search.py:149-164. */
typedef struct qw_problem {
    int64_t n;          /* 'dimension' */
    int32_t topology;   /* QW_TOPO_* */
    int32_t schedule;   /* QW_SCHED_* ('step_function') */
    int32_t gamma_on;   /* QW_GAMMA_ON_* */
    int32_t flags;      /* QW_FLAG_* */
    int64_t target;     /* marked vertex; <0 = int((n-1)/2) as BCB.py:96-97 */
    double  step_size;  /* >0: h = step_size, n = int(T/h) steps (QW_Search.py:153-154);
                           else h = T/n_steps */
    int64_t n_steps;    /* uniform step count when step_size <= 0 and no per-point array */
} qw_problem;

enum {
    QW_FLAG_NONE = 0,
    QW_FLAG_FORCE_GENERIC = 1,   /* use the global-memory kernel whatever n is (tests) */
    QW_FLAG_FORCE_STREAM = 2,    /* use the tile-fused streaming kernel (cycle graph) */
    QW_FLAG_EXACT_SUM = 4,       /* complete graph: block-reduce sum(u) in every stage
                                    instead of once per step (tests) */
    QW_FLAG_STREAM_KB1 = 8,      /* streaming kernel: 1 RK4 step per pass over HBM (the plain
                                    32 B/site-update roofline) instead of the default 8 */
    QW_FLAG_STREAM_KB2 = 16,     /* streaming kernel: 2 steps per pass; KB1 | KB2: 4 steps */
    QW_FLAG_CYCLE_LEGACY = 32,   /* first-generation resident kernels (cycle: stage code without
                                    edge-first ordering; complete: blocking reduction per
                                    step); kept for A/B measurements */
    QW_FLAG_CYCLE_SPLIT = 64,    /* second generation with mbarrier arrive/wait instead of one
                                    __syncthreads per stage (measured slower; A/B only) */
    QW_FLAG_OCC4 = 128,          /* resident kernels, <= 128 threads: 4 blocks/SM at 128
                                    registers instead of the default 3 at 168 (A/B only);
                                    nested cycle kernel, 5 or 6 warps: one block per SM at 200
                                    registers instead of two at 168 (A/B);
                                    complete graph: one auxiliary warp with both duties
                                    instead of scalar warp + table warp (above 3072 sites) or
                                    of duties folded into the worker warps (up to 3072) (A/B);
                                    complete graph above 4096 vertices: a point split into
                                    blocks of 2049 .. 4096 sites instead of the measured-cost
                                    rule over blocks of 1024 .. 2048 sites (A/B);
                                    cluster kernel: CTAs of up to 4096 sites above 8192 (A/B) */
    QW_FLAG_R8 = 256,            /* resident kernels: at most 8 sites per thread (default: 16
                                    where n % 16 == 0 and n/16 >= 32; A/B only) */
    QW_FLAG_NO_PACK = 512,       /* do not use the warp-packed kernel for small rings; half-ring
                                    kernels: one point per warp and 9 or 17 sites per lane only
                                    (one warp per point), 17 sites per thread only (block per
                                    point) instead of the geometry chosen for the fill (A/B) */
    QW_FLAG_COMPLETE_FREE = 1024,/* complete graph: barrier-free variant (stage sums through an
                                    acquire/release table) instead of one barrier per stage;
                                    measured equal, A/B only */
    QW_FLAG_NO_HORNER = 2048,    /* resident kernels: classic stage form (stage input +
                                    accumulator, 30 FP64 instructions per site-update) instead
                                    of the nested form with tabulated oracle corrections (24);
                                    A/B and cross-checks */
    QW_FLAG_NO_ROTATE = 4096,    /* nested-form kernels: keep the marked vertex in warp 0 of
                                    every block instead of alternating the owner warp with the
                                    physical warp slot (A/B) */
    QW_FLAG_MIRROR = 8192,       /* cycle graph: integrate only the half ring w .. w + n/2 with
                                    reflecting ends (psi is mirror symmetric about the marked
                                    vertex for the uniform initial state every caller of the
                                    reference uses): same p, norm and psi with half the
                                    arithmetic.  6 <= n < 160: warp-packed half rings;
                                    160 <= n <= 1087: one warp per point, or -- half rings of
                                    at most 16 lanes (n <= 543) in batches that fill the
                                    machine -- two or four points per warp; up to
                                    8702: block-per-point kernel on the half ring; above:
                                    streaming kernel with half the HBM traffic.
                                    Complete graph: all unmarked vertices carry one common
                                    amplitude, one thread integrates the pair (psi_w, psi_other)
                                    of a point, any n, O(n_steps) work.
                                    Opt-in; ignored where it does not apply. */
    QW_FLAG_FORCE_CLUSTER = 16384, /* cycle graph, 4097 .. 32768 sites: the thread-block-cluster
                                    kernel also where the bulk-copy streaming kernel is preferred
                                    (above 8192 sites when n is a multiple of 128); tests, A/B */
    QW_FLAG_FIXED_GEOMETRY = 32768 /* kernel and lane geometry as for a call that fills the
                                    machine, whatever its size (no latency geometries, no
                                    parallel-in-time path): a point's result then depends on the
                                    problem only, not on how a sweep is cut into calls -- the
                                    shards of a multi-GPU heatmap are bitwise the rows of the
                                    one-GPU heatmap (parallel_routine sets it) */
};

/* ---- the hot path -------------------------------------------------------------------- */

/* Batch of independent (T_q, gamma_q) points: replaces the body of the loops
 * Schroedinger_Solver.py:438-441 / BCB.py:250-253, i.e. evaluate_probability (BCB.py:193-203)
 * -> solve_schrodinger_equation (BCB.py:177-190) -> schrodinger_equation (BCB.py:163-174)
 * -> generate_hamiltonian (BCB.py:28-99), with the fixed-step RK4 of QW_Search.py:149-164
 * in place of solve_ivp.  Also the robustness / localization ensembles (Robustness.py).
 *   T, gamma     [n_points] doubles
 *   n_steps      [n_points] int64 or NULL (then prob->n_steps; ignored if step_size > 0)
 *   order        [n_points] int64 or NULL: the q-th thread block handles point order[q]
 *                (host-side longest-first ordering when work per point differs)
 *   p, norm      [n_points] doubles: |psi_w|^2/||psi||^2 and ||psi||^2 (norm nullable)
 *   psi          NULL or complex128[n_points][n]: psi(T), not renormalised
 */
int qw_rk4_batch_dev(const qw_problem *prob, const double *T, const double *gamma,
                     const int64_t *n_steps, const int64_t *order, int64_t n_points,
                     double *p, double *norm, void *psi, void *stream);

/* Rows [beta_begin, beta_end) of the heatmap probability[beta][T]: replaces
 * grid_probability_evaluation (Schroedinger_Solver.py:426-444) / grid_eval (BCB.py:239-259);
 * one call = one Ray task of parallel_routine (BCB.py:279-281).
 *   time_array [n_time], beta_array [n_beta] doubles
 *   p, norm    [(beta_end-beta_begin)][n_time] row-major (norm nullable)
 */
int qw_rk4_grid_dev(const qw_problem *prob, const double *time_array, int64_t n_time,
                    const double *beta_array, int64_t n_beta, int64_t beta_begin,
                    int64_t beta_end, double *p, double *norm, void *stream);

/* Several heatmaps in as few launches as possible: the reference's production sweep
 * (BCB_latest.py:309-320: 4 schedules x 35 dimensions, 300 points each) and the three schedules
 * of BASELINE config 2 are many SMALL problems, each far too small to fill the GPU.  Jobs whose
 * warp-packed kernel coincides (same sites-per-lane geometry and step rule; dimension, schedule,
 * target, gamma placement and axes may differ) go out as ONE grid in which every warp looks its
 * problem up in a job table; the others are launched one after the other on the same stream.
 * Two or more such grids run CONCURRENTLY on streams owned by the library (forked from and
 * joined back into `stream`: work queued on `stream` before the call is seen by all of them,
 * work queued after it waits for all of them), and with a fixed step size a grid walks the
 * problems T column by T column, longest first.  Output buffers of different jobs must not
 * overlap.  Results are bitwise those of separate qw_rk4_grid_dev calls.  All pointers are
 * device pointers; jobs itself is a host array. */
typedef struct qw_grid_job {
    qw_problem prob;
    const double *time_array;
    int64_t n_time;
    const double *beta_array;
    int64_t n_beta, beta_begin, beta_end;
    double *p;             /* [(beta_end-beta_begin)][n_time] */
    double *norm;          /* same shape, nullable */
} qw_grid_job;

int qw_rk4_multi_grid_dev(const qw_grid_job *jobs, int64_t n_jobs, void *stream);

/* Host-buffer forms of the two calls above (copies inside; device = CUDA ordinal).  Staging
 * (one pinned host buffer, one device buffer, one stream) is kept per host thread and reused, so
 * a call allocates nothing; small calls are zero-copy (the kernels read and write the pinned
 * buffer).  Up to 16 points on a graph of at most 32 vertices take the parallel-in-time kernels
 * (qw_pit.cu): the single-point calls of the reference are latency bound. */
int qw_rk4_batch_host(const qw_problem *prob, const double *T, const double *gamma,
                      const int64_t *n_steps, int64_t n_points, double *p, double *norm,
                      void *psi, int device);
int qw_rk4_grid_host(const qw_problem *prob, const double *time_array, int64_t n_time,
                     const double *beta_array, int64_t n_beta, int64_t beta_begin,
                     int64_t beta_end, double *p, double *norm, int device);

/* Exact time-independent evolution psi(t) = exp(-i H t) psi0 with H = L - gamma|w><w| (or
 * gamma L - |w><w| under QW_GAMMA_ON_LAPLACIAN): replaces evaluate_time_independent_probability
 * (BCB.py:137-159), whose dense scipy.linalg.expm becomes a Chebyshev expansion over the
 * implicit stencil, accurate to ~1e-14.  Cycle graph with a register-resident dimension only
 * (QW_E_UNSUPPORTED otherwise; schedule 'static' through qw_rk4_* covers the rest).
 * p = |psi_w|^2 (not renormalised, as in the reference), norm = ||psi||^2 (nullable), psi
 * nullable complex128[n_points][n].  Points whose a*|t| exceeds the Bessel table (~3800,
 * a = half the spectral width) come back as NaN.  prob->schedule / step rule are ignored. */
int qw_ti_exact_batch_dev(const qw_problem *prob, const double *time, const double *gamma,
                          const int64_t *order, int64_t n_points, double *p, double *norm,
                          void *psi, void *stream);
int qw_ti_exact_grid_dev(const qw_problem *prob, const double *time_array, int64_t n_time,
                         const double *beta_array, int64_t n_beta, int64_t beta_begin,
                         int64_t beta_end, double *p, double *norm, void *stream);

/* Adaptive Dormand-Prince RK45 with SciPy's step-size controller: the integrator the reference
 * ships, solve_ivp(schrodinger_equation, [0, T], y0, method='RK45', atol, rtol)
 * (BCB.py:177-190, Schroedinger_Solver.py:88-103), on the implicit operator.  Reproduces the
 * reference's printed probabilities to ~1e-10.  rtol / atol = the module globals
 * rtolerance / atolerance (BCB.py:339-340).  Dimension up to 1024 (2048 if even).
 * nfev: optional int64[n_points], right-hand-side evaluations (solve_ivp's sol.nfev).
 * prob->schedule applies; the step rule is ignored.  p is renormalised as in qw_rk4_*. */
int qw_rk45_batch_dev(const qw_problem *prob, const double *T, const double *gamma,
                      const int64_t *order, int64_t n_points, double rtol, double atol,
                      double *p, double *norm, void *psi, int64_t *nfev, void *stream);
int qw_rk45_grid_dev(const qw_problem *prob, const double *time_array, int64_t n_time,
                     const double *beta_array, int64_t n_beta, int64_t beta_begin,
                     int64_t beta_end, double rtol, double atol, double *p, double *norm,
                     void *stream);

/* One right-hand side dy = -i H(t) y: replaces schrodinger_equation(t, y, beta, T)
 * (BCB.py:163-174) with the implicit operator; y, dy complex128[n] on the device. */
int qw_rhs_dev(const qw_problem *prob, double t, double T, double gamma, const void *y,
               void *dy, void *stream);

/* ---- fused post-integration reductions ----------------------------------------------- */

/* Ensemble statistics over p[n] (BASELINE config 4 / thesis section 2.3):
 * out[0] = mean, out[1] = population std, out[2] = min, out[3] = max,
 * out[4+k] = fraction with p >= thresholds[k]  (k < n_thresholds <= 8).
 * out is a device buffer of 4 + n_thresholds doubles; deterministic summation order. */
int qw_ensemble_stats_dev(const double *p, int64_t n, const double *thresholds,
                          int32_t n_thresholds, double *out, void *stream);

/* Robustness map of a heatmap probability[n_beta][n_time] (Robustness.py:25-31, unrounded):
 * R[j][i] = ((p[j][i]-p[j+v][i]) + (p[j][i]-p[j-v][i]))/2 with python index wrap for
 * j-v < 0 and NaN where j+v >= n_beta; plus, per time column, the first-maximum beta
 * index (Robustness.py:41-45).  R is [n_beta][n_time] (nullable), argmax_beta [n_time]
 * int64 (nullable). */
int qw_heatmap_robustness_dev(const double *p, int64_t n_beta, int64_t n_time,
                              int64_t variation, double *R, int64_t *argmax_beta,
                              void *stream);

/* Heatmap summary (thesis Eq. 2.10-2.11, the numbers the thesis tabulates per N):
 * out[0] = max p, out[1], out[2] = its beta / time index (first maximum in numpy's flattened
 * [beta][T] order); out[3] = min of tau = T/p over cells with T >= t_min (T_min = (pi/2)
 * sqrt(N), thesis p.41), out[4], out[5] = its indices; I = 1/p follows from out[0].
 * out: 6 doubles on the device. */
int qw_heatmap_summary_dev(const double *p, const double *time_array, int64_t n_beta,
                           int64_t n_time, double t_min, double *out, void *stream);

/* Adiabatic diagnostics of H(s) = (1-g(s)) L - g(s) gamma |w><w| (or the gamma-on-Laplacian
 * form) for a batch of (s_q, gamma_q), s = t/T in [0, 1]: replaces compute_energy_diff(s, beta)
 * and compute_gamma(s, beta) (Schroedinger_Solver.py:208-253; a dense scipy.linalg.eig of
 * generate_hamiltonian_derivative(s, beta, 0) per call there), the two ingredients of
 * adiabatic_theorem_check (Schroedinger_Solver.py:255-280).
 *   gap[q]      = E_1 - E_0, the two lowest eigenvalues of H(s_q)
 *   coupling[q] = |<phi_1| dH/ds |phi_0>| (nullable; NaN at s = 1 where H has no Laplacian part)
 *   e0[q]       = E_0 (nullable)
 * From the secular equation of the rank-one perturbation (cycle graph, dimension <= 4096) or
 * the 2 x 2 symmetric subspace (complete graph); no matrix is built.  reference_derivative != 0
 * reproduces the reference's cube-root slope 1/(3 cbrt(s)) (Schroedinger_Solver.py:204-206)
 * instead of d/ds s^(1/3).  prob->schedule applies, the step rule is ignored. */
int qw_adiabatic_batch_dev(const qw_problem *prob, const double *s, const double *gamma,
                           int64_t n_points, int32_t reference_derivative, double *gap,
                           double *coupling, double *e0, void *stream);

/* The whole spectrum of H(s_q) in ascending order, eigenvalues[q][0 .. n-1], for a batch of
 * (s_q, gamma_q): replaces compute_eigenvalues(dimension, gamma, time, TIME) of
 * Eigenvalues_Crossing.py:59-71 (time_dependent_hamiltonian + scipy.linalg.eig + sort per call;
 * its __main__, lines 88-112, samples s for the level-crossing figure).  s = time / TIME.
 * All roots of the secular equation of the rank-one perturbation, one thread per root (cycle
 * graph, dimension <= 4096), or the 2 x 2 symmetric block plus the (n-2)-fold level a n
 * (complete graph).  gamma >= 0 with the gamma-on-Laplacian placement (NaN otherwise).
 * prob->schedule applies, the step rule is ignored. */
int qw_spectrum_batch_dev(const qw_problem *prob, const double *s, const double *gamma,
                          int64_t n_points, double *eigenvalues, void *stream);

/* ---- introspection ------------------------------------------------------------------- */

typedef struct qw_plan {
    int32_t path;            /* 0 resident-cycle, 1 resident-complete, 2 generic, 3 stream,
                                4 warp-packed cycle, 5 / 6 resident cycle / complete in nested form,
                                7 warp-packed cycle in nested form, 8 warp-packed cycle of any
                                small dimension, 9 mirror-reduced cycle, 10 symmetry-reduced
                                complete graph (both QW_FLAG_MIRROR), 11 warp-packed complete
                                graph of any small dimension, 12 cycle graph resident across a
                                thread-block cluster (4097 .. 32768 sites), 13 warp-packed half
                                ring of a small cycle graph (QW_FLAG_MIRROR, 6 .. 159 sites),
                                14 parallel-in-time (a handful of points, at most 32 vertices) */
    int32_t sites_per_thread;
    int32_t threads_per_block;
    int32_t blocks_per_sm;   /* occupancy from cudaOccupancyMaxActiveBlocksPerMultiprocessor */
    int32_t regs_per_thread;
    int32_t smem_bytes;
    int32_t sm_count;
    int32_t steps_per_pass;  /* streaming path: RK4 steps fused per pass over HBM; packed
                                path: points per warp; cluster path: CTAs per cluster; complete
                                graph above 4096 vertices: blocks per point; else 0 */
} qw_plan;

/* Which kernel a large heatmap call (qw_rk4_grid_*) of this problem would run on the current
 * device.  Batch calls with a fixed step size use the block-per-point kernels instead of the
 * warp-packed ones (no common step count within a warp). */
int qw_plan_query(const qw_problem *prob, qw_plan *plan);

/* Sustained DFMA throughput of the current device in TFLOP/s (FMA = 2 flops), measured
 * with a dependent-free register-only kernel running for about `millis` ms: the FP64
 * roofline denominator (MEASURED_PEAKS.json carries no FP64 figure). */
int qw_fp64_peak_probe(double millis, double *tflops, void *stream);

/* Scratch of the large-N paths (streaming, catch-all, split, parallel-in-time, job tables) comes
 * from a memory pool owned by the library (one per device; up to 1 GiB stays cached between
 * calls).  This synchronises the current device and returns the pool's memory to the driver. */
int qw_scratch_trim(void);

/* Kernels launched by this library in this process so far (bench.py's gpu_launches). */
int64_t qw_launch_count(void);

const char *qw_last_error(void);
int qw_version(void);

#ifdef __cplusplus
}
#endif
#endif /* QW_B200_H */
