// C ABI of the library (include/qw_b200.h): validation, kernel selection, host-buffer
// forms.  No torch types; the Python shim passes raw device pointers and a stream handle.
#include <algorithm>
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <vector>

#include "qw_common.cuh"

cudaError_t qw_launch_resident(QwParams prm, cudaStream_t st, qw_plan *plan);
int qw_resident_sites_per_thread(int64_t n);
cudaError_t qw_generic_geometry(const QwParams &prm, int *blocks, size_t *scratch_bytes,
                                qw_plan *plan);
cudaError_t qw_launch_generic(const QwParams &prm, int blocks, cudaStream_t st);
bool qw_stream_supported(const QwParams &prm);
cudaError_t qw_stream_plan(const QwParams &prm, qw_plan *plan);
cudaError_t qw_launch_stream(const QwParams &prm, cudaStream_t st, int64_t *launches);
cudaError_t qw_launch_rhs(int64_t n, int topology, int64_t w, double a, double d,
                          const double2 *y, double2 *dy, cudaStream_t st);
cudaError_t qw_launch_stats(const double *p, int64_t n, const double *thr, int nthr, double *out,
                            cudaStream_t st);
cudaError_t qw_launch_robustness(const double *p, int64_t nb, int64_t nt, int64_t v, double *R,
                                 int64_t *arg, cudaStream_t st, int *launches);
cudaError_t qw_run_fp64_probe(double millis, double *tflops, cudaStream_t st, int *launches);
cudaError_t qw_launch_summary(const double *p, const double *tarr, int64_t nb, int64_t nt,
                              double t_min, double *out, cudaStream_t st);
bool qw_packed_supported(const QwParams &prm);
cudaError_t qw_launch_packed_multi(const QwParams *prms, int n, unsigned char *taken,
                                   cudaStream_t st, int64_t *launches);
cudaError_t qw_launch_packed(const QwParams &prm, cudaStream_t st, qw_plan *plan);
bool qw_rk45_supported(const QwParams &prm);
cudaError_t qw_launch_rk45(QwParams prm, double rtol, double atol, int64_t *nfev, cudaStream_t st);
int qw_horner_sites_per_thread(const QwParams &prm);
int qw_horner_split(const QwParams &prm);
int qw_cluster_size(const QwParams &prm);
cudaError_t qw_launch_cluster(QwParams prm, cudaStream_t st, qw_plan *plan);
cudaError_t qw_launch_horner(QwParams prm, cudaStream_t st, qw_plan *plan);
cudaError_t qw_launch_adiabatic(int64_t n, int topology, int schedule, int gamma_on, int ref,
                                const double *s, const double *gamma, int64_t n_points,
                                double *gap, double *coupling, double *e0, cudaStream_t st);
int qw_adiabatic_max_dimension();
cudaError_t qw_launch_spectrum(int64_t n, int topology, int schedule, int gamma_on, const double *s,
                               const double *gamma, int64_t n_points, double *out, cudaStream_t st);
int qw_mirror_sites_per_lane(const QwParams &prm);
cudaError_t qw_launch_mirror(const QwParams &prm, cudaStream_t st, qw_plan *plan);
bool qw_symmetric_supported(const QwParams &prm);
cudaError_t qw_launch_symmetric(const QwParams &prm, cudaStream_t st, qw_plan *plan);
bool qw_pit_supported(const QwParams &prm, int64_t steps_bound);
size_t qw_pit_scratch_bytes(const QwParams &prm, int64_t steps_bound);
cudaError_t qw_launch_pit(const QwParams &prm, int64_t steps_bound, double2 *scratch,
                          cudaStream_t st);
bool qw_cheb_supported(const QwParams &prm);
cudaError_t qw_launch_cheb(QwParams prm, cudaStream_t st);

namespace {

thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};

int fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int cuda_fail(cudaError_t e, const char *what)
{
    snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
    return (int)e;
}

int validate(const qw_problem *prob, QwParams *prm)
{
    if (!prob) return fail(QW_E_NULL, "qw_problem is NULL");
    if (prob->topology != QW_TOPO_COMPLETE && prob->topology != QW_TOPO_CYCLE)
        return fail(QW_E_ENUM, "undefined topology %d", prob->topology);
    if (prob->schedule < QW_SCHED_CERF3 || prob->schedule > QW_SCHED_STATIC)
        return fail(QW_E_ENUM, "step_function value not defined: %d", prob->schedule);
    if (prob->gamma_on != QW_GAMMA_ON_ORACLE && prob->gamma_on != QW_GAMMA_ON_LAPLACIAN)
        return fail(QW_E_ENUM, "undefined gamma_on %d", prob->gamma_on);
    if (prob->n < 1) return fail(QW_E_DIMENSION, "dimension %lld < 1", (long long)prob->n);
    if (prob->topology == QW_TOPO_CYCLE && prob->n < 3)
        return fail(QW_E_DIMENSION, "cycle graph needs dimension >= 3, got %lld",
                    (long long)prob->n);
    int64_t w = prob->target < 0 ? (prob->n - 1) / 2 : prob->target;
    if (w >= prob->n) return fail(QW_E_DIMENSION, "target %lld out of range", (long long)w);
    if (!(prob->step_size > 0.0) && prob->n_steps < 0)
        return fail(QW_E_STEPS, "need step_size > 0 or n_steps >= 0");
    memset(prm, 0, sizeof(*prm));
    prm->n = prob->n;
    prm->topology = prob->topology;
    prm->schedule = prob->schedule;
    prm->gamma_on = prob->gamma_on;
    prm->flags = prob->flags;
    prm->target = w;
    prm->step_size = prob->step_size > 0.0 ? prob->step_size : 0.0;
    prm->n_steps_uniform = prob->n_steps;
    return 0;
}

enum { PATH_RESIDENT = 0, PATH_GENERIC = 2, PATH_STREAM = 3, PATH_PACKED = 4, PATH_HORNER = 5,
       PATH_MIRROR = 9, PATH_SYMMETRIC = 10, PATH_CLUSTER = 12, PATH_PIT = 14 };

int choose_path(const QwParams &prm)
{
    if (prm.flags & QW_FLAG_FORCE_GENERIC) return PATH_GENERIC;
    if ((prm.flags & QW_FLAG_FORCE_STREAM) && qw_stream_supported(prm)) return PATH_STREAM;
    // a handful of points on a tiny graph: cut the time axis (latency regime)
    if (qw_pit_supported(prm, prm.steps_bound)) return PATH_PIT;
    if (qw_mirror_sites_per_lane(prm) > 0) return PATH_MIRROR;
    if (qw_symmetric_supported(prm)) return PATH_SYMMETRIC;
    const bool horner = !(prm.flags & (QW_FLAG_CYCLE_LEGACY | QW_FLAG_CYCLE_SPLIT |
                                       QW_FLAG_NO_HORNER | QW_FLAG_COMPLETE_FREE)) &&
                        qw_horner_sites_per_thread(prm) > 0;
    const bool packed = !(prm.flags & (QW_FLAG_CYCLE_LEGACY | QW_FLAG_NO_PACK)) &&
                        qw_packed_supported(prm);
    // above 512 sites a point needs more than one warp: block-per-point nested form
    if (horner && (prm.n > 512 || !packed)) return PATH_HORNER;
    if (packed) return PATH_PACKED;
    if (horner) return PATH_HORNER;
    if (qw_resident_sites_per_thread(prm.n) > 0) return PATH_RESIDENT;
    if (qw_cluster_size(prm) > 0) return PATH_CLUSTER;     // rings of 4097 .. 32768 sites
    if (qw_stream_supported(prm)) return PATH_STREAM;
    return PATH_GENERIC;
}

// per-device state of the library: its scratch pool and the SM count
struct DeviceState {
    std::atomic<int> ready{0};
    cudaMemPool_t pool = nullptr;
    int sms = 0;
};
DeviceState g_dev[64];
std::mutex g_dev_mutex;

DeviceState *device_state()
{
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    DeviceState &d = g_dev[dev];
    if (d.ready.load(std::memory_order_acquire)) return &d;
    std::lock_guard<std::mutex> lock(g_dev_mutex);
    if (d.ready.load(std::memory_order_relaxed)) return &d;
    if (cudaDeviceGetAttribute(&d.sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
        d.sms <= 0)
        d.sms = 148;
    cudaMemPoolProps props = {};
    props.allocType = cudaMemAllocationTypePinned;
    props.handleTypes = cudaMemHandleTypeNone;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id = dev;
    if (cudaMemPoolCreate(&d.pool, &props) == cudaSuccess) {
        uint64_t keep = 1ull << 30;        // cached between calls; more goes back to the driver
        cudaMemPoolSetAttribute(d.pool, cudaMemPoolAttrReleaseThreshold, &keep);
    } else {
        d.pool = nullptr;                  // fall back to the default pool, attributes untouched
        cudaGetLastError();
    }
    d.ready.store(1, std::memory_order_release);
    return &d;
}

int run(QwParams &prm, cudaStream_t st)
{
    if (prm.n_points == 0) return 0;
    cudaError_t e;
    const int path = choose_path(prm);
    switch (path) {
    case PATH_RESIDENT:
        e = qw_launch_resident(prm, st, nullptr);
        if (e != cudaSuccess) return cuda_fail(e, "resident kernel launch");
        g_launches += 1;
        return 0;
    case PATH_MIRROR:
        e = qw_launch_mirror(prm, st, nullptr);
        if (e != cudaSuccess) return cuda_fail(e, "mirror kernel launch");
        g_launches += 1;
        return 0;
    case PATH_SYMMETRIC:
        e = qw_launch_symmetric(prm, st, nullptr);
        if (e != cudaSuccess) return cuda_fail(e, "symmetric complete-graph kernel launch");
        g_launches += 1;
        return 0;
    case PATH_HORNER: {
        void *scratch = nullptr;
        if (const int split = qw_horner_split(prm)) {
            // partial norms of the blocks of a split point + one arrival counter per point
            const size_t part = (size_t)prm.n_points * (size_t)split * 2 * sizeof(double);
            const size_t done = (size_t)prm.n_points * sizeof(unsigned int);
            e = qw_scratch_alloc(&scratch, part + done, st);
            if (e != cudaSuccess) return cuda_fail(e, "scratch allocation");
            prm.split_part = (double *)scratch;
            prm.split_done = (unsigned int *)((char *)scratch + part);
            e = cudaMemsetAsync(prm.split_done, 0, done, st);
            if (e != cudaSuccess) { qw_scratch_free(scratch, st); return cuda_fail(e, "scratch reset"); }
        }
        e = qw_launch_horner(prm, st, nullptr);
        if (scratch) qw_scratch_free(scratch, st);
        if (e != cudaSuccess) return cuda_fail(e, "nested-form kernel launch");
        g_launches += 1;
        return 0;
    }
    case PATH_PIT: {
        const int64_t bound = prm.steps_bound > 0 ? prm.steps_bound : prm.n_steps_uniform;
        void *scratch = nullptr;
        e = qw_scratch_alloc(&scratch, qw_pit_scratch_bytes(prm, bound), st);
        if (e != cudaSuccess) return cuda_fail(e, "scratch allocation");
        e = qw_launch_pit(prm, bound, (double2 *)scratch, st);
        qw_scratch_free(scratch, st);
        if (e != cudaSuccess) return cuda_fail(e, "parallel-in-time kernel launch");
        g_launches += 2;
        return 0;
    }
    case PATH_CLUSTER:
        e = qw_launch_cluster(prm, st, nullptr);
        if (e != cudaSuccess) return cuda_fail(e, "cluster kernel launch");
        g_launches += 1;
        return 0;
    case PATH_PACKED:
        e = qw_launch_packed(prm, st, nullptr);
        if (e != cudaSuccess) return cuda_fail(e, "packed kernel launch");
        g_launches += 1;
        return 0;
    case PATH_STREAM: {
        int64_t nl = 0;
        e = qw_launch_stream(prm, st, &nl);
        g_launches += nl;
        if (e == cudaErrorMemoryAllocation) {
            cudaGetLastError();
            return fail(QW_E_UNSUPPORTED, "dimension %lld: the two state buffers of one point do "
                        "not fit the free device memory", (long long)prm.n);
        }
        if (e != cudaSuccess) return cuda_fail(e, "streaming kernel launch");
        return 0;
    }
    default: {
        int blocks = 0;
        size_t bytes = 0;
        e = qw_generic_geometry(prm, &blocks, &bytes, nullptr);
        if (e == cudaErrorMemoryAllocation) {
            cudaGetLastError();
            return fail(QW_E_UNSUPPORTED, "dimension %lld: five state vectors of one point do not "
                        "fit the free device memory", (long long)prm.n);
        }
        if (e != cudaSuccess) return cuda_fail(e, "generic kernel geometry");
        void *scratch = nullptr;
        e = qw_scratch_alloc(&scratch, bytes, st);
        if (e != cudaSuccess) return cuda_fail(e, "scratch allocation");
        prm.scratch = (double2 *)scratch;
        e = qw_launch_generic(prm, blocks, st);
        g_launches += 1;
        cudaError_t e2 = qw_scratch_free(scratch, st);
        if (e != cudaSuccess) return cuda_fail(e, "generic kernel launch");
        if (e2 != cudaSuccess) return cuda_fail(e2, "scratch free");
        return 0;
    }
    }
}

struct DevBuf {
    void *p = nullptr;
    cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 8); }
    ~DevBuf() { if (p) cudaFree(p); }
};

}  // namespace

cudaError_t qw_scratch_alloc(void **ptr, size_t bytes, cudaStream_t st)
{
    DeviceState *d = device_state();
    if (d && d->pool) return cudaMallocFromPoolAsync(ptr, bytes ? bytes : 16, d->pool, st);
    return cudaMallocAsync(ptr, bytes ? bytes : 16, st);
}

cudaError_t qw_scratch_free(void *ptr, cudaStream_t st) { return cudaFreeAsync(ptr, st); }

int qw_sm_count()
{
    DeviceState *d = device_state();
    return d ? d->sms : 148;
}

extern "C" {

int qw_scratch_trim(void)
{
    DeviceState *d = device_state();
    if (!d || !d->pool) return 0;
    cudaError_t e = cudaDeviceSynchronize();
    if (e == cudaSuccess) e = cudaMemPoolTrimTo(d->pool, 0);
    return e == cudaSuccess ? 0 : cuda_fail(e, "scratch pool trim");
}

static int batch_dev_impl(const qw_problem *prob, const double *T, const double *gamma,
                          const int64_t *n_steps, const int64_t *order, int64_t n_points,
                          double *p, double *norm, void *psi, void *stream, int64_t steps_bound)
{
    QwParams prm;
    int rc = validate(prob, &prm);
    if (rc) return rc;
    if (n_points < 0) return fail(QW_E_SIZE, "n_points < 0");
    if (n_points > 0 && (!T || !gamma || !p)) return fail(QW_E_NULL, "T, gamma and p are required");
    prm.T = T; prm.gamma = gamma; prm.n_steps = n_steps; prm.order = order;
    prm.n_points = n_points;
    prm.p = p; prm.norm = norm; prm.psi = (double2 *)psi;
    prm.steps_bound = steps_bound;
    return run(prm, (cudaStream_t)stream);
}

int qw_rk4_batch_dev(const qw_problem *prob, const double *T, const double *gamma,
                     const int64_t *n_steps, const int64_t *order, int64_t n_points,
                     double *p, double *norm, void *psi, void *stream)
{
    return batch_dev_impl(prob, T, gamma, n_steps, order, n_points, p, norm, psi, stream, 0);
}

int qw_rk4_grid_dev(const qw_problem *prob, const double *time_array, int64_t n_time,
                    const double *beta_array, int64_t n_beta, int64_t beta_begin,
                    int64_t beta_end, double *p, double *norm, void *stream)
{
    QwParams prm;
    int rc = validate(prob, &prm);
    if (rc) return rc;
    if (n_time < 0 || n_beta < 0 || beta_begin < 0 || beta_end < beta_begin || beta_end > n_beta)
        return fail(QW_E_SIZE, "bad grid extents");
    prm.n_points = (beta_end - beta_begin) * n_time;
    if (prm.n_points > 0 && (!time_array || !beta_array || !p))
        return fail(QW_E_NULL, "time_array, beta_array and p are required");
    prm.time_array = time_array; prm.beta_array = beta_array;
    prm.n_time = n_time; prm.beta_begin = beta_begin;
    prm.p = p; prm.norm = norm;
    return run(prm, (cudaStream_t)stream);
}

int qw_rk4_multi_grid_dev(const qw_grid_job *jobs, int64_t n_jobs, void *stream)
{
    if (n_jobs < 0 || n_jobs > (1 << 20)) return fail(QW_E_SIZE, "bad number of jobs");
    if (n_jobs == 0) return 0;
    if (!jobs) return fail(QW_E_NULL, "jobs is NULL");
    std::vector<QwParams> prms((size_t)n_jobs);
    for (int64_t k = 0; k < n_jobs; ++k) {
        const qw_grid_job &j = jobs[k];
        QwParams &prm = prms[(size_t)k];
        int rc = validate(&j.prob, &prm);
        if (rc) return rc;
        if (j.n_time < 0 || j.n_beta < 0 || j.beta_begin < 0 || j.beta_end < j.beta_begin ||
            j.beta_end > j.n_beta)
            return fail(QW_E_SIZE, "bad grid extents in job %lld", (long long)k);
        prm.n_points = (j.beta_end - j.beta_begin) * j.n_time;
        if (prm.n_points > 0 && (!j.time_array || !j.beta_array || !j.p))
            return fail(QW_E_NULL, "time_array, beta_array and p are required (job %lld)",
                        (long long)k);
        prm.time_array = j.time_array; prm.beta_array = j.beta_array;
        prm.n_time = j.n_time; prm.beta_begin = j.beta_begin;
        prm.p = j.p; prm.norm = j.norm;
    }
    cudaStream_t st = (cudaStream_t)stream;
    // problems that share a warp-packed kernel: one grid per kernel
    std::vector<unsigned char> taken((size_t)n_jobs, 0);
    std::vector<unsigned char> eligible((size_t)n_jobs, 0);
    std::vector<QwParams> packed;
    std::vector<int64_t> where;
    for (int64_t k = 0; k < n_jobs; ++k)
        if (prms[(size_t)k].n_points > 0 && choose_path(prms[(size_t)k]) == PATH_PACKED) {
            packed.push_back(prms[(size_t)k]);
            where.push_back(k);
        }
    if (packed.size() >= 2) {
        std::vector<unsigned char> got(packed.size(), 0);
        int64_t nl = 0;
        cudaError_t e = qw_launch_packed_multi(packed.data(), (int)packed.size(), got.data(), st, &nl);
        g_launches += nl;
        if (e != cudaSuccess) return cuda_fail(e, "multi-problem packed launch");
        for (size_t i = 0; i < packed.size(); ++i) taken[(size_t)where[i]] = got[i];
    }
    for (int64_t k = 0; k < n_jobs; ++k) {
        if (taken[(size_t)k]) continue;
        int rc = run(prms[(size_t)k], st);
        if (rc) return rc;
    }
    return 0;
}

static int run_ti_exact(QwParams &prm, cudaStream_t st)
{
    if (!qw_cheb_supported(prm))
        return fail(QW_E_UNSUPPORTED, "exact propagator: cycle graph with a register-resident "
                                      "dimension only (use schedule 'static' with RK4 otherwise)");
    if (prm.n_points == 0) return 0;
    cudaError_t e = qw_launch_cheb(prm, st);
    g_launches += 1;
    if (e != cudaSuccess) return cuda_fail(e, "Chebyshev kernel launch");
    return 0;
}

int qw_ti_exact_batch_dev(const qw_problem *prob, const double *time, const double *gamma,
                          const int64_t *order, int64_t n_points, double *p, double *norm,
                          void *psi, void *stream)
{
    QwParams prm;
    qw_problem tmp;
    if (!prob) return fail(QW_E_NULL, "qw_problem is NULL");
    tmp = *prob;
    if (!(tmp.step_size > 0.0) && tmp.n_steps < 0) tmp.n_steps = 0;   // no step rule needed
    int rc = validate(&tmp, &prm);
    if (rc) return rc;
    if (n_points < 0) return fail(QW_E_SIZE, "n_points < 0");
    if (n_points > 0 && (!time || !gamma || !p))
        return fail(QW_E_NULL, "time, gamma and p are required");
    prm.T = time; prm.gamma = gamma; prm.order = order; prm.n_points = n_points;
    prm.p = p; prm.norm = norm; prm.psi = (double2 *)psi;
    return run_ti_exact(prm, (cudaStream_t)stream);
}

int qw_ti_exact_grid_dev(const qw_problem *prob, const double *time_array, int64_t n_time,
                         const double *beta_array, int64_t n_beta, int64_t beta_begin,
                         int64_t beta_end, double *p, double *norm, void *stream)
{
    QwParams prm;
    qw_problem tmp;
    if (!prob) return fail(QW_E_NULL, "qw_problem is NULL");
    tmp = *prob;
    if (!(tmp.step_size > 0.0) && tmp.n_steps < 0) tmp.n_steps = 0;
    int rc = validate(&tmp, &prm);
    if (rc) return rc;
    if (n_time < 0 || n_beta < 0 || beta_begin < 0 || beta_end < beta_begin || beta_end > n_beta)
        return fail(QW_E_SIZE, "bad grid extents");
    prm.n_points = (beta_end - beta_begin) * n_time;
    if (prm.n_points > 0 && (!time_array || !beta_array || !p))
        return fail(QW_E_NULL, "time_array, beta_array and p are required");
    prm.time_array = time_array; prm.beta_array = beta_array;
    prm.n_time = n_time; prm.beta_begin = beta_begin;
    prm.p = p; prm.norm = norm;
    return run_ti_exact(prm, (cudaStream_t)stream);
}

static int run_rk45(QwParams &prm, double rtol, double atol, int64_t *nfev, cudaStream_t st)
{
    if (!qw_rk45_supported(prm))
        return fail(QW_E_UNSUPPORTED, "RK45: dimension up to 1024 (or even, up to 2048)");
    if (!(rtol > 0.0) || !(atol > 0.0)) return fail(QW_E_STEPS, "rtol and atol must be > 0");
    if (prm.n_points == 0) return 0;
    // solve_ivp raises rtol to 100 eps when it is smaller (scipy/integrate/_ivp/common.py)
    if (rtol < 100.0 * 2.220446049250313e-16) rtol = 100.0 * 2.220446049250313e-16;
    cudaError_t e = qw_launch_rk45(prm, rtol, atol, nfev, st);
    g_launches += 1;
    if (e != cudaSuccess) return cuda_fail(e, "RK45 kernel launch");
    return 0;
}

int qw_rk45_batch_dev(const qw_problem *prob, const double *T, const double *gamma,
                      const int64_t *order, int64_t n_points, double rtol, double atol,
                      double *p, double *norm, void *psi, int64_t *nfev, void *stream)
{
    QwParams prm;
    qw_problem tmp;
    if (!prob) return fail(QW_E_NULL, "qw_problem is NULL");
    tmp = *prob;
    if (!(tmp.step_size > 0.0) && tmp.n_steps < 0) tmp.n_steps = 0;   // no step rule needed
    int rc = validate(&tmp, &prm);
    if (rc) return rc;
    if (n_points < 0) return fail(QW_E_SIZE, "n_points < 0");
    if (n_points > 0 && (!T || !gamma || !p)) return fail(QW_E_NULL, "T, gamma and p are required");
    prm.T = T; prm.gamma = gamma; prm.order = order; prm.n_points = n_points;
    prm.p = p; prm.norm = norm; prm.psi = (double2 *)psi;
    return run_rk45(prm, rtol, atol, nfev, (cudaStream_t)stream);
}

int qw_rk45_grid_dev(const qw_problem *prob, const double *time_array, int64_t n_time,
                     const double *beta_array, int64_t n_beta, int64_t beta_begin,
                     int64_t beta_end, double rtol, double atol, double *p, double *norm,
                     void *stream)
{
    QwParams prm;
    qw_problem tmp;
    if (!prob) return fail(QW_E_NULL, "qw_problem is NULL");
    tmp = *prob;
    if (!(tmp.step_size > 0.0) && tmp.n_steps < 0) tmp.n_steps = 0;
    int rc = validate(&tmp, &prm);
    if (rc) return rc;
    if (n_time < 0 || n_beta < 0 || beta_begin < 0 || beta_end < beta_begin || beta_end > n_beta)
        return fail(QW_E_SIZE, "bad grid extents");
    prm.n_points = (beta_end - beta_begin) * n_time;
    if (prm.n_points > 0 && (!time_array || !beta_array || !p))
        return fail(QW_E_NULL, "time_array, beta_array and p are required");
    prm.time_array = time_array; prm.beta_array = beta_array;
    prm.n_time = n_time; prm.beta_begin = beta_begin;
    prm.p = p; prm.norm = norm;
    return run_rk45(prm, rtol, atol, nullptr, (cudaStream_t)stream);
}

// Staging of the host-buffer entry points, kept per host thread: one pinned host buffer, one
// device buffer and one stream, grown on demand and reused -- a call makes no allocation
// (cudaMalloc + cudaFree cost more than a small kernel: the round-1 path spent 224 of its 432 us
// on them and on the framework's tensors for BASELINE config 1).
namespace {
struct HostStage {
    int device = -1;
    cudaStream_t stream = nullptr;
    void *pinned = nullptr, *dev = nullptr;
    size_t pinned_bytes = 0, dev_bytes = 0;
    cudaError_t prepare(int d, size_t host_need, size_t dev_need)
    {
        cudaError_t e = cudaSetDevice(d);
        if (e != cudaSuccess) return e;
        if (d != device) {
            release();
            device = d;
            if ((e = cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking))) return e;
        }
        if (host_need > pinned_bytes) {
            if (pinned) cudaFreeHost(pinned);
            pinned = nullptr; pinned_bytes = 0;
            const size_t want = host_need < 65536 ? 65536 : host_need;
            if ((e = cudaHostAlloc(&pinned, want, cudaHostAllocDefault))) return e;
            pinned_bytes = want;
        }
        if (dev_need > dev_bytes) {
            if (dev) cudaFree(dev);
            dev = nullptr; dev_bytes = 0;
            if ((e = cudaMalloc(&dev, dev_need))) return e;
            dev_bytes = dev_need;
        }
        return cudaSuccess;
    }
    void release()
    {
        if (device >= 0 && cudaSetDevice(device) == cudaSuccess) {
            if (stream) cudaStreamDestroy(stream);
            if (pinned) cudaFreeHost(pinned);
            if (dev) cudaFree(dev);
        }
        stream = nullptr; pinned = nullptr; dev = nullptr;
        pinned_bytes = dev_bytes = 0; device = -1;
    }
    // no destructor work: at thread exit the CUDA context may already be gone
};
thread_local HostStage g_stage;

int64_t steps_of(double T, double h) { const double r = T / h; return r > 0.0 ? (int64_t)r : 0; }
}  // namespace

int qw_rk4_batch_host(const qw_problem *prob, const double *T, const double *gamma,
                      const int64_t *n_steps, int64_t n_points, double *p, double *norm,
                      void *psi, int device)
{
    if (n_points < 0) return fail(QW_E_SIZE, "n_points < 0");
    if (n_points > 0 && (!T || !gamma || !p)) return fail(QW_E_NULL, "T, gamma and p are required");
    if (!prob) return fail(QW_E_NULL, "qw_problem is NULL");
    if (n_points == 0) return 0;
    // the largest step count of the call (the host has T): lets a handful of points on a tiny
    // graph take the parallel-in-time path under every step rule
    int64_t bound = 0;
    if (prob->step_size > 0.0) {
        for (int64_t q = 0; q < n_points; ++q) bound = std::max(bound, steps_of(T[q], prob->step_size));
    } else if (n_steps) {
        for (int64_t q = 0; q < n_points; ++q) bound = std::max(bound, n_steps[q]);
    } else {
        bound = prob->n_steps;
    }
    // layout of the staging buffers: T | gamma | n_steps | (pad) | psi | p | norm -- psi is
    // written with 16-byte stores, so the output section starts on a 16-byte boundary
    const size_t np = (size_t)n_points;
    const size_t in_bytes = (np * 8 * (n_steps ? 3 : 2) + 15) / 16 * 16;
    const size_t psib = psi ? np * (size_t)prob->n * sizeof(double2) : 0;
    const size_t out_bytes = psib + np * 16;
    // small calls: the kernels read and write the pinned buffer directly (unified addressing) --
    // no copy commands at all, one launch and one synchronisation
    const bool zero_copy = in_bytes + out_bytes <= 65536;
    cudaError_t e = g_stage.prepare(device, in_bytes + out_bytes, zero_copy ? 0 : in_bytes + out_bytes);
    if (e != cudaSuccess) return cuda_fail(e, "staging buffers");
    char *h = (char *)g_stage.pinned;
    memcpy(h, T, np * 8);
    memcpy(h + np * 8, gamma, np * 8);
    if (n_steps) memcpy(h + np * 16, n_steps, np * 8);
    cudaStream_t st = g_stage.stream;
    char *d = zero_copy ? h : (char *)g_stage.dev;
    if (!zero_copy && (e = cudaMemcpyAsync(d, h, in_bytes, cudaMemcpyHostToDevice, st)))
        return cuda_fail(e, "H2D copy");
    void *dpsi = psi ? (void *)(d + in_bytes) : nullptr;
    double *dp = (double *)(d + in_bytes + psib), *dn = dp + np;
    int rc = batch_dev_impl(prob, (double *)d, (double *)(d + np * 8),
                            n_steps ? (int64_t *)(d + np * 16) : nullptr, nullptr, n_points, dp, dn,
                            dpsi, st, bound > 0 ? bound : 0);
    if (rc) return rc;
    if (!zero_copy && (e = cudaMemcpyAsync(h + in_bytes, d + in_bytes, out_bytes,
                                           cudaMemcpyDeviceToHost, st)))
        return cuda_fail(e, "D2H copy");
    if ((e = cudaStreamSynchronize(st))) return cuda_fail(e, "kernel execution");
    memcpy(p, h + in_bytes + psib, np * 8);
    if (norm) memcpy(norm, h + in_bytes + psib + np * 8, np * 8);
    if (psi) memcpy(psi, h + in_bytes, psib);
    return 0;
}

int qw_rk4_grid_host(const qw_problem *prob, const double *time_array, int64_t n_time,
                     const double *beta_array, int64_t n_beta, int64_t beta_begin,
                     int64_t beta_end, double *p, double *norm, int device)
{
    if (n_time < 0 || n_beta < 0 || beta_begin < 0 || beta_end < beta_begin || beta_end > n_beta)
        return fail(QW_E_SIZE, "bad grid extents");
    if (!time_array || !beta_array || !p) return fail(QW_E_NULL, "time_array, beta_array and p are required");
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return cuda_fail(e, "cudaSetDevice");
    const size_t np = (size_t)(beta_end - beta_begin) * (size_t)n_time * sizeof(double);
    DevBuf dT, dB, dP, dN;
    if ((e = dT.alloc((size_t)n_time * 8)) || (e = dB.alloc((size_t)n_beta * 8)) ||
        (e = dP.alloc(np)) || (e = dN.alloc(np)))
        return cuda_fail(e, "cudaMalloc");
    cudaStream_t st = nullptr;
    if ((e = cudaMemcpyAsync(dT.p, time_array, (size_t)n_time * 8, cudaMemcpyHostToDevice, st)) ||
        (e = cudaMemcpyAsync(dB.p, beta_array, (size_t)n_beta * 8, cudaMemcpyHostToDevice, st)))
        return cuda_fail(e, "H2D copy");
    int rc = qw_rk4_grid_dev(prob, (double *)dT.p, n_time, (double *)dB.p, n_beta, beta_begin,
                             beta_end, (double *)dP.p, (double *)dN.p, st);
    if (rc) return rc;
    if ((e = cudaMemcpyAsync(p, dP.p, np, cudaMemcpyDeviceToHost, st))) return cuda_fail(e, "D2H copy");
    if (norm && (e = cudaMemcpyAsync(norm, dN.p, np, cudaMemcpyDeviceToHost, st)))
        return cuda_fail(e, "D2H copy");
    if ((e = cudaStreamSynchronize(st))) return cuda_fail(e, "kernel execution");
    return 0;
}

int qw_rhs_dev(const qw_problem *prob, double t, double T, double gamma, const void *y, void *dy,
               void *stream)
{
    QwParams prm;
    qw_problem tmp;
    if (!prob) return fail(QW_E_NULL, "qw_problem is NULL");
    tmp = *prob;
    if (!(tmp.step_size > 0.0) && tmp.n_steps < 0) tmp.n_steps = 0;   // irrelevant here
    int rc = validate(&tmp, &prm);
    if (rc) return rc;
    if (!y || !dy) return fail(QW_E_NULL, "y and dy are required");
    // BCB.py:72-73: time == 0 or T == 0 -> H = L; every schedule has g(0) = 0 anyway
    double2 ad;
    if ((t == 0.0 || T == 0.0) && prm.schedule != QW_SCHED_STATIC)
        ad = make_double2(prm.gamma_on == QW_GAMMA_ON_ORACLE ? 1.0 : gamma, 0.0);
    else
        ad = qw_operator_scalars(t / T, gamma, prm.schedule, prm.gamma_on);
    const double a = ad.x, d = ad.y;
    cudaError_t e = qw_launch_rhs(prm.n, prm.topology, prm.target, a, d, (const double2 *)y,
                                  (double2 *)dy, (cudaStream_t)stream);
    g_launches += 1;
    if (e != cudaSuccess) return cuda_fail(e, "rhs kernel launch");
    return 0;
}

int qw_ensemble_stats_dev(const double *p, int64_t n, const double *thresholds,
                          int32_t n_thresholds, double *out, void *stream)
{
    if (!p || !out) return fail(QW_E_NULL, "p and out are required");
    if (n < 1 || n_thresholds < 0 || n_thresholds > 8) return fail(QW_E_SIZE, "bad sizes");
    if (n_thresholds > 0 && !thresholds) return fail(QW_E_NULL, "thresholds is NULL");
    cudaError_t e = qw_launch_stats(p, n, thresholds, n_thresholds, out, (cudaStream_t)stream);
    g_launches += 1;
    if (e != cudaSuccess) return cuda_fail(e, "stats kernel launch");
    return 0;
}

int qw_heatmap_robustness_dev(const double *p, int64_t n_beta, int64_t n_time, int64_t variation,
                              double *R, int64_t *argmax_beta, void *stream)
{
    if (!p) return fail(QW_E_NULL, "p is required");
    if (n_beta < 1 || n_time < 1 || variation < 0 || variation > n_beta)
        return fail(QW_E_SIZE, "bad sizes");
    int nl = 0;
    cudaError_t e = qw_launch_robustness(p, n_beta, n_time, variation, R, argmax_beta,
                                         (cudaStream_t)stream, &nl);
    g_launches += nl;
    if (e != cudaSuccess) return cuda_fail(e, "robustness kernel launch");
    return 0;
}

int qw_heatmap_summary_dev(const double *p, const double *time_array, int64_t n_beta,
                           int64_t n_time, double t_min, double *out, void *stream)
{
    if (!p || !time_array || !out) return fail(QW_E_NULL, "p, time_array and out are required");
    if (n_beta < 1 || n_time < 1) return fail(QW_E_SIZE, "bad sizes");
    cudaError_t e = qw_launch_summary(p, time_array, n_beta, n_time, t_min, out,
                                      (cudaStream_t)stream);
    g_launches += 1;
    if (e != cudaSuccess) return cuda_fail(e, "summary kernel launch");
    return 0;
}

int qw_adiabatic_batch_dev(const qw_problem *prob, const double *s, const double *gamma,
                           int64_t n_points, int32_t reference_derivative, double *gap,
                           double *coupling, double *e0, void *stream)
{
    QwParams prm;
    qw_problem tmp;
    if (!prob) return fail(QW_E_NULL, "qw_problem is NULL");
    tmp = *prob;
    if (!(tmp.step_size > 0.0) && tmp.n_steps < 0) tmp.n_steps = 0;   // no step rule needed
    int rc = validate(&tmp, &prm);
    if (rc) return rc;
    if (n_points < 0) return fail(QW_E_SIZE, "n_points < 0");
    if (n_points == 0) return 0;
    if (!s || !gamma || !gap) return fail(QW_E_NULL, "s, gamma and gap are required");
    if (prm.topology == QW_TOPO_CYCLE && prm.n > qw_adiabatic_max_dimension())
        return fail(QW_E_UNSUPPORTED, "adiabatic diagnostics: cycle graph up to dimension %d",
                    qw_adiabatic_max_dimension());
    cudaError_t e = qw_launch_adiabatic(prm.n, prm.topology, prm.schedule, prm.gamma_on,
                                        reference_derivative != 0, s, gamma, n_points, gap,
                                        coupling, e0, (cudaStream_t)stream);
    g_launches += 1;
    if (e != cudaSuccess) return cuda_fail(e, "adiabatic kernel launch");
    return 0;
}

int qw_spectrum_batch_dev(const qw_problem *prob, const double *s, const double *gamma,
                          int64_t n_points, double *eigenvalues, void *stream)
{
    QwParams prm;
    qw_problem tmp;
    if (!prob) return fail(QW_E_NULL, "qw_problem is NULL");
    tmp = *prob;
    if (!(tmp.step_size > 0.0) && tmp.n_steps < 0) tmp.n_steps = 0;   // no step rule needed
    int rc = validate(&tmp, &prm);
    if (rc) return rc;
    if (n_points < 0) return fail(QW_E_SIZE, "n_points < 0");
    if (n_points == 0) return 0;
    if (!s || !gamma || !eigenvalues) return fail(QW_E_NULL, "s, gamma and eigenvalues are required");
    if (prm.topology == QW_TOPO_CYCLE && prm.n > qw_adiabatic_max_dimension())
        return fail(QW_E_UNSUPPORTED, "spectrum: cycle graph up to dimension %d",
                    qw_adiabatic_max_dimension());
    cudaError_t e = qw_launch_spectrum(prm.n, prm.topology, prm.schedule, prm.gamma_on, s, gamma,
                                       n_points, eigenvalues, (cudaStream_t)stream);
    g_launches += 1;
    if (e != cudaSuccess) return cuda_fail(e, "spectrum kernel launch");
    return 0;
}

int qw_plan_query(const qw_problem *prob, qw_plan *plan)
{
    QwParams prm;
    int rc = validate(prob, &prm);
    if (rc) return rc;
    if (!plan) return fail(QW_E_NULL, "plan is NULL");
    memset(plan, 0, sizeof(*plan));
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&plan->sm_count, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return cuda_fail(e, "device query");
    prm.n_points = 1 << 20;
    // the plan describes a large heatmap call (grid mode): with a fixed step size the packed
    // kernels need the T columns of a grid (never dereferenced here)
    static const double dummy_axis[1] = {1.0};
    prm.time_array = dummy_axis;
    prm.beta_array = dummy_axis;
    prm.n_time = 1024;
    switch (choose_path(prm)) {
    case PATH_RESIDENT: e = qw_launch_resident(prm, nullptr, plan); break;
    case PATH_PACKED: e = qw_launch_packed(prm, nullptr, plan); break;
    case PATH_HORNER: e = qw_launch_horner(prm, nullptr, plan); break;
    case PATH_MIRROR: e = qw_launch_mirror(prm, nullptr, plan); break;
    case PATH_SYMMETRIC: e = qw_launch_symmetric(prm, nullptr, plan); break;
    case PATH_CLUSTER: e = qw_launch_cluster(prm, nullptr, plan); break;
    case PATH_PIT: plan->path = 14; plan->sites_per_thread = 1; plan->threads_per_block = 32 * (int)prm.n; break;
    case PATH_STREAM: e = qw_stream_plan(prm, plan); break;
    default: {
        int blocks; size_t bytes;
        e = qw_generic_geometry(prm, &blocks, &bytes, plan);
    }
    }
    if (e != cudaSuccess) return cuda_fail(e, "plan query");
    return 0;
}

int qw_fp64_peak_probe(double millis, double *tflops, void *stream)
{
    if (!tflops) return fail(QW_E_NULL, "tflops is NULL");
    int nl = 0;
    cudaError_t e = qw_run_fp64_probe(millis, tflops, (cudaStream_t)stream, &nl);
    g_launches += nl;
    if (e != cudaSuccess) return cuda_fail(e, "fp64 probe");
    return 0;
}

int64_t qw_launch_count(void) { return g_launches.load(); }
const char *qw_last_error(void) { return g_err; }
int qw_version(void) { return QW_ABI_VERSION; }

}  // extern "C"
