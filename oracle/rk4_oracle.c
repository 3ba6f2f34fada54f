/*
 * Plain-C restatement of the fixed-step RK4 heatmap path.  TEST INFRASTRUCTURE ONLY:
 * the checker for the CUDA path at sizes numpy cannot reach, and the timed host-CPU
 * baseline ("port") of bench.py.  The product library never links or calls this file.
 *
 * It follows, function by function, the reference's Python (paths under
 * /root/reference/Code):
 *   schedule_g        BCB.py:75-87, QW_Search.py:104-109 (thesis Eq. 2.5-2.6)
 *   operator_scalars  BCB.py:93-97 (gamma on the oracle), thesis Eq. 2.14 p.47 (gamma on
 *                     the Laplacian, no reference code), BCB.py:101-135 (static)
 *   apply_h           BCB.py:28-64 (L of the cycle / complete graph) + BCB.py:163-174
 *                     (RHS -i H y), restated as the implicit stencil / rank-one operator
 *   qwo_rk4_point     QW_Search.py:149-164 (fixed-step rule), BCB.py:177-199 (initial
 *                     state 1/sqrt(N), renormalised success probability)
 *   qwo_rk4_batch     BCB.py:239-259 (independent points; OpenMP replaces the Ray fan-out
 *                     of BCB.py:279-285)
 * Pinned against the reference's own functions through tests/golden (see
 * oracle/__init__.py).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

enum { TOPO_COMPLETE = 0, TOPO_CYCLE = 1 };
enum { S_CERF3 = 0, S_LIN = 1, S_SQRT = 2, S_CBRT = 3, S_CERF5 = 4, S_CERF7 = 5, S_STATIC = 6 };
enum { G_ORACLE = 0, G_LAPLACIAN = 1 };

static double schedule_g(double tau, int schedule)
{
    double x;
    switch (schedule) {
    case S_LIN:   return tau;
    case S_SQRT:  return sqrt(tau);
    case S_CBRT:  return cbrt(tau);
    case S_CERF3: x = 2 * tau - 1; return 0.5 * (1 + x * x * x);
    case S_CERF5: x = 2 * tau - 1; return 0.5 * (1 + x * x * x * x * x);
    case S_CERF7: x = 2 * tau - 1; return 0.5 * (1 + x * x * x * x * x * x * x);
    default:      return 1.0;
    }
}

static void operator_scalars(double tau, double gamma, int schedule, int gamma_on,
                             double *a, double *d)
{
    if (schedule == S_STATIC) {
        if (gamma_on == G_ORACLE) { *a = 1.0; *d = -gamma; }
        else                      { *a = gamma; *d = -1.0; }
        return;
    }
    double g = schedule_g(tau, schedule);
    if (gamma_on == G_ORACLE) { *a = 1 - g;           *d = -g * gamma; }
    else                      { *a = (1 - g) * gamma; *d = -g; }
}

/* v = (a L + d |w><w|) u, split re/im */
static void apply_h(int64_t n, int topo, int64_t w, double a, double d,
                    const double *ur, const double *ui, double *vr, double *vi)
{
    if (topo == TOPO_CYCLE) {
        vr[0] = a * (2 * ur[0] - ur[n - 1] - ur[1]);
        vi[0] = a * (2 * ui[0] - ui[n - 1] - ui[1]);
        for (int64_t j = 1; j < n - 1; ++j) {
            vr[j] = a * (2 * ur[j] - ur[j - 1] - ur[j + 1]);
            vi[j] = a * (2 * ui[j] - ui[j - 1] - ui[j + 1]);
        }
        vr[n - 1] = a * (2 * ur[n - 1] - ur[n - 2] - ur[0]);
        vi[n - 1] = a * (2 * ui[n - 1] - ui[n - 2] - ui[0]);
    } else {
        double sr = 0, si = 0;
        for (int64_t j = 0; j < n; ++j) { sr += ur[j]; si += ui[j]; }
        double nn = (double)n;
        for (int64_t j = 0; j < n; ++j) {
            vr[j] = a * (nn * ur[j] - sr);
            vi[j] = a * (nn * ui[j] - si);
        }
    }
    vr[w] += d * ur[w];
    vi[w] += d * ui[w];
}

/* One (T, gamma) point.  step_size > 0: h = step_size, tau = n h / T (n_steps given by the
 * caller as int(T/h)); else h = T / n_steps.  psi_out (nullable) is interleaved re,im. */
int qwo_rk4_point(int64_t n, int topo, int schedule, int gamma_on, int64_t w,
                  double T, double gamma, int64_t n_steps, double step_size,
                  double *p_out, double *norm_out, double *psi_out)
{
    if (n < 1 || (topo == TOPO_CYCLE && n < 3) || w < 0 || w >= n) return -1;
    double *buf = (double *)malloc(sizeof(double) * 10 * (size_t)n);
    if (!buf) return -2;
    double *yr = buf, *yi = buf + n, *ur = buf + 2 * n, *ui = buf + 3 * n;
    double *ar = buf + 4 * n, *ai = buf + 5 * n, *vr = buf + 6 * n, *vi = buf + 7 * n;
    double *u2r = buf + 8 * n, *u2i = buf + 9 * n;
    double amp = 1.0 / sqrt((double)n);
    for (int64_t j = 0; j < n; ++j) { yr[j] = amp; yi[j] = 0.0; }

    double h;
    if (step_size > 0) h = step_size;
    else if (T == 0.0 || n_steps <= 0) { h = 0.0; n_steps = 0; }
    else h = T / (double)n_steps;

    for (int64_t s = 0; s < n_steps; ++s) {
        double t = (double)s * h, a, d;
        /* k = -i v  =>  k.re = v.im, k.im = -v.re */
        operator_scalars(t / T, gamma, schedule, gamma_on, &a, &d);
        apply_h(n, topo, w, a, d, yr, yi, vr, vi);
        for (int64_t j = 0; j < n; ++j) {
            ar[j] = yr[j] + (h / 6.0) * vi[j];  ai[j] = yi[j] - (h / 6.0) * vr[j];
            ur[j] = yr[j] + (0.5 * h) * vi[j];  ui[j] = yi[j] - (0.5 * h) * vr[j];
        }
        operator_scalars((t + 0.5 * h) / T, gamma, schedule, gamma_on, &a, &d);
        apply_h(n, topo, w, a, d, ur, ui, vr, vi);
        for (int64_t j = 0; j < n; ++j) {
            ar[j] += (h / 3.0) * vi[j];          ai[j] -= (h / 3.0) * vr[j];
            u2r[j] = yr[j] + (0.5 * h) * vi[j];  u2i[j] = yi[j] - (0.5 * h) * vr[j];
        }
        apply_h(n, topo, w, a, d, u2r, u2i, vr, vi);
        for (int64_t j = 0; j < n; ++j) {
            ar[j] += (h / 3.0) * vi[j];  ai[j] -= (h / 3.0) * vr[j];
            ur[j] = yr[j] + h * vi[j];   ui[j] = yi[j] - h * vr[j];
        }
        operator_scalars((t + h) / T, gamma, schedule, gamma_on, &a, &d);
        apply_h(n, topo, w, a, d, ur, ui, vr, vi);
        for (int64_t j = 0; j < n; ++j) {
            yr[j] = ar[j] + (h / 6.0) * vi[j];
            yi[j] = ai[j] - (h / 6.0) * vr[j];
        }
    }
    double norm = 0;
    for (int64_t j = 0; j < n; ++j) norm += yr[j] * yr[j] + yi[j] * yi[j];
    double pw = yr[w] * yr[w] + yi[w] * yi[w];
    if (p_out) *p_out = pw / norm;
    if (norm_out) *norm_out = norm;
    if (psi_out)
        for (int64_t j = 0; j < n; ++j) { psi_out[2 * j] = yr[j]; psi_out[2 * j + 1] = yi[j]; }
    free(buf);
    return 0;
}

/* Independent points, fanned over host threads.  n_threads <= 0: all available.  Returns
 * the thread count actually used (>0) or a negative error. */
int qwo_rk4_batch(int64_t n, int topo, int schedule, int gamma_on, int64_t w,
                  const double *T, const double *gamma, const int64_t *n_steps,
                  double step_size, int64_t n_points,
                  double *p_out, double *norm_out, double *psi_out, int n_threads)
{
    int used = 1, err = 0;
#ifdef _OPENMP
    if (n_threads <= 0) n_threads = omp_get_max_threads();
    used = n_threads;
#pragma omp parallel for schedule(dynamic, 1) num_threads(n_threads)
#endif
    for (int64_t q = 0; q < n_points; ++q) {
        int rc = qwo_rk4_point(n, topo, schedule, gamma_on, w, T[q], gamma[q], n_steps[q],
                               step_size, &p_out[q], norm_out ? &norm_out[q] : 0,
                               psi_out ? psi_out + 2 * (size_t)n * (size_t)q : 0);
        if (rc) {
#ifdef _OPENMP
#pragma omp atomic write
#endif
            err = rc;
        }
    }
    return err ? err : used;
}

int qwo_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
