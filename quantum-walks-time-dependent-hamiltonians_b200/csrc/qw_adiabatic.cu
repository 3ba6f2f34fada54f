// Adiabatic diagnostics of H(s) = a(s) L + d(s)|w><w| for batches of (s, gamma) points:
// the gap E1 - E0 and the coupling |<phi_1| dH/ds |phi_0>| between the two lowest eigenstates
// (Schroedinger_Solver.py:146-253: generate_hamiltonian_derivative,
// compute_eigenvalues_eigenvectors, compute_gamma, compute_energy_diff -- a dense
// scipy.linalg.eig per point there).
//
// No matrix is built.  H is a rank-one perturbation of a L, whose spectrum is known:
//   * every eigenvector with phi(w) != 0 satisfies the secular equation
//       1 + d G(E) = 0,   G(E) = <w|(a L - E)^-1|w> = (1/N) sum_k 1 / (a lambda_k - E),
//     lambda_k = 2 - 2 cos(2 pi k / N) on the cycle; for d < 0 < a the ground state is its root
//     in [d, 0) and the first excited state its root in (0, a lambda_1) (the antisymmetric
//     partner of that level stays at a lambda_1 and does not couple);
//   * phi(w)^2 = 1 / (d^2 G2(E)),  G2(E) = (1/N) sum_k 1 / (a lambda_k - E)^2;
//   * <phi_1|L|phi_0> = -(d/a) phi_1(w) phi_0(w) (from H phi_0 = E_0 phi_0 and orthogonality),
//     so <phi_1| a' L + d' P |phi_0> = phi_1(w) phi_0(w) (d' - a' d / a).
// On the complete graph the states with phi(w) != 0 live in span{|w>, uniform rest}: a 2 x 2
// problem in closed form.  One warp per point; the k-sums are split over the lanes.
#include "qw_common.cuh"

namespace {

constexpr int AD_WARPS = 4;
constexpr int AD_MAXN = 4096;          // lambda table in shared memory

// g'(s); `ref` reproduces the reference's cube-root derivative 1 / (3 cbrt(s))
// (Schroedinger_Solver.py:204-206; the derivative of s^(1/3) is 1 / (3 cbrt(s)^2))
__device__ __forceinline__ double schedule_slope(double s, int schedule, int ref)
{
    double x;
    switch (schedule) {
    case QW_SCHED_LIN:   return 1.0;
    case QW_SCHED_SQRT:  return 0.5 / sqrt(s);
    case QW_SCHED_CBRT:  x = cbrt(s); return ref ? 1.0 / (3.0 * x) : 1.0 / (3.0 * x * x);
    case QW_SCHED_CERF3: x = 2.0 * s - 1.0; return 3.0 * x * x;
    case QW_SCHED_CERF5: x = 2.0 * s - 1.0; x *= x; return 5.0 * x * x;
    case QW_SCHED_CERF7: x = 2.0 * s - 1.0; x *= x; return 7.0 * x * x * x;
    default:             return 0.0;
    }
}

// (G, G2)(E) for the cycle, summed over the warp in a fixed order
__device__ __forceinline__ double2 resolvent(const double *lam, int n, double a, double E, int lane)
{
    double g = 0.0, g2 = 0.0;
    for (int k = lane; k < n; k += 32) {
        const double r = 1.0 / (a * lam[k] - E);
        g += r;
        g2 = fma(r, r, g2);
    }
    g = qw_warp_sum(g);
    g2 = qw_warp_sum(g2);
    return make_double2(g / n, g2 / n);
}

// root of G(E) = target in (lo, hi), G increasing there with G(lo+) < target < G(hi-)
__device__ __forceinline__ double secular_root(const double *lam, int n, double a, double target,
                                               double lo, double hi, int lane)
{
    for (int it = 0; it < 256; ++it) {
        const double mid = 0.5 * (lo + hi);
        if (!(mid > lo && mid < hi)) break;
        if (resolvent(lam, n, a, mid, lane).x < target) lo = mid; else hi = mid;
    }
    return 0.5 * (lo + hi);
}

__global__ void __launch_bounds__(32 * AD_WARPS)
adiabatic_kernel(int64_t n, int topology, int schedule, int gamma_on, int ref, const double *s_arr,
                 const double *gamma_arr, int64_t n_points, double *gap, double *coupling,
                 double *e0_out)
{
    __shared__ double lam[AD_MAXN];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (topology == QW_TOPO_CYCLE) {
        for (int k = threadIdx.x; k < n; k += blockDim.x)
            lam[k] = 2.0 - 2.0 * cospi(2.0 * (double)k / (double)n);
        __syncthreads();
    }
    const int64_t q = (int64_t)blockIdx.x * AD_WARPS + warp;
    if (q >= n_points) return;
    const double s = s_arr[q], gamma = gamma_arr[q];
    const double2 ad = qw_operator_scalars(s, gamma, schedule, gamma_on);
    const double a = ad.x, d = ad.y;
    // dH/ds = a' L + d' P
    const double slope = schedule_slope(s, schedule, ref);
    const double gl = gamma_on == QW_GAMMA_ON_ORACLE ? 1.0 : gamma;
    const double gd = gamma_on == QW_GAMMA_ON_ORACLE ? gamma : 1.0;
    const double da = -slope * gl, dd = -slope * gd;
    const double nn = (double)n;

    double E0, E1, p0, p1;                   // eigenvalues, phi(w)^2
    if (topology == QW_TOPO_CYCLE) {
        const double lam1 = a * lam[1];
        if (d == 0.0 || a == 0.0) {          // no perturbation / no Laplacian: limits
            E0 = a == 0.0 ? d : 0.0;
            E1 = a == 0.0 ? 0.0 : lam1;
            p0 = a == 0.0 ? 1.0 : 1.0 / nn;
            p1 = a == 0.0 ? 0.0 : 2.0 / nn;
        } else {
            const double target = -1.0 / d;
            E0 = secular_root(lam, (int)n, a, target, d, 0.0, lane);
            E1 = secular_root(lam, (int)n, a, target, 0.0, lam1, lane);
            p0 = 1.0 / (d * d * resolvent(lam, (int)n, a, E0, lane).y);
            p1 = 1.0 / (d * d * resolvent(lam, (int)n, a, E1, lane).y);
        }
    } else {
        // span{|w>, (N-1)^-1/2 sum_{j != w} |j>}: [[a(N-1) + d, -a sqrt(N-1)], [., a]]
        const double h11 = a * (nn - 1.0) + d, h22 = a, off = a * sqrt(nn - 1.0);
        const double m = 0.5 * (h11 + h22), del = 0.5 * (h11 - h22);
        const double r = sqrt(del * del + off * off);
        E0 = m - r;
        E1 = m + r;
        if (off == 0.0) { p0 = del < 0.0 ? 1.0 : 0.0; p1 = 1.0 - p0; }
        else {
            const double y0 = h11 - E0, y1 = h11 - E1;
            p0 = off * off / (off * off + y0 * y0);
            p1 = off * off / (off * off + y1 * y1);
        }
    }
    if (lane == 0) {
        gap[q] = E1 - E0;
        if (coupling)
            coupling[q] = (a == 0.0 && d != 0.0) ? NAN
                                                 : sqrt(p0 * p1) * fabs(dd - (d == 0.0 ? 0.0 : da * d / a));
        if (e0_out) e0_out[q] = E0;
    }
}

// ---------------------------------------------------------------------------------------------
// The whole sorted spectrum of H(s) = a L + d |w><w| for a batch of (s, gamma) points
// (Eigenvalues_Crossing.py:59-71, compute_eigenvalues: a dense scipy.linalg.eig + sort per call;
// its __main__ samples 100 values of s for the level-crossing figure).
//
// Cycle graph: L has the distinct eigenvalues mu_m = 2 - 2 cos(2 pi m / N), m = 0 .. M = N/2;
// m = 1 .. D = (N-1)/2 are doubly degenerate.  Of each degenerate pair the combination that
// vanishes on w is untouched by the oracle and stays at a mu_m; the other one, like the two
// singlets, has weight omega_m = |<w|v_m>|^2 = 2/N (1/N) on w and moves to a root of
//     f(E) = 1 + d sum_m omega_m / (a mu_m - E) = 0,
// exactly one per interval between consecutive poles plus one beyond the end the perturbation
// pushes towards (below a mu_0 - |d| .. a mu_0 for d < 0, above for d > 0).  f is monotonic
// between poles: bisection to the last bit, one thread per root.  The interlacing also fixes the
// sorted order, so the output needs no sort.  Complete graph: two roots of the 2 x 2 problem
// above plus N - 2 copies of a N.
constexpr int SP_THREADS = 128;
constexpr int SP_MAXM = AD_MAXN / 2 + 1;

__device__ __forceinline__ double secular_f(const double *pole, int m1, int dbl, double d,
                                            double inv_n, double E)
{
    double acc = 0.0;
    for (int m = 0; m < m1; ++m) {
        const double w = (m >= 1 && m <= dbl) ? 2.0 : 1.0;
        acc += w / (pole[m] - E);
    }
    return fma(d * inv_n, acc, 1.0);
}

__global__ void __launch_bounds__(SP_THREADS)
spectrum_kernel(int64_t n, int topology, int schedule, int gamma_on, const double *s_arr,
                const double *gamma_arr, int64_t n_points, double *out)
{
    __shared__ double pole[SP_MAXM];
    const int64_t q = blockIdx.x;
    const double s = s_arr[q], gamma = gamma_arr[q];
    const double2 ad = qw_operator_scalars(s, gamma, schedule, gamma_on);
    const double a = ad.x, d = ad.y, nn = (double)n;
    double *o = out + q * n;
    const int t = threadIdx.x;
    if (!(a >= 0.0)) {                                 // gamma < 0 on the Laplacian: not ordered
        for (int64_t i = t; i < n; i += SP_THREADS) o[i] = NAN;
        return;
    }
    if (topology == QW_TOPO_COMPLETE) {
        double E0 = d, E1 = d;
        if (n >= 2) {
            const double h11 = a * (nn - 1.0) + d, h22 = a, off = a * sqrt(nn - 1.0);
            const double m = 0.5 * (h11 + h22), del = 0.5 * (h11 - h22);
            const double r = sqrt(del * del + off * off);
            E0 = m - r;
            E1 = m + r;
        }
        const double mid = a * nn;                     // N - 2 fold
        const int64_t below = (n >= 2 ? (E0 <= mid) + (E1 <= mid) : 1);   // roots placed first
        for (int64_t i = t; i < n; i += SP_THREADS) {
            double v;
            if (n == 1) v = d;
            else if (i < below) v = i == 0 ? E0 : E1;
            else if (i < below + (n - 2)) v = mid;
            else v = (i - (n - 2)) == 0 ? E0 : E1;
            o[i] = v;
        }
        return;
    }
    const int M = (int)(n / 2), D = (int)((n - 1) / 2);
    for (int m = t; m <= M; m += SP_THREADS)
        pole[m] = a * (2.0 - 2.0 * cospi(2.0 * (double)m / nn));
    __syncthreads();
    if (a == 0.0) {                                    // no Laplacian: {d, 0 x (N - 1)}
        for (int64_t i = t; i < n; i += SP_THREADS)
            o[i] = d < 0.0 ? (i == 0 ? d : 0.0) : (i == n - 1 ? d : 0.0);
        return;
    }
    const double inv_n = 1.0 / nn;
    for (int j = t; j <= M; j += SP_THREADS) {
        double root;
        if (d == 0.0) root = pole[j];
        else {
            double lo, hi;
            if (d < 0.0) { lo = j == 0 ? pole[0] + d : pole[j - 1]; hi = pole[j]; }
            else { lo = pole[j]; hi = j == M ? pole[M] + d : pole[j + 1]; }
            for (int it = 0; it < 1100; ++it) {        // until the bracket is two neighbours
                const double mid = 0.5 * (lo + hi);
                if (!(mid > lo && mid < hi)) break;
                const double f = secular_f(pole, M + 1, D, d, inv_n, mid);
                // d < 0: f falls from +inf to -inf across the interval; d > 0: it rises
                if ((f > 0.0) == (d < 0.0)) lo = mid; else hi = mid;
            }
            root = 0.5 * (lo + hi);
        }
        const bool dbl = j >= 1 && j <= D;
        if (d <= 0.0) {                                // r_0, then (r_j, a mu_j) pairs
            const int pos = j == 0 ? 0 : j + (j - 1 < D ? j - 1 : D);
            o[pos] = root;
            if (dbl) o[pos + 1] = pole[j];
        } else {                                       // (a mu_j, r_j)
            const int pos = j + (j < D ? j : D);
            o[pos] = root;
            if (dbl) o[pos - 1] = pole[j];
        }
    }
}

}  // namespace

cudaError_t qw_launch_adiabatic(int64_t n, int topology, int schedule, int gamma_on, int ref,
                                const double *s, const double *gamma, int64_t n_points,
                                double *gap, double *coupling, double *e0, cudaStream_t st)
{
    const unsigned blocks = (unsigned)((n_points + AD_WARPS - 1) / AD_WARPS);
    adiabatic_kernel<<<blocks, 32 * AD_WARPS, 0, st>>>(n, topology, schedule, gamma_on, ref, s,
                                                       gamma, n_points, gap, coupling, e0);
    return cudaGetLastError();
}

int qw_adiabatic_max_dimension() { return AD_MAXN; }

cudaError_t qw_launch_spectrum(int64_t n, int topology, int schedule, int gamma_on, const double *s,
                               const double *gamma, int64_t n_points, double *out, cudaStream_t st)
{
    spectrum_kernel<<<(unsigned)n_points, SP_THREADS, 0, st>>>(n, topology, schedule, gamma_on, s,
                                                               gamma, n_points, out);
    return cudaGetLastError();
}
