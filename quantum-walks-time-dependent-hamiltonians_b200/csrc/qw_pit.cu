// Parallel-in-time evaluation of the fixed-step RK4 of QW_Search.py:149-164 for a HANDFUL of
// points on a small graph (BASELINE config 1: cycle N = 16, one (T, gamma) point; the objective
// calls of the reference's SciPy optimisers, Schroedinger_Solver.py:283-361, one point at a time).
//
// One point offers no parallelism over points, and a ring of 16 sites none over sites: the
// warp-packed kernel needs 4 x ~50 ns per step, 208 us for 1024 steps -- slower than one host
// core.  But an RK4 step is a LINEAR map of the state, psi_{n+1} = P_n psi_n with an N x N complex
// matrix P_n that depends on the step's operator scalars only, so the time axis can be cut:
//   phase A  the S steps are cut into chunks of K; for every chunk c and every basis vector e_j
//            one warp integrates e_j through the chunk's steps with the ordinary stage code
//            (one site per lane, halo by shuffles): column j of M_c = P_{cK+K-1} ... P_{cK}.
//            All chunks and columns at once: N x ceil(S / K) warps, K steps deep;
//   phase B  one warp per point applies the chunk matrices in order, psi <- M_c psi (lane i
//            holds row i; the next chunk's columns are in flight while this one is summed).
// Depth K + ceil(S / K) matrix-vector products instead of S steps: K ~ sqrt(2 S).  The result is
// the RK4 result up to the rounding of the regrouped sums (parity 1e-10 is checked in the tests).
// Any dimension 3 <= N <= 32 (cycle) / 1 <= N <= 32 (complete), both step rules, psi on request.
#include <cstdlib>

#include "qw_common.cuh"

namespace {

constexpr int PIT_KMAX = 256;            // steps per chunk the coefficient table holds

// (a, d) of the three stage times of the chunk's steps -> shared memory
__device__ __forceinline__ void pit_table(double2 *tab, const QwParams &prm, const QwPoint &pt,
                                          const int64_t n0, const int cnt)
{
    for (int i = threadIdx.x; i < 3 * cnt; i += blockDim.x) {
        const int s = i / 3, k = i - 3 * s;
        double t = (double)(n0 + s) * pt.h;
        if (k == 1) t += 0.5 * pt.h;
        if (k == 2) t += pt.h;
        tab[i] = qw_operator_scalars(t / pt.T, pt.gamma, prm.schedule, prm.gamma_on);
    }
    __syncthreads();
}

// -i (a L + d |w><w|) x for the lane's site; x of the neighbours by shuffles
template <bool CYCLE>
__device__ __forceinline__ double2 pit_rhs(const double2 x, const double2 ad, const int left,
                                           const int right, const bool used, const bool marked,
                                           const double nd)
{
    const unsigned full = 0xffffffffu;
    double lr, li;
    if (CYCLE) {
        const double ar = __shfl_sync(full, x.x, left), ai = __shfl_sync(full, x.y, left);
        const double br = __shfl_sync(full, x.x, right), bi = __shfl_sync(full, x.y, right);
        lr = fma(2.0, x.x, -(ar + br));
        li = fma(2.0, x.y, -(ai + bi));
    } else {
        double sr = used ? x.x : 0.0, si = used ? x.y : 0.0;
        sr = qw_warp_sum(sr);
        si = qw_warp_sum(si);
        lr = fma(nd, x.x, -sr);
        li = fma(nd, x.y, -si);
    }
    double hr = ad.x * lr, hi = ad.x * li;                 // H x
    if (marked) { hr = fma(ad.y, x.x, hr); hi = fma(ad.y, x.y, hi); }
    return make_double2(hi, -hr);                          // -i H x
}

// block (chunk, point); warp j integrates basis vector e_j through the chunk
template <bool CYCLE>
__global__ void __launch_bounds__(1024) pit_chunk_kernel(const QwParams prm, const int K,
                                                         const int chunks_max,
                                                         double2 *__restrict__ M)
{
    __shared__ double2 tab[3 * PIT_KMAX];
    const int n = (int)prm.n;
    const int lane = threadIdx.x & 31, col = threadIdx.x >> 5;
    const int64_t q = blockIdx.y;
    const int c = blockIdx.x;
    const QwPoint pt = qw_point(prm, q);
    const int64_t n0 = (int64_t)c * K;
    if (n0 >= pt.steps || c >= chunks_max) return;         // block-uniform: chunk not needed
    const int cnt = (int)((pt.steps - n0 < K) ? (pt.steps - n0) : K);
    pit_table(tab, prm, pt, n0, cnt);

    const bool used = lane < n;
    const int left = used ? (lane == 0 ? n - 1 : lane - 1) : lane;
    const int right = used ? (lane == n - 1 ? 0 : lane + 1) : lane;
    const bool marked = used && lane == (int)prm.target;
    const double nd = (double)n, h = pt.h;
    double2 y = make_double2((used && lane == col) ? 1.0 : 0.0, 0.0);
    for (int s = 0; s < cnt; ++s) {
        const double2 ad1 = tab[3 * s], ad2 = tab[3 * s + 1], ad4 = tab[3 * s + 2];
        const double2 k1 = pit_rhs<CYCLE>(y, ad1, left, right, used, marked, nd);
        const double2 u2 = make_double2(fma(0.5 * h, k1.x, y.x), fma(0.5 * h, k1.y, y.y));
        const double2 k2 = pit_rhs<CYCLE>(u2, ad2, left, right, used, marked, nd);
        const double2 u3 = make_double2(fma(0.5 * h, k2.x, y.x), fma(0.5 * h, k2.y, y.y));
        const double2 k3 = pit_rhs<CYCLE>(u3, ad2, left, right, used, marked, nd);
        const double2 u4 = make_double2(fma(h, k3.x, y.x), fma(h, k3.y, y.y));
        const double2 k4 = pit_rhs<CYCLE>(u4, ad4, left, right, used, marked, nd);
        y.x = fma(h / 6.0, (k1.x + k4.x) + 2.0 * (k2.x + k3.x), y.x);
        y.y = fma(h / 6.0, (k1.y + k4.y) + 2.0 * (k2.y + k3.y), y.y);
    }
    if (used)            // column `col` of M_c, rows contiguous
        M[(((size_t)q * chunks_max + c) * n + col) * n + lane] = y;
}

__device__ __forceinline__ void pit_cp_async16(void *smem, const void *gmem)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(
                     (uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}

// one warp per point: psi <- M_c psi for c = 0, 1, ...; then p, norm, psi.  The chunk matrices
// stream through a double buffer in shared memory (cp.async): chunk c + 1 is in flight while
// chunk c is applied.
__global__ void __launch_bounds__(32) pit_combine_kernel(const QwParams prm, const int K,
                                                         const int chunks_max,
                                                         const double2 *__restrict__ M)
{
    __shared__ __align__(16) double2 buf[2][32 * 32];
    const unsigned full = 0xffffffffu;
    const int n = (int)prm.n, lane = threadIdx.x, nn = n * n;
    const int64_t q = blockIdx.x;
    const QwPoint pt = qw_point(prm, q);
    const bool used = lane < n;
    int chunks = (int)((pt.steps + K - 1) / K);
    if (chunks > chunks_max) chunks = chunks_max;          // cannot happen: the host's bound
    double2 psi = make_double2(used ? 1.0 / sqrt((double)n) : 0.0, 0.0);
    const double2 *Mq = M + (size_t)q * chunks_max * nn;
    if (chunks > 0)
        for (int i = lane; i < nn; i += 32) pit_cp_async16(&buf[0][i], &Mq[i]);
    asm volatile("cp.async.commit_group;" ::: "memory");
    for (int c = 0; c < chunks; ++c) {
        if (c + 1 < chunks)
            for (int i = lane; i < nn; i += 32)
                pit_cp_async16(&buf[(c + 1) & 1][i], &Mq[(size_t)(c + 1) * nn + i]);
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 1;" ::: "memory");
        __syncwarp();
        const double2 *m = buf[c & 1];
        // row `lane` of M_c times psi; four partial sums shorten the dependent chain
        double ar[4] = {0.0, 0.0, 0.0, 0.0}, ai[4] = {0.0, 0.0, 0.0, 0.0};
        for (int j0 = 0; j0 < n; j0 += 4) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int j = j0 + k;
                const double xr = __shfl_sync(full, psi.x, j & 31), xi = __shfl_sync(full, psi.y, j & 31);
                const double2 e = (used && j < n) ? m[j * n + lane] : make_double2(0.0, 0.0);
                ar[k] = fma(e.x, xr, fma(-e.y, xi, ar[k]));
                ai[k] = fma(e.x, xi, fma(e.y, xr, ai[k]));
            }
        }
        psi = make_double2((ar[0] + ar[1]) + (ar[2] + ar[3]), (ai[0] + ai[1]) + (ai[2] + ai[3]));
        __syncwarp();                                      // buf[c & 1] is free for chunk c + 2
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    const double mine = used ? fma(psi.x, psi.x, psi.y * psi.y) : 0.0;
    const double nrm = qw_warp_sum(mine);
    const double pw = __shfl_sync(full, mine, (int)prm.target);
    if (lane == 0) {
        prm.p[pt.q] = pw / nrm;
        if (prm.norm) prm.norm[pt.q] = nrm;
    }
    if (prm.psi && used) prm.psi[(size_t)pt.q * (size_t)n + lane] = psi;
}

}  // namespace

// steps per chunk for a run of at most `steps` steps: ~ sqrt(2 S) (a step costs about half a
// matrix-vector product), at least 8, at most the coefficient table
int qw_pit_chunk(int64_t steps)
{
    static double factor = 0.0;
    if (factor == 0.0) {
        const char *env = getenv("QW_PIT_FACTOR");           // development knob
        factor = env ? atof(env) : 2.0;
        if (!(factor > 0.0)) factor = 2.0;
    }
    int k = (int)ceil(sqrt(factor * (double)steps));
    if (k < 8) k = 8;
    if (k > PIT_KMAX) k = PIT_KMAX;
    return k;
}

// A handful of points on a graph of at most 32 vertices with enough steps to cut, and a bound on
// the step count known on the HOST (the uniform n_steps rule, or steps_bound from a caller that
// has T on the host): the latency regime.
bool qw_pit_supported(const QwParams &prm, int64_t steps_bound)
{
    if (prm.flags & (QW_FLAG_FORCE_GENERIC | QW_FLAG_FORCE_STREAM | QW_FLAG_NO_PACK |
                     QW_FLAG_NO_HORNER | QW_FLAG_CYCLE_LEGACY | QW_FLAG_EXACT_SUM |
                     QW_FLAG_FIXED_GEOMETRY))
        return false;
    if (prm.n > 32 || (prm.topology == QW_TOPO_CYCLE && prm.n < 3)) return false;
    if (prm.n_points < 1 || prm.n_points > 16 || prm.order) return false;
    if (steps_bound <= 0) {
        if (prm.step_size > 0.0 || prm.n_steps) return false;
        steps_bound = prm.n_steps_uniform;
    }
    return steps_bound >= 128 && steps_bound <= (int64_t)PIT_KMAX * 4096;
}

size_t qw_pit_scratch_bytes(const QwParams &prm, int64_t steps_bound)
{
    const int K = qw_pit_chunk(steps_bound);
    const int64_t chunks = (steps_bound + K - 1) / K;
    return (size_t)prm.n_points * (size_t)chunks * (size_t)(prm.n * prm.n) * sizeof(double2);
}

cudaError_t qw_launch_pit(const QwParams &prm, int64_t steps_bound, double2 *scratch,
                          cudaStream_t st)
{
    const int K = qw_pit_chunk(steps_bound);
    const int chunks = (int)((steps_bound + K - 1) / K);
    const dim3 grid((unsigned)chunks, (unsigned)prm.n_points, 1);
    const int threads = 32 * (int)prm.n;
    if (prm.topology == QW_TOPO_CYCLE)
        pit_chunk_kernel<true><<<grid, threads, 0, st>>>(prm, K, chunks, scratch);
    else
        pit_chunk_kernel<false><<<grid, threads, 0, st>>>(prm, K, chunks, scratch);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    pit_combine_kernel<<<(unsigned)prm.n_points, 32, 0, st>>>(prm, K, chunks, scratch);
    return cudaGetLastError();
}
