// Warp-packed register-resident RK4 for small rings (cycle graph, N = R * TPP with TPP a
// power of two <= 32), sm_100a.  BASELINE configs 1, 2 and 4 (N = 16, 64, 256).
//
// A point occupies TPP lanes of a warp, so a warp integrates 32/TPP points at once and every
// lane still carries R (up to 16) sites: the dependent chain DADD -> DFMA -> DFMA of a stage
// is hidden behind 2R independent chains per lane instead of behind other warps.  Halo values
// travel by warp shuffles -- no shared-memory mailbox and no barrier anywhere in the time
// loop.  Requires one common step count (n_steps rule): the normalised stage times
// tau = (n + c)/S are then the same for every point, so the schedule values g(tau) live in
// one small shared table per warp and each lane only scales them by its own h and gamma.
// Same arithmetic as qw_resident.cu (reference path BCB.py:163-203, step rule
// QW_Search.py:149-164).
#include <algorithm>
#include <cstring>
#include <vector>

#include "qw_stage.cuh"
#include <algorithm>
#include <mutex>
#include <vector>

#include "qw_horner_level.cuh"

namespace {

struct PackedShared {
    double2 fac[3 * QW_CHUNK];       // (1 - g, -g) at tau_n, tau_n + 1/(2S), tau_n + 1/S
};

// Which point does group `grp` of warp `warp_id` integrate?  All points of a warp must share
// the step count and the normalised stage times, so that one schedule table serves them:
//  * n_steps rule: any points will do -- warp_id * groups + grp in launch order;
//  * fixed step size on a heatmap (grid mode): the points of one T COLUMN share T, hence
//    steps = int(T / h) and tau = t / T; a warp takes `groups` consecutive beta rows of one
//    column, columns from the longest T to the shortest.
__device__ __forceinline__ bool packed_column_mode(const QwParams &prm)
{
    return prm.step_size > 0.0;
}

__device__ __forceinline__ QwPoint packed_point(const QwParams &prm, const int64_t warp_id,
                                                const int grp, const int groups, bool &valid)
{
    if (!packed_column_mode(prm)) {
        int64_t b = warp_id * groups + grp;
        valid = b < prm.n_points;
        if (!valid) b = prm.n_points - 1;                // idle groups shadow the last point
        return qw_point(prm, b);
    }
    const int64_t rows = prm.n_points / prm.n_time;
    const int64_t wpc = (rows + groups - 1) / groups;     // warps per column
    const int64_t slot = warp_id / wpc;
    int64_t row = (warp_id - slot * wpc) * groups + grp;
    valid = row < rows;
    if (!valid) row = rows - 1;
    const bool ascending = prm.time_array[0] <= prm.time_array[prm.n_time - 1];
    const int64_t col = ascending ? prm.n_time - 1 - slot : slot;
    QwPoint pt;
    pt.q = row * prm.n_time + col;
    pt.T = prm.time_array[col];
    pt.gamma = prm.beta_array[prm.beta_begin + row];
    pt.h = prm.step_size;
    const double r = pt.T / pt.h;
    pt.steps = (r > 0.0) ? (int64_t)r : 0;
    return pt;
}

// Step count of the warp.  It is the same in every lane, but in column mode it comes from a
// load through a generic pointer, which the compiler counts as divergent -- and a possibly
// divergent trip count puts divergence handling (BRA.DIV, two copies of every shuffle) into
// the time loop (C1 latency 209 -> 258 us).  Passing the value through shared memory (`slot`)
// makes it a load from a uniform shared address, which the compiler knows to be uniform.
__device__ __forceinline__ int64_t packed_steps_raw(const QwParams &prm, const int64_t warp_id,
                                                    const int groups);

__device__ __forceinline__ int64_t packed_steps(const QwParams &prm, const int64_t warp_id,
                                                const int groups, int64_t *slot)
{
    if (!packed_column_mode(prm)) return prm.n_steps_uniform;
    const int64_t raw = packed_steps_raw(prm, warp_id, groups);
    if (threadIdx.x == 0) *slot = raw;
    __syncwarp();
    return *slot;
}

__device__ __forceinline__ int64_t packed_steps_raw(const QwParams &prm, const int64_t warp_id,
                                                    const int groups)
{
    const int64_t rows = prm.n_points / prm.n_time;
    const int64_t wpc = (rows + groups - 1) / groups;
    const int64_t slot = warp_id / wpc;
    const bool ascending = prm.time_array[0] <= prm.time_array[prm.n_time - 1];
    const double r = prm.time_array[ascending ? prm.n_time - 1 - slot : slot] / prm.step_size;
    return (r > 0.0) ? (int64_t)r : 0;
}

// warps needed for the whole batch
static int64_t packed_warps(const QwParams &prm, int groups)
{
    if (prm.step_size > 0.0) {
        const int64_t rows = prm.n_points / prm.n_time;
        return prm.n_time * ((rows + groups - 1) / groups);
    }
    return (prm.n_points + groups - 1) / groups;
}

// (1 - g, -g) of the three stage times of step n: tau = (n + k/2) / S under the n_steps rule,
// tau = t / T with t = n h (+ h/2, + h) under the fixed step size (as qw_fill_table)
__device__ __forceinline__ double2 packed_factor(const QwParams &prm, const QwPoint &pt,
                                                 const int64_t n, const int k, const double inv_s)
{
    if (prm.schedule == QW_SCHED_STATIC) return make_double2(1.0, -1.0);
    double tau;
    if (packed_column_mode(prm)) {
        double t = (double)n * pt.h;
        if (k == 1) t += 0.5 * pt.h;
        if (k == 2) t += pt.h;
        tau = t / pt.T;
    } else {
        tau = ((double)n + 0.5 * (double)k) * inv_s;
    }
    const double g = qw_schedule(tau, prm.schedule);
    return make_double2(1.0 - g, -g);
}

// COLUMN: fixed step size on a heatmap (packed_point).  A template switch, because the step
// count then comes from memory and lives in a vector register; with the n_steps rule it is a
// kernel parameter and the loop control stays in the uniform datapath (C1: 5 % latency).
// The kernels below are bodies (`prm` by reference, the warp's index within its problem as an
// argument) with two entry points each: the single-problem kernel, whose parameters sit in the
// constant bank, and a multi-problem kernel, whose warps look their problem up in a job table
// (qw_rk4_multi_grid_dev: many small heatmaps in one grid).
template <int R, bool COLUMN>
__device__ __forceinline__ void packed_body(const QwParams &prm, const int64_t warp_id,
                                            const int tpp_log2)
{
    __shared__ PackedShared sh;
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x;
    const int tpp = 1 << tpp_log2, groups = 32 >> tpp_log2;
    const int li = lane & (tpp - 1), grp = lane >> tpp_log2;
    bool valid;
    const QwPoint pt = packed_point(prm, warp_id, grp, groups, valid);
    __shared__ int64_t steps_slot;
    const int64_t S = COLUMN ? packed_steps(prm, warp_id, groups, &steps_slot)
                             : prm.n_steps_uniform;
    const bool own = (li == 0);
    const int left = (li == 0) ? lane + tpp - 1 : lane - 1;
    const int right = (li == tpp - 1) ? lane - tpp + 1 : lane + 1;

    // per-lane stage weights: A = wa * a(t), B = wb * a(t), DA = da * d(t)/dfac, ...
    const double h = pt.h;                       // 0 when T == 0: the state never moves
    const double gl = prm.gamma_on == QW_GAMMA_ON_LAPLACIAN ? pt.gamma : 1.0;
    const double gd = prm.gamma_on == QW_GAMMA_ON_ORACLE ? pt.gamma : 1.0;
    const double a_half = 0.5 * h * gl, a_full = h * gl, a_6 = h / 6.0 * gl, a_3 = h / 3.0 * gl;
    const double d_half = 0.5 * h * gd, d_full = h * gd, d_6 = h / 6.0 * gd, d_3 = h / 3.0 * gd;

    State<R> s;
    const double amp = 1.0 / sqrt((double)prm.n);
#pragma unroll
    for (int r = 0; r < R; ++r) {
        s.yr[r] = amp; s.yi[r] = 0.0;
        s.ur[r] = 0.0; s.ui[r] = 0.0; s.ar[r] = 0.0; s.ai[r] = 0.0;
    }
    const double2 zero2 = make_double2(0.0, 0.0);
    const double inv_s = 1.0 / (double)S;

    for (int64_t n0 = 0; n0 < S; n0 += QW_CHUNK) {
        __syncwarp();
        for (int i = lane; i < 3 * QW_CHUNK; i += 32) {
            const int st = i / 3, k = i - 3 * st;
            sh.fac[i] = packed_factor(prm, pt, n0 + st, k, inv_s);
        }
        __syncwarp();
        const int cnt = (int)((S - n0 < QW_CHUNK) ? (S - n0) : QW_CHUNK);
        for (int sidx = 0; sidx < cnt; ++sidx) {
#define QW_PACKED_STAGE(K, TIDX, WA, WB, DWA, DWB)                                      \
            {                                                                           \
                const double2 f = sh.fac[3 * sidx + TIDX];                              \
                const double xfr = (K == 0) ? s.yr[0] : s.ur[0];                        \
                const double xfi = (K == 0) ? s.yi[0] : s.ui[0];                        \
                const double xlr = (K == 0) ? s.yr[R - 1] : s.ur[R - 1];                \
                const double xli = (K == 0) ? s.yi[R - 1] : s.ui[R - 1];                \
                const double2 hl = make_double2(__shfl_sync(full, xlr, left),           \
                                                __shfl_sync(full, xli, left));          \
                const double2 hr = make_double2(__shfl_sync(full, xfr, right),          \
                                                __shfl_sync(full, xfi, right));         \
                const double2 dd = make_double2((DWA) * f.y, (DWB) * f.y);              \
                rk_stage<R, K, true>(s, (WA) * f.x, (WB) * f.x, hl, hr, zero2, 0.0, own, dd); \
            }
            QW_PACKED_STAGE(0, 0, a_half, a_6, d_half, d_6)
            QW_PACKED_STAGE(1, 1, a_half, a_3, d_half, d_3)
            QW_PACKED_STAGE(2, 1, a_full, a_3, d_full, d_3)
            QW_PACKED_STAGE(3, 2, 0.0, a_6, 0.0, d_6)
#undef QW_PACKED_STAGE
        }
    }

    // epilogue: segmented reductions over the tpp lanes of each point
    double nrm = 0.0;
#pragma unroll
    for (int r = 0; r < R; ++r) nrm = fma(s.yr[r], s.yr[r], fma(s.yi[r], s.yi[r], nrm));
    for (int o = tpp >> 1; o > 0; o >>= 1) nrm += __shfl_xor_sync(full, nrm, o);
    if (own && valid) {
        prm.p[pt.q] = fma(s.yr[0], s.yr[0], s.yi[0] * s.yi[0]) / nrm;
        if (prm.norm) prm.norm[pt.q] = nrm;
    }
    if (prm.psi && valid) {              // undo the relabelling: logical l -> (l + w) mod n
        double2 *out = prm.psi + (size_t)pt.q * (size_t)prm.n;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            int64_t j = (int64_t)li * R + r + prm.target;
            if (j >= prm.n) j -= prm.n;
            out[j] = make_double2(s.yr[r], s.yi[r]);
        }
    }
}

template <int R, int MINB, bool COLUMN>
__global__ void __launch_bounds__(32, MINB) rk4_cycle_packed(const QwParams prm, const int tpp_log2)
{
    packed_body<R, COLUMN>(prm, blockIdx.x, tpp_log2);
}

// One entry per problem of a multi-problem launch: warps [warp_begin, warp_end) of the grid.
struct PackedJob {
    QwParams prm;
    int64_t warp_begin, warp_end;
    int32_t a, b;                    // tpp_log2, - / lanes per point, lanes with R sites
    int64_t warp_offset;             // index within the problem of the entry's first warp (a problem
                                     // with a fixed step size has one entry per T column)
};

// the calling warp's job -> shared memory; returns the warp's index within the job
__device__ __forceinline__ int64_t packed_job_fetch(const PackedJob *jobs, const int n_jobs,
                                                    PackedJob *mine)
{
    const int64_t w = blockIdx.x;
    int lo = 0, hi = n_jobs - 1;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (w >= jobs[mid].warp_end) lo = mid + 1; else hi = mid;
    }
    static_assert(sizeof(PackedJob) % 8 == 0, "copied in 8-byte words");
    const unsigned long long *src = reinterpret_cast<const unsigned long long *>(jobs + lo);
    unsigned long long *dst = reinterpret_cast<unsigned long long *>(mine);
    for (int i = threadIdx.x; i < (int)(sizeof(PackedJob) / 8); i += 32) dst[i] = src[i];
    __syncwarp();
    return w - mine->warp_begin + mine->warp_offset;
}

template <int R, int MINB, bool COLUMN>
__global__ void __launch_bounds__(32, MINB) rk4_cycle_packed_multi(const PackedJob *jobs,
                                                                   const int n_jobs)
{
    __shared__ PackedJob job;
    const int64_t w = packed_job_fetch(jobs, n_jobs, &job);
    packed_body<R, COLUMN>(job.prm, w, job.a);
}

// ---------------------------------------------------------------------------------------------
// Packed layout for ANY small ring (the reference's own runs use odd dimensions 3 .. 71):
// n = L (R - 1) + nfull sites of a point are dealt out over L lanes, R sites to the first
// nfull lanes and R - 1 to the others (possible whenever n >= R^2); a warp integrates
// G = 32 / L points, lanes beyond G L idle.  The last slot of a short lane is a ghost that is
// loaded with the right neighbour's first value before every stage, and a short lane hands
// slot R-2 to its left neighbour: two selects per stage, no FP64 work.  Stage form.
// MIRROR (QW_FLAG_MIRROR, rings below 160 sites; larger ones: rk4_cycle_mirror): the lanes hold
// the half ring w .. w + n/2 (n/2 + 1 sites) with reflecting ends -- left of the marked vertex
// its own right neighbour, right of the antipode the site before it (even n) or the last site
// itself (odd n).  psi is mirror symmetric about the marked vertex for the uniform start, so p,
// norm and psi are those of the full ring with half the arithmetic.
template <int R, bool MIRROR>
__device__ __forceinline__ void packed_any_body(const QwParams &prm, const int64_t warp_id,
                                                const int L, const int nfull)
{
    static_assert(R >= 2, "a short lane needs at least one real site");
    __shared__ PackedShared sh;
    const unsigned full_mask = 0xffffffffu;
    const int lane = threadIdx.x;
    const int groups = 32 / L;
    const int grp = lane / L, li = lane - grp * L;
    const bool lane_used = grp < groups;
    bool valid;
    const QwPoint pt = packed_point(prm, warp_id, lane_used ? grp : groups - 1, groups, valid);
    valid = valid && lane_used;
    __shared__ int64_t steps_slot;
    const int64_t S = packed_steps(prm, warp_id, groups, &steps_slot);
    const bool own = lane_used && li == 0;
    const bool full = !lane_used || li < nfull;        // R sites (idle lanes: R zeros)
    const int left = !lane_used ? lane : (li == 0 ? lane + L - 1 : lane - 1);
    const int right = !lane_used ? lane : (li == L - 1 ? lane - L + 1 : lane + 1);
    const bool m_first = MIRROR && lane_used && li == 0, m_last = MIRROR && lane_used && li == L - 1;
    const bool even = (prm.n & 1) == 0;

    const double h = pt.h;                       // 0 when T == 0: the state never moves
    const double gl = prm.gamma_on == QW_GAMMA_ON_LAPLACIAN ? pt.gamma : 1.0;
    const double gd = prm.gamma_on == QW_GAMMA_ON_ORACLE ? pt.gamma : 1.0;
    const double a_half = 0.5 * h * gl, a_full = h * gl, a_6 = h / 6.0 * gl, a_3 = h / 3.0 * gl;
    const double d_half = 0.5 * h * gd, d_full = h * gd, d_6 = h / 6.0 * gd, d_3 = h / 3.0 * gd;

    State<R> s;
    const double amp = lane_used ? 1.0 / sqrt((double)prm.n) : 0.0;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        s.yr[r] = amp; s.yi[r] = 0.0;
        s.ur[r] = 0.0; s.ui[r] = 0.0; s.ar[r] = 0.0; s.ai[r] = 0.0;
    }
    const double2 zero2 = make_double2(0.0, 0.0);
    const double inv_s = 1.0 / (double)S;

    for (int64_t n0 = 0; n0 < S; n0 += QW_CHUNK) {
        __syncwarp();
        for (int i = lane; i < 3 * QW_CHUNK; i += 32) {
            const int st = i / 3, k = i - 3 * st;
            sh.fac[i] = packed_factor(prm, pt, n0 + st, k, inv_s);
        }
        __syncwarp();
        const int cnt = (int)((S - n0 < QW_CHUNK) ? (S - n0) : QW_CHUNK);
        for (int sidx = 0; sidx < cnt; ++sidx) {
#define QW_ANY_STAGE(K, TIDX, WA, WB, DWA, DWB)                                         \
            {                                                                           \
                const double2 f = sh.fac[3 * sidx + TIDX];                              \
                const double xfr = (K == 0) ? s.yr[0] : s.ur[0];                        \
                const double xfi = (K == 0) ? s.yi[0] : s.ui[0];                        \
                const double xlr = (K == 0) ? (full ? s.yr[R - 1] : s.yr[R > 1 ? R - 2 : 0]) \
                                            : (full ? s.ur[R - 1] : s.ur[R > 1 ? R - 2 : 0]); \
                const double xli = (K == 0) ? (full ? s.yi[R - 1] : s.yi[R > 1 ? R - 2 : 0]) \
                                            : (full ? s.ui[R - 1] : s.ui[R > 1 ? R - 2 : 0]); \
                double2 hl = make_double2(__shfl_sync(full_mask, xlr, left),            \
                                          __shfl_sync(full_mask, xli, left));           \
                double2 hr = make_double2(__shfl_sync(full_mask, xfr, right),           \
                                          __shfl_sync(full_mask, xfi, right));          \
                if (MIRROR) {                 /* reflecting ends of the half ring: selects only */ \
                    /* left of w: slot 1 (lane 0 holds >= 2 real sites).  Right of the      */ \
                    /* antipode: odd n the last real site itself, even n the site before it */ \
                    /* (a short lane of R = 2 holds one site: then the left neighbour's)    */ \
                    constexpr int A = R - 1, B = R > 1 ? R - 2 : 0, C = R > 2 ? R - 3 : 0;  \
                    const double f1r = (K == 0) ? s.yr[1] : s.ur[1], f1i = (K == 0) ? s.yi[1] : s.ui[1]; \
                    const double ar = (K == 0) ? s.yr[A] : s.ur[A], ai = (K == 0) ? s.yi[A] : s.ui[A]; \
                    const double br = (K == 0) ? s.yr[B] : s.ur[B], bi = (K == 0) ? s.yi[B] : s.ui[B]; \
                    const double cr = (R > 2) ? ((K == 0) ? s.yr[C] : s.ur[C]) : hl.x;      \
                    const double ci = (R > 2) ? ((K == 0) ? s.yi[C] : s.ui[C]) : hl.y;      \
                    const double er = even ? (full ? br : cr) : (full ? ar : br);           \
                    const double ei = even ? (full ? bi : ci) : (full ? ai : bi);           \
                    hr = make_double2(m_last ? er : hr.x, m_last ? ei : hr.y);              \
                    hl = make_double2(m_first ? f1r : hl.x, m_first ? f1i : hl.y);          \
                }                                                                       \
                if (!full) {                  /* ghost := the right neighbour's value */ \
                    if (K == 0) { s.yr[R - 1] = hr.x; s.yi[R - 1] = hr.y; }             \
                    else { s.ur[R - 1] = hr.x; s.ui[R - 1] = hr.y; }                    \
                }                                                                       \
                const double2 dd = make_double2((DWA) * f.y, (DWB) * f.y);              \
                rk_stage<R, K, true>(s, (WA) * f.x, (WB) * f.x, hl, hr, zero2, 0.0, own, dd); \
            }
            QW_ANY_STAGE(0, 0, a_half, a_6, d_half, d_6)
            QW_ANY_STAGE(1, 1, a_half, a_3, d_half, d_3)
            QW_ANY_STAGE(2, 1, a_full, a_3, d_full, d_3)
            QW_ANY_STAGE(3, 2, 0.0, a_6, 0.0, d_6)
#undef QW_ANY_STAGE
        }
    }

    // epilogue: ||psi||^2 over the real slots, summed over the L lanes of each point
    double nrm = 0.0;
#pragma unroll
    for (int r = 0; r < R - 1; ++r) nrm = fma(s.yr[r], s.yr[r], fma(s.yi[r], s.yi[r], nrm));
    if (full) nrm = fma(s.yr[R - 1], s.yr[R - 1], fma(s.yi[R - 1], s.yi[R - 1], nrm));
    if (MIRROR) {          // every site of the half ring counts twice, except w and (even n) the antipode
        nrm *= 2.0;
        if (m_first) nrm -= fma(s.yr[0], s.yr[0], s.yi[0] * s.yi[0]);
        if (m_last && even) {
            const double ar = full ? s.yr[R - 1] : s.yr[R > 1 ? R - 2 : 0];
            const double ai = full ? s.yi[R - 1] : s.yi[R > 1 ? R - 2 : 0];
            nrm -= fma(ar, ar, ai * ai);
        }
    }
    double tot = nrm;
    const int gbase = lane - li;
    for (int k = 1; k < L; ++k) {
        int src = li + k;
        if (src >= L) src -= L;
        tot += __shfl_sync(full_mask, nrm, lane_used ? gbase + src : lane);
    }
    if (own && valid) {
        prm.p[pt.q] = fma(s.yr[0], s.yr[0], s.yi[0] * s.yi[0]) / tot;
        if (prm.norm) prm.norm[pt.q] = tot;
    }
    if (prm.psi && valid) {              // undo the relabelling: slot r of lane li -> w + off(li) + r
        double2 *out = prm.psi + (size_t)pt.q * (size_t)prm.n;
        const int64_t off = (int64_t)li * (R - 1) + (li < nfull ? li : nfull);
#pragma unroll
        for (int r = 0; r < R; ++r) {
            if (r == R - 1 && !full) continue;
            int64_t j = off + r + prm.target;
            if (j >= prm.n) j -= prm.n;
            out[j] = make_double2(s.yr[r], s.yi[r]);
            if (MIRROR) {                                         // the mirror image w - l
                int64_t jm = prm.target - (off + r);
                if (jm < 0) jm += prm.n;
                out[jm] = make_double2(s.yr[r], s.yi[r]);
            }
        }
    }
}

template <int R, int MINB, bool MIRROR = false>
__global__ void __launch_bounds__(32, MINB) rk4_cycle_packed_any(const QwParams prm, const int L,
                                                                 const int nfull)
{
    packed_any_body<R, MIRROR>(prm, blockIdx.x, L, nfull);
}

#ifndef QW_ANY4_MINB
#define QW_ANY4_MINB 16
#endif
#ifndef QW_ANY2_MINB
#define QW_ANY2_MINB 16
#endif
template <int R, int MINB, bool MIRROR>
__global__ void __launch_bounds__(32, MINB) rk4_cycle_packed_any_multi(const PackedJob *jobs,
                                                                       const int n_jobs)
{
    __shared__ PackedJob job;
    const int64_t w = packed_job_fetch(jobs, n_jobs, &job);
    packed_any_body<R, MIRROR>(job.prm, w, job.a, job.b);
}

// ---------------------------------------------------------------------------------------------
// The same packed layout for small COMPLETE graphs (4 <= n < 256; the block-per-point kernels
// need a block per point there and leave most lanes of it idle).  (L x)_j = n x_j - sum(x); the
// sum of a stage input is never reduced over the lanes: the column sums of L vanish, so only
// the oracle term changes it,
//   S(u_{k+1}) = S(y) - i DA_k x_w^(k),   S(y+) = S(y) - i sum_k DB_k x_w^(k),
// two FMAs per stage and sum in the lane that owns the marked vertex, broadcast to the lanes of
// the point by one shuffle per component.  S(y) starts at n / sqrt(n) and is not re-anchored:
// the difference between the tracked and the true sum is multiplied by the step polynomial at
// the top eigenvalue of L in every step, |.| <= 1 inside the stability region (qw_horner.cu).
// The ghost slot of a short lane evolves like a site of its own and is never read.
template <int R, int MINB>
__global__ void __launch_bounds__(32, MINB) rk4_complete_packed_any(const QwParams prm, const int L,
                                                                    const int nfull)
{
    __shared__ PackedShared sh;
    const unsigned full_mask = 0xffffffffu;
    const int lane = threadIdx.x;
    const int groups = 32 / L;
    const int grp = lane / L, li = lane - grp * L;
    const bool lane_used = grp < groups;
    bool valid;
    const QwPoint pt = packed_point(prm, blockIdx.x, lane_used ? grp : groups - 1, groups, valid);
    valid = valid && lane_used;
    __shared__ int64_t steps_slot;
    const int64_t S = packed_steps(prm, blockIdx.x, groups, &steps_slot);
    const bool own = lane_used && li == 0;
    const bool full = !lane_used || li < nfull;        // R sites (idle lanes: R zeros)
    const int gbase = lane_used ? lane - li : lane;    // the lane that owns the marked vertex

    const double h = pt.h;                       // 0 when T == 0: the state never moves
    const double gl = prm.gamma_on == QW_GAMMA_ON_LAPLACIAN ? pt.gamma : 1.0;
    const double gd = prm.gamma_on == QW_GAMMA_ON_ORACLE ? pt.gamma : 1.0;
    const double a_half = 0.5 * h * gl, a_full = h * gl, a_6 = h / 6.0 * gl, a_3 = h / 3.0 * gl;
    const double d_half = 0.5 * h * gd, d_full = h * gd, d_6 = h / 6.0 * gd, d_3 = h / 3.0 * gd;
    const double nd = lane_used ? (double)prm.n : 0.0;

    State<R> s;
    const double amp = lane_used ? 1.0 / sqrt((double)prm.n) : 0.0;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        s.yr[r] = amp; s.yi[r] = 0.0;
        s.ur[r] = 0.0; s.ui[r] = 0.0; s.ar[r] = 0.0; s.ai[r] = 0.0;
    }
    double2 Sy = make_double2(nd * amp, 0.0);    // sum of psi_n over the point's sites
    double2 Sx = Sy, Sa = Sy;                    // sum of the stage input, of the accumulator
    const double2 zero2 = make_double2(0.0, 0.0);
    const double inv_s = 1.0 / (double)S;

    for (int64_t n0 = 0; n0 < S; n0 += QW_CHUNK) {
        __syncwarp();
        for (int i = lane; i < 3 * QW_CHUNK; i += 32) {
            const int st = i / 3, k = i - 3 * st;
            sh.fac[i] = packed_factor(prm, pt, n0 + st, k, inv_s);
        }
        __syncwarp();
        const int cnt = (int)((S - n0 < QW_CHUNK) ? (S - n0) : QW_CHUNK);
        for (int sidx = 0; sidx < cnt; ++sidx) {
#define QW_CPL_STAGE(K, TIDX, WA, WB, DWA, DWB)                                         \
            {                                                                           \
                const double2 f = sh.fac[3 * sidx + TIDX];                              \
                const double x0r = (K == 0) ? s.yr[0] : s.ur[0];                        \
                const double x0i = (K == 0) ? s.yi[0] : s.ui[0];                        \
                const double2 dd = make_double2((DWA) * f.y, (DWB) * f.y);              \
                rk_stage<R, K, false>(s, (WA) * f.x, (WB) * f.x, zero2, zero2, Sx, nd, own, dd); \
                const double bx = (K == 0) ? Sy.x : Sa.x, by = (K == 0) ? Sy.y : Sa.y;  \
                Sa = make_double2(fma(dd.y, x0i, bx), fma(-dd.y, x0r, by));             \
                const double2 nx = (K < 3) ? make_double2(fma(dd.x, x0i, Sy.x),         \
                                                          fma(-dd.x, x0r, Sy.y)) : Sa;  \
                Sx = make_double2(__shfl_sync(full_mask, nx.x, gbase),                  \
                                  __shfl_sync(full_mask, nx.y, gbase));                 \
                if (K == 3) Sy = Sx;                                                    \
            }
            QW_CPL_STAGE(0, 0, a_half, a_6, d_half, d_6)
            QW_CPL_STAGE(1, 1, a_half, a_3, d_half, d_3)
            QW_CPL_STAGE(2, 1, a_full, a_3, d_full, d_3)
            QW_CPL_STAGE(3, 2, 0.0, a_6, 0.0, d_6)
#undef QW_CPL_STAGE
        }
    }

    // epilogue: ||psi||^2 over the real slots, summed over the L lanes of each point
    double nrm = 0.0;
#pragma unroll
    for (int r = 0; r < R - 1; ++r) nrm = fma(s.yr[r], s.yr[r], fma(s.yi[r], s.yi[r], nrm));
    if (full) nrm = fma(s.yr[R - 1], s.yr[R - 1], fma(s.yi[R - 1], s.yi[R - 1], nrm));
    double tot = nrm;
    for (int k = 1; k < L; ++k) {
        int src = li + k;
        if (src >= L) src -= L;
        tot += __shfl_sync(full_mask, nrm, lane_used ? gbase + src : lane);
    }
    if (own && valid) {
        prm.p[pt.q] = fma(s.yr[0], s.yr[0], s.yi[0] * s.yi[0]) / tot;
        if (prm.norm) prm.norm[pt.q] = tot;
    }
    if (prm.psi && valid) {              // undo the relabelling: slot r of lane li -> w + off(li) + r
        double2 *out = prm.psi + (size_t)pt.q * (size_t)prm.n;
        const int64_t off = (int64_t)li * (R - 1) + (li < nfull ? li : nfull);
#pragma unroll
        for (int r = 0; r < R; ++r) {
            if (r == R - 1 && !full) continue;
            int64_t j = off + r + prm.target;
            if (j >= prm.n) j -= prm.n;
            out[j] = make_double2(s.yr[r], s.yi[r]);
        }
    }
}

template <int R, int MINB>
cudaError_t launch_packed_complete(const QwParams &prm, cudaStream_t st, qw_plan *plan)
{
    const int L = (int)((prm.n + R - 1) / R);
    const int nfull = (int)(prm.n - (int64_t)L * (R - 1));
    const int groups = 32 / L;
    if (plan) {
        cudaFuncAttributes fa;
        cudaError_t e = cudaFuncGetAttributes(&fa, rk4_complete_packed_any<R, MINB>);
        if (e != cudaSuccess) return e;
        int nb = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, rk4_complete_packed_any<R, MINB>, 32,
                                                          0);
        if (e != cudaSuccess) return e;
        plan->path = 11;
        plan->sites_per_thread = R;
        plan->threads_per_block = 32;
        plan->blocks_per_sm = nb;
        plan->regs_per_thread = fa.numRegs;
        plan->smem_bytes = (int)fa.sharedSizeBytes;
        plan->steps_per_pass = groups;          // points per warp
        return cudaSuccess;
    }
    const int64_t blocks = packed_warps(prm, groups);
    rk4_complete_packed_any<R, MINB><<<(unsigned)blocks, 32, 0, st>>>(prm, L, nfull);
    return cudaGetLastError();
}

// sites of a point that the any-dimension kernel integrates: the ring, or the half ring
// w .. w + n/2 under QW_FLAG_MIRROR
static int64_t packed_any_sites(const QwParams &prm, bool mirror)
{
    return mirror ? prm.n / 2 + 1 : prm.n;
}

template <int R, int MINB, bool MIRROR>
cudaError_t launch_packed_any(const QwParams &prm, cudaStream_t st, qw_plan *plan)
{
    const int64_t m = packed_any_sites(prm, MIRROR);
    const int L = (int)((m + R - 1) / R);
    const int nfull = (int)(m - (int64_t)L * (R - 1));
    const int groups = 32 / L;
    if (plan) {
        cudaFuncAttributes fa;
        cudaError_t e = cudaFuncGetAttributes(&fa, rk4_cycle_packed_any<R, MINB, MIRROR>);
        if (e != cudaSuccess) return e;
        int nb = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, rk4_cycle_packed_any<R, MINB, MIRROR>,
                                                          32, 0);
        if (e != cudaSuccess) return e;
        plan->path = MIRROR ? 13 : 8;
        plan->sites_per_thread = R;
        plan->threads_per_block = 32;
        plan->blocks_per_sm = nb;
        plan->regs_per_thread = fa.numRegs;
        plan->smem_bytes = (int)fa.sharedSizeBytes;
        plan->steps_per_pass = groups;          // points per warp
        return cudaSuccess;
    }
    const int64_t blocks = packed_warps(prm, groups);
    rk4_cycle_packed_any<R, MINB, MIRROR><<<(unsigned)blocks, 32, 0, st>>>(prm, L, nfull);
    return cudaGetLastError();
}

// sites per lane of the any-dimension packed kernel: the largest R in {8, 4, 2} with n >= R^2
// whose ceil(n / R) lanes fit a warp; 0 if none (n < 4 or n > 256)
static int packed_any_r(int64_t n)
{
    const int cand[3] = {8, 4, 2};
    for (int i = 0; i < 3; ++i) {
        const int R = cand[i];
        if (n >= R * R && (n + R - 1) / R <= 32) return R;
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------
// Nested (Horner) form of the step for the packed layout (qw_horner_level.cuh):
// N = 16 * TPP with TPP = 8, 16 or 32 lanes per point (N = 128, 256, 512; above that the
// block-per-point kernel of qw_horner.cu takes over).  24 instead of 30 FP64 instructions per
// site-update and two instead of three complex vectors in registers (12 instead of 8 warps
// per SM).  The step coefficients depend on the point (h, gamma), so every point of the warp
// has its own table: a refill computes 32 coefficient sets -- 32 / G steps for each of the
// G = 32 / TPP points -- one per lane.  The marked vertex is slot 3 of the first lane of a
// point's group; eight lanes of the group evaluate its four corrections at once.
// Oracle correction (round 2, as the one-warp form of qw_horner.cu): the first lane of a group
// publishes its window y_{w-3} .. y_w in shared memory (mirror symmetry: qw_horner_level.cuh), the
// group's TPP lanes form the four dot products y_w + sigma_l -- lane 8 qg + l the terms of 32 / TPP
// window sites in output l, then a butterfly over qg -- and the owner takes them as the ADDEND of
// its marked slot: 4 / 5 / 8 FP64 warp-instructions per step for 32 / 16 / 8 lanes per point
// instead of 22, two shuffle stages instead of 24 shuffles.
template <int MINB, int TPP_LOG2>
__global__ void __launch_bounds__(32, MINB) rk4_cycle_packed_horner(const QwParams prm)
{
    constexpr int R = 16;
    constexpr int tpp_log2 = TPP_LOG2;
    constexpr int QPL = 32 >> TPP_LOG2;          // sums s_q per lane: 1, 2, 4
    __shared__ HornerStepBase tab[32];           // [group][step of the chunk]
    __shared__ double2 win[4][4];                // [group]: slots 0 .. 3 of the group's first lane
    __shared__ double2 tq2[4][4];                // [group]: y_w + sigma_4 .. y_w + sigma_1
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x;
    constexpr int tpp = 1 << tpp_log2, groups = 32 >> tpp_log2, ch = 32 / groups;   // ch = tpp
    const int li = lane & (tpp - 1), grp = lane >> tpp_log2;
    bool valid;
    const QwPoint pt = packed_point(prm, blockIdx.x, grp, groups, valid);
    __shared__ int64_t steps_slot;
    const int64_t S = packed_steps(prm, blockIdx.x, groups, &steps_slot);
    const int left = (li == 0) ? lane + tpp - 1 : lane - 1;
    const int right = (li == tpp - 1) ? lane - tpp + 1 : lane + 1;
    const bool own = li == 0;
    const uint32_t a_win = smem_u32(&win[grp][0]), a_tq = smem_u32(&tq2[grp][0]);
    // the lane's refill duty: step (lane % ch) of point (lane / ch) of this warp
    const int fgrp = lane / ch, fstep = lane - fgrp * ch;
    bool fvalid;
    const QwPoint fpt = packed_point(prm, blockIdx.x, fgrp, groups, fvalid);

    HState<R> s;
    const double amp = 1.0 / sqrt((double)prm.n);
#pragma unroll
    for (int r = 0; r < R; ++r) {
        s.yr[r] = amp; s.yi[r] = 0.0;
        s.zr[r] = 0.0; s.zi[r] = 0.0;
    }

    if (own) {
#pragma unroll
        for (int r = 0; r <= HW0; ++r) win[grp][r] = make_double2(s.yr[r], s.yi[r]);
    }
    for (int64_t n0 = 0; n0 < S; n0 += ch) {
        __syncwarp();
        if (n0 + fstep < S && fpt.steps == 0) {        // T == 0: identity step (BCB.py:72-73)
            HornerStepBase &o = tab[lane];
#pragma unroll
            for (int k = 0; k < 4; ++k) o.rho[k] = 0.0;
#pragma unroll
            for (int k = 0; k < 32; ++k)           // y_w + 0: the identity on y_w
                o.coef2[k] = make_double2(k < 8 && !(k & 1) ? 1.0 : 0.0, k < 8 && (k & 1) ? 1.0 : 0.0);
        } else if (n0 + fstep < S) {
            const double t0 = (double)(n0 + fstep) * fpt.h;
            const double2 ad1 = qw_operator_scalars(t0 / fpt.T, fpt.gamma, prm.schedule, prm.gamma_on);
            const double2 ad2 = qw_operator_scalars((t0 + 0.5 * fpt.h) / fpt.T, fpt.gamma,
                                                    prm.schedule, prm.gamma_on);
            const double2 ad4 = qw_operator_scalars((t0 + fpt.h) / fpt.T, fpt.gamma, prm.schedule,
                                                    prm.gamma_on);
            HornerCoef hc;
            horner_coefficients(hc, fpt.h, ad1, ad2, ad4, 2.0, 6.0);
            horner_cycle_sum_basis(hc.c);
            horner_step_store(tab[lane], hc, true);
        }
        __syncwarp();
        const int cnt = (int)((S - n0 < ch) ? (S - n0) : ch);
        for (int sidx = 0; sidx < cnt; ++sidx) {
            const HornerStepBase &hs = tab[grp * ch + sidx];
#define QW_PH_LEVEL(LEVEL)                                                                  \
            {                                                                               \
                const double xfr = (LEVEL == 4) ? s.yr[0] : s.zr[0];                        \
                const double xfi = (LEVEL == 4) ? s.yi[0] : s.zi[0];                        \
                const double xlr = (LEVEL == 4) ? s.yr[R - 1] : s.zr[R - 1];                \
                const double xli = (LEVEL == 4) ? s.yi[R - 1] : s.zi[R - 1];                \
                const double2 hl = make_double2(__shfl_sync(full, xlr, left),               \
                                                __shfl_sync(full, xli, left));              \
                const double2 hr = make_double2(__shfl_sync(full, xfr, right),              \
                                                __shfl_sync(full, xfi, right));             \
                const double rho = hs.rho[LEVEL - 1];                                       \
                if (LEVEL == 4) {                 /* the window was published at level 1 */ \
                    const int qg = li >> 3, l = li & 7;                                     \
                    double v = 0.0;                                                         \
                    _Pragma("unroll")                                                       \
                    for (int j = 0; j < QPL; ++j) {                                         \
                        const int q = qg * QPL + j;                                         \
                        const double2 ya = lds2(a_win + 16 * (HW0 - q));                    \
                        const double2 cf = hs.coef2[8 * q + l];                             \
                        v = fma(cf.y, ya.y, fma(cf.x, ya.x, v));                            \
                    }                                                                       \
                    if (TPP_LOG2 >= 4) v += __shfl_xor_sync(full, v, 8);                    \
                    if (TPP_LOG2 >= 5) v += __shfl_xor_sync(full, v, 16);                   \
                    if (li < 8) reinterpret_cast<double *>(&tq2[grp][0])[li] = v;           \
                    __syncwarp();                                                           \
                }                                                                           \
                const double2 tw = lds2(a_tq + 16 * (4 - LEVEL));                           \
                const double wr = own ? tw.x : s.yr[HW0], wi = own ? tw.y : s.yi[HW0];      \
                double2 o0, oL;                                                             \
                hedges<R, LEVEL, HW0, HW0>(s, rho, hl, hr, o0, oL, wr, wi, a_win, (int)own); \
                hinterior<R, LEVEL, HW0, HW0>(s, rho, o0, oL, wr, wi, a_win, (int)own);     \
                if (LEVEL == 1) {                 /* y[HW0] completes the next step's window */ \
                    sts2_if(a_win + 16 * HW0, s.yr[HW0], s.yi[HW0], (int)own);              \
                    __syncwarp();                                                           \
                }                                                                           \
            }
            QW_PH_LEVEL(4)
            QW_PH_LEVEL(3)
            QW_PH_LEVEL(2)
            QW_PH_LEVEL(1)
#undef QW_PH_LEVEL
        }
    }

    // epilogue: segmented reductions over the tpp lanes of each point
    double nrm = 0.0;
#pragma unroll
    for (int r = 0; r < R; ++r) nrm = fma(s.yr[r], s.yr[r], fma(s.yi[r], s.yi[r], nrm));
    for (int o = tpp >> 1; o > 0; o >>= 1) nrm += __shfl_xor_sync(full, nrm, o);
    if (li == 0 && valid) {
        prm.p[pt.q] = fma(s.yr[HW0], s.yr[HW0], s.yi[HW0] * s.yi[HW0]) / nrm;
        if (prm.norm) prm.norm[pt.q] = nrm;
    }
    if (prm.psi && valid) {              // undo the relabelling: slot r of lane li -> w + li R + r - 3
        double2 *out = prm.psi + (size_t)pt.q * (size_t)prm.n;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            int64_t j = (int64_t)li * R + r - HW0 + prm.target;
            if (j < 0) j += prm.n;
            if (j >= prm.n) j -= prm.n;
            out[j] = make_double2(s.yr[r], s.yi[r]);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Mirror-reduced nested kernel (QW_FLAG_MIRROR, opt-in).  The reference always starts from the
// uniform state (BCB.py:179-180) and H(t) is invariant under the reflection of the ring about
// the marked vertex, so psi_{w+k} = psi_{w-k} at all times -- bit for bit in the full-ring
// kernels, whose neighbour sums are commutative.  Integrating the half ring w .. w + n/2 with
// reflecting ends (left of w: psi_{w+1}; right of the antipode: psi_{n/2-1} for even n, the
// last site itself for odd n) gives the same numbers with half the arithmetic, and for
// n <= 1086 the n/2 + 1 sites of a point fit the registers of ONE warp (17 sites per lane):
// halo values travel by shuffles, there is no mailbox and no barrier.  s_k = 2 psi_{w+k} in the
// correction basis; ||psi||^2 counts the interior sites twice.  One point per warp, any step
// rule.  Not the default: the benchmark's unit is the full-ring site-update.
template <int R, int MINB>
__global__ void __launch_bounds__(32, MINB) rk4_cycle_mirror(const QwParams prm, const int L,
                                                             const int nfull)
{
    __shared__ HornerStepBase tab[32];            // 12 blocks per SM: 18 KB of shared memory each
    __shared__ double2 win[4], priv[32];          // lane 0's slots 1 .. 3; every lane's y[0]
    __shared__ double tq[8];                      // y_w + sigma_4 .. y_w + sigma_1 (re, im)
    __shared__ int64_t steps_slot;
    const unsigned fm = 0xffffffffu;
    const int lane = threadIdx.x;
    const QwPoint pt = qw_point(prm, blockIdx.x);
    if (lane == 0) steps_slot = pt.steps;         // uniform to the compiler (see packed_steps)
    __syncwarp();
    const int64_t S = steps_slot;
    const bool active = lane < L;
    const bool full = lane < nfull || !active;
    const bool first = lane == 0, lastl = lane == L - 1;
    const bool even = (prm.n & 1) == 0;
    const int left = first ? lane : lane - 1, right = lane + 1 < 32 ? lane + 1 : lane;
    // Oracle correction as in the one-warp form of qw_horner.cu (HR_SELF): the 32 lanes form the
    // four dot products from lane 0's window in shared memory (6 FP64 warp-instructions), lane 0
    // takes y_w + sigma_l as the addend of slot 0, every other lane its own y[0] -- from shared
    // memory, so that no select is needed (round 1: 22 instructions and 36 shuffles per step).
    const int q = lane >> 3;
    const uint32_t wq = q == 0 ? smem_u32(priv) : smem_u32(win + q);   // y_{w-q} = y_{w+q} = slot q
    const uint32_t a_priv = smem_u32(priv + lane);
    const uint32_t a_t = first ? smem_u32(tq) : a_priv, a_ts = first ? 16u : 0u;
    const uint32_t a_win = smem_u32(win);

    HState<R> s;
    const double amp = active ? 1.0 / sqrt((double)prm.n) : 0.0;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        s.yr[r] = (r == R - 1 && !full) ? 0.0 : amp; s.yi[r] = 0.0;
        s.zr[r] = 0.0; s.zi[r] = 0.0;
    }

    priv[lane] = make_double2(s.yr[0], s.yi[0]);
    if (first) {
#pragma unroll
        for (int r = 1; r < 4; ++r) win[r] = make_double2(s.yr[r], s.yi[r]);
    }
    for (int64_t n0 = 0; n0 < S; n0 += 32) {
        __syncwarp();
        if (n0 + lane < S) {
            const double t0 = (double)(n0 + lane) * pt.h;
            const double2 ad1 = qw_operator_scalars(t0 / pt.T, pt.gamma, prm.schedule, prm.gamma_on);
            const double2 ad2 = qw_operator_scalars((t0 + 0.5 * pt.h) / pt.T, pt.gamma,
                                                    prm.schedule, prm.gamma_on);
            const double2 ad4 = qw_operator_scalars((t0 + pt.h) / pt.T, pt.gamma, prm.schedule,
                                                    prm.gamma_on);
            HornerCoef hc;
            horner_coefficients(hc, pt.h, ad1, ad2, ad4, 2.0, 6.0);
            horner_cycle_sum_basis(hc.c);
            horner_step_store(tab[lane], hc, true);
        }
        __syncwarp();
        const int cnt = (int)((S - n0 < 32) ? (S - n0) : 32);
        for (int sidx = 0; sidx < cnt; ++sidx) {
            const HornerStepBase &hs = tab[sidx];
#define QW_MX_R(LV, r) ((LV == 4) ? s.yr[r] : s.zr[r])
#define QW_MX_I(LV, r) ((LV == 4) ? s.yi[r] : s.zi[r])
#define QW_MIRROR_LEVEL(LEVEL)                                                              \
            {                                                                               \
                /* last real slot (a short lane's slot R-1 is a ghost) and the one before */ \
                const double lr = full ? QW_MX_R(LEVEL, R - 1) : QW_MX_R(LEVEL, R - 2);                   \
                const double li_ = full ? QW_MX_I(LEVEL, R - 1) : QW_MX_I(LEVEL, R - 2);                  \
                const double br = full ? QW_MX_R(LEVEL, R - 2) : QW_MX_R(LEVEL, R - 3);                   \
                const double bi = full ? QW_MX_I(LEVEL, R - 2) : QW_MX_I(LEVEL, R - 3);                   \
                double2 hl = make_double2(__shfl_sync(fm, lr, left), __shfl_sync(fm, li_, left)); \
                double2 hr = make_double2(__shfl_sync(fm, QW_MX_R(LEVEL, 0), right),               \
                                          __shfl_sync(fm, QW_MX_I(LEVEL, 0), right));              \
                if (first) hl = make_double2(QW_MX_R(LEVEL, 1), QW_MX_I(LEVEL, 1));     /* mirror about w */ \
                if (lastl) hr = even ? make_double2(br, bi) : make_double2(lr, li_);        \
                if (!full) {                                                                \
                    if (LEVEL == 4) { s.yr[R - 1] = hr.x; s.yi[R - 1] = hr.y; }             \
                    else { s.zr[R - 1] = hr.x; s.zi[R - 1] = hr.y; }                        \
                }                                                                           \
                const double rho = hs.rho[LEVEL - 1];                                       \
                if (LEVEL == 4) {                                                           \
                    const double v = horner_window_dot(wq, hs, lane);                   \
                    if (lane < 8) tq[lane] = v;                                             \
                    __syncwarp();                                                           \
                }                                                                           \
                const double2 tw = lds2(a_t + (4 - LEVEL) * a_ts);                          \
                double2 o0, oL;                                                             \
                hedges<R, LEVEL, 0, 4>(s, rho, hl, hr, o0, oL, tw.x, tw.y, a_win, (int)first); \
                hinterior<R, LEVEL, 0, 4>(s, rho, o0, oL, tw.x, tw.y, a_win, (int)first);   \
                if (LEVEL == 1) {                 /* the next step's y[0]; window complete */ \
                    sts2(a_priv, s.yr[0], s.yi[0]);                                         \
                    __syncwarp();                                                           \
                }                                                                           \
            }
            QW_MIRROR_LEVEL(4)
            QW_MIRROR_LEVEL(3)
            QW_MIRROR_LEVEL(2)
            QW_MIRROR_LEVEL(1)
#undef QW_MIRROR_LEVEL
#undef QW_MX_R
#undef QW_MX_I
        }
    }

    // ||psi||^2: every site of the half ring counts twice, except w and (even n) the antipode
    double nrm = 0.0;
#pragma unroll
    for (int r = 0; r < R - 1; ++r) nrm = fma(s.yr[r], s.yr[r], fma(s.yi[r], s.yi[r], nrm));
    if (full) nrm = fma(s.yr[R - 1], s.yr[R - 1], fma(s.yi[R - 1], s.yi[R - 1], nrm));
    nrm *= 2.0;
    const double pw = fma(s.yr[0], s.yr[0], s.yi[0] * s.yi[0]);
    if (first) nrm -= pw;
    if (lastl && even) {
        const double ar = full ? s.yr[R - 1] : s.yr[R - 2], ai = full ? s.yi[R - 1] : s.yi[R - 2];
        nrm -= fma(ar, ar, ai * ai);
    }
    if (!active) nrm = 0.0;
    nrm = qw_warp_sum(nrm);
    if (first) {
        prm.p[pt.q] = pw / nrm;
        if (prm.norm) prm.norm[pt.q] = nrm;
    }
    if (prm.psi && active) {              // both mirror images: w + l and w - l
        double2 *out = prm.psi + (size_t)pt.q * (size_t)prm.n;
        const int64_t off = (int64_t)lane * (R - 1) + (lane < nfull ? lane : nfull);
#pragma unroll
        for (int r = 0; r < R; ++r) {
            if (r == R - 1 && !full) continue;
            int64_t a = prm.target + off + r, b = prm.target - off - r;
            if (a >= prm.n) a -= prm.n;
            if (b < 0) b += prm.n;
            out[a] = make_double2(s.yr[r], s.yi[r]);
            out[b] = make_double2(s.yr[r], s.yi[r]);
        }
    }
}

template <int R, int MINB>
cudaError_t launch_mirror(const QwParams &prm, cudaStream_t st, qw_plan *plan)
{
    const int64_t m = prm.n / 2 + 1;                 // sites of the half ring
    const int L = (int)((m + R - 1) / R);
    const int nfull = (int)(m - (int64_t)L * (R - 1));
    if (plan) {
        cudaFuncAttributes fa;
        cudaError_t e = cudaFuncGetAttributes(&fa, rk4_cycle_mirror<R, MINB>);
        if (e != cudaSuccess) return e;
        int nb = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, rk4_cycle_mirror<R, MINB>, 32, 0);
        if (e != cudaSuccess) return e;
        plan->path = 9;
        plan->sites_per_thread = R;
        plan->threads_per_block = 32;
        plan->blocks_per_sm = nb;
        plan->regs_per_thread = fa.numRegs;
        plan->smem_bytes = (int)fa.sharedSizeBytes;
        plan->steps_per_pass = 1;
        return cudaSuccess;
    }
    rk4_cycle_mirror<R, MINB><<<(unsigned)prm.n_points, 32, 0, st>>>(prm, L, nfull);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// The same half-ring kernel for half rings that fill at most half a warp: LP = 8 or 16 lanes per
// point, G = 4 or 2 points per warp (BASELINE config 4, N = 256: 129 sites = 8 lanes of 17 / 16
// sites; one point per warp left 17 of 32 lanes idle and lost to the FULL-ring nested kernel,
// 29.9 against 22.5 ms for 1e5 points).  As in rk4_cycle_packed_horner every point of the warp
// has its own coefficient table -- a refill computes 32 / G steps for each of the G points, one
// set per lane -- its own window (slots 1 .. 3 of the group's first lane; slot 0 in priv) and its
// own four corrections: lane 8 qg + l of a group sums the terms of QPL window sites in output l,
// a butterfly over qg follows for 16 lanes per point.  All points of a warp share the step
// count (qw_packed_supported: the n_steps rule, or a heatmap column under a fixed step size).
template <int R, int MINB, int LLOG2>
__global__ void __launch_bounds__(32, MINB) rk4_cycle_mirror_packed(const QwParams prm,
                                                                    const int L, const int nfull)
{
    constexpr int LP = 1 << LLOG2, G = 32 >> LLOG2, CH = LP;     // CH steps per refill and point
    constexpr int QPL = 32 >> LLOG2;                             // window terms per lane: 4, 2
    __shared__ HornerStepBase tab[32];            // [group][step of the chunk]
    __shared__ double2 win[G][4], priv[32];       // slots 1 .. 3 of a group's first lane; every lane's y[0]
    __shared__ double tq[G][8];                   // y_w + sigma_4 .. y_w + sigma_1 (re, im)
    const unsigned fm = 0xffffffffu;
    const int lane = threadIdx.x;
    const int li = lane & (LP - 1), grp = lane >> LLOG2;
    bool valid;
    const QwPoint pt = packed_point(prm, blockIdx.x, grp, G, valid);
    // the step count passes through tq (written next inside the time loop, behind a __syncwarp):
    // with four points per warp the block must not exceed 18432 bytes for 12 blocks per SM
    const int64_t S = packed_steps(prm, blockIdx.x, G, reinterpret_cast<int64_t *>(&tq[0][0]));
    const bool active = li < L;
    const bool full = li < nfull || !active;
    const bool first = li == 0, lastl = li == L - 1;
    const bool even = (prm.n & 1) == 0;
    const int left = first ? lane : lane - 1, right = lane + 1 < 32 ? lane + 1 : lane;
    // the lane's terms of the four dot products: window site q = qg QPL + j in output l
    const int qg = li >> 3, l = li & 7;
    uint32_t wq[QPL];
#pragma unroll
    for (int j = 0; j < QPL; ++j) {
        const int q = qg * QPL + j;
        wq[j] = q == 0 ? smem_u32(priv + (lane - li)) : smem_u32(&win[grp][q]);
    }
    const uint32_t a_priv = smem_u32(priv + lane);
    const uint32_t a_t = first ? smem_u32(&tq[grp][0]) : a_priv, a_ts = first ? 16u : 0u;
    const uint32_t a_win = smem_u32(&win[grp][0]);
    // the lane's refill duty: step (lane % CH) of point (lane / CH) of this warp
    const int fgrp = lane / CH, fstep = lane - fgrp * CH;
    bool fvalid;
    const QwPoint fpt = packed_point(prm, blockIdx.x, fgrp, G, fvalid);

    HState<R> s;
    const double amp = active ? 1.0 / sqrt((double)prm.n) : 0.0;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        s.yr[r] = (r == R - 1 && !full) ? 0.0 : amp; s.yi[r] = 0.0;
        s.zr[r] = 0.0; s.zi[r] = 0.0;
    }

    priv[lane] = make_double2(s.yr[0], s.yi[0]);
    if (first) {
#pragma unroll
        for (int r = 1; r < 4; ++r) win[grp][r] = make_double2(s.yr[r], s.yi[r]);
    }
    for (int64_t n0 = 0; n0 < S; n0 += CH) {
        __syncwarp();
        if (n0 + fstep < S && fpt.steps == 0) {        // T == 0: identity step (BCB.py:72-73)
            HornerStepBase &o = tab[lane];
#pragma unroll
            for (int k = 0; k < 4; ++k) o.rho[k] = 0.0;
#pragma unroll
            for (int k = 0; k < 32; ++k)           // y_w + 0: the identity on y_w
                o.coef2[k] = make_double2(k < 8 && !(k & 1) ? 1.0 : 0.0, k < 8 && (k & 1) ? 1.0 : 0.0);
        } else if (n0 + fstep < S) {
            const double t0 = (double)(n0 + fstep) * fpt.h;
            const double2 ad1 = qw_operator_scalars(t0 / fpt.T, fpt.gamma, prm.schedule, prm.gamma_on);
            const double2 ad2 = qw_operator_scalars((t0 + 0.5 * fpt.h) / fpt.T, fpt.gamma,
                                                    prm.schedule, prm.gamma_on);
            const double2 ad4 = qw_operator_scalars((t0 + fpt.h) / fpt.T, fpt.gamma, prm.schedule,
                                                    prm.gamma_on);
            HornerCoef hc;
            horner_coefficients(hc, fpt.h, ad1, ad2, ad4, 2.0, 6.0);
            horner_cycle_sum_basis(hc.c);
            horner_step_store(tab[lane], hc, true);
        }
        __syncwarp();
        const int cnt = (int)((S - n0 < CH) ? (S - n0) : CH);
        for (int sidx = 0; sidx < cnt; ++sidx) {
            const HornerStepBase &hs = tab[grp * CH + sidx];
#define QW_MX_R(LV, r) ((LV == 4) ? s.yr[r] : s.zr[r])
#define QW_MX_I(LV, r) ((LV == 4) ? s.yi[r] : s.zi[r])
#define QW_MIRROR_LEVEL(LEVEL)                                                              \
            {                                                                               \
                /* last real slot (a short lane's slot R-1 is a ghost) and the one before */ \
                const double lr = full ? QW_MX_R(LEVEL, R - 1) : QW_MX_R(LEVEL, R - 2);     \
                const double li_ = full ? QW_MX_I(LEVEL, R - 1) : QW_MX_I(LEVEL, R - 2);    \
                const double br = full ? QW_MX_R(LEVEL, R - 2) : QW_MX_R(LEVEL, R - 3);     \
                const double bi = full ? QW_MX_I(LEVEL, R - 2) : QW_MX_I(LEVEL, R - 3);     \
                double2 hl = make_double2(__shfl_sync(fm, lr, left), __shfl_sync(fm, li_, left)); \
                double2 hr = make_double2(__shfl_sync(fm, QW_MX_R(LEVEL, 0), right),        \
                                          __shfl_sync(fm, QW_MX_I(LEVEL, 0), right));       \
                if (first) hl = make_double2(QW_MX_R(LEVEL, 1), QW_MX_I(LEVEL, 1));     /* mirror about w */ \
                if (lastl) hr = even ? make_double2(br, bi) : make_double2(lr, li_);        \
                if (!full) {                                                                \
                    if (LEVEL == 4) { s.yr[R - 1] = hr.x; s.yi[R - 1] = hr.y; }             \
                    else { s.zr[R - 1] = hr.x; s.zi[R - 1] = hr.y; }                        \
                }                                                                           \
                const double rho = hs.rho[LEVEL - 1];                                       \
                if (LEVEL == 4) {                                                           \
                    double v = 0.0;                                                         \
                    _Pragma("unroll")                                                       \
                    for (int j = 0; j < QPL; ++j) {                                         \
                        const double2 ya = lds2(wq[j]);                                     \
                        const double2 cf = hs.coef2[8 * (qg * QPL + j) + l];                \
                        v = fma(cf.y, ya.y, fma(cf.x, ya.x, v));                            \
                    }                                                                       \
                    if (LLOG2 >= 4) v += __shfl_xor_sync(fm, v, 8);                         \
                    if (li < 8) tq[grp][li] = v;                                            \
                    __syncwarp();                                                           \
                }                                                                           \
                const double2 tw = lds2(a_t + (4 - LEVEL) * a_ts);                          \
                double2 o0, oL;                                                             \
                hedges<R, LEVEL, 0, 4>(s, rho, hl, hr, o0, oL, tw.x, tw.y, a_win, (int)first); \
                hinterior<R, LEVEL, 0, 4>(s, rho, o0, oL, tw.x, tw.y, a_win, (int)first);   \
                if (LEVEL == 1) {                 /* the next step's y[0]; window complete */ \
                    sts2(a_priv, s.yr[0], s.yi[0]);                                         \
                    __syncwarp();                                                           \
                }                                                                           \
            }
            QW_MIRROR_LEVEL(4)
            QW_MIRROR_LEVEL(3)
            QW_MIRROR_LEVEL(2)
            QW_MIRROR_LEVEL(1)
#undef QW_MIRROR_LEVEL
#undef QW_MX_R
#undef QW_MX_I
        }
    }

    // ||psi||^2: every site of the half ring counts twice, except w and (even n) the antipode
    double nrm = 0.0;
#pragma unroll
    for (int r = 0; r < R - 1; ++r) nrm = fma(s.yr[r], s.yr[r], fma(s.yi[r], s.yi[r], nrm));
    if (full) nrm = fma(s.yr[R - 1], s.yr[R - 1], fma(s.yi[R - 1], s.yi[R - 1], nrm));
    nrm *= 2.0;
    const double pw = fma(s.yr[0], s.yr[0], s.yi[0] * s.yi[0]);
    if (first) nrm -= pw;
    if (lastl && even) {
        const double ar = full ? s.yr[R - 1] : s.yr[R - 2], ai = full ? s.yi[R - 1] : s.yi[R - 2];
        nrm -= fma(ar, ar, ai * ai);
    }
    if (!active) nrm = 0.0;
#pragma unroll
    for (int o = LP >> 1; o > 0; o >>= 1) nrm += __shfl_xor_sync(fm, nrm, o);
    if (first && valid) {
        prm.p[pt.q] = pw / nrm;
        if (prm.norm) prm.norm[pt.q] = nrm;
    }
    if (prm.psi && active && valid) {     // both mirror images: w + l and w - l
        double2 *out = prm.psi + (size_t)pt.q * (size_t)prm.n;
        const int64_t off = (int64_t)li * (R - 1) + (li < nfull ? li : nfull);
#pragma unroll
        for (int r = 0; r < R; ++r) {
            if (r == R - 1 && !full) continue;
            int64_t a = prm.target + off + r, b = prm.target - off - r;
            if (a >= prm.n) a -= prm.n;
            if (b < 0) b += prm.n;
            out[a] = make_double2(s.yr[r], s.yi[r]);
            out[b] = make_double2(s.yr[r], s.yi[r]);
        }
    }
}

template <int R, int MINB, int LLOG2>
cudaError_t launch_mirror_packed(const QwParams &prm, cudaStream_t st, qw_plan *plan)
{
    const int64_t m = prm.n / 2 + 1;                 // sites of the half ring
    const int L = (int)((m + R - 1) / R);
    const int nfull = (int)(m - (int64_t)L * (R - 1));
    constexpr int G = 32 >> LLOG2;
    auto kernel = rk4_cycle_mirror_packed<R, MINB, LLOG2>;
    if (plan) {
        cudaFuncAttributes fa;
        cudaError_t e = cudaFuncGetAttributes(&fa, kernel);
        if (e != cudaSuccess) return e;
        int nb = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, 32, 0);
        if (e != cudaSuccess) return e;
        plan->path = 9;
        plan->sites_per_thread = R;
        plan->threads_per_block = 32;
        plan->blocks_per_sm = nb;
        plan->regs_per_thread = fa.numRegs;
        plan->smem_bytes = (int)fa.sharedSizeBytes;
        plan->steps_per_pass = G;               // points per warp
        return cudaSuccess;
    }
    const int64_t blocks = packed_warps(prm, G);
    kernel<<<(unsigned)blocks, 32, 0, st>>>(prm, L, nfull);
    return cudaGetLastError();
}

template <int TPP_LOG2>
cudaError_t launch_packed_horner_t(const QwParams &prm, cudaStream_t st, qw_plan *plan)
{
    constexpr int MINB = 12;
    constexpr int groups = 32 >> TPP_LOG2;
    auto kernel = rk4_cycle_packed_horner<MINB, TPP_LOG2>;
    if (plan) {
        cudaFuncAttributes fa;
        cudaError_t e = cudaFuncGetAttributes(&fa, kernel);
        if (e != cudaSuccess) return e;
        int nb = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, 32, 0);
        if (e != cudaSuccess) return e;
        plan->path = 7;
        plan->sites_per_thread = 16;
        plan->threads_per_block = 32;
        plan->blocks_per_sm = nb;
        plan->regs_per_thread = fa.numRegs;
        plan->smem_bytes = (int)fa.sharedSizeBytes;
        plan->steps_per_pass = groups;          // points per warp
        return cudaSuccess;
    }
    const int64_t blocks = packed_warps(prm, groups);
    kernel<<<(unsigned)blocks, 32, 0, st>>>(prm);
    return cudaGetLastError();
}

cudaError_t launch_packed_horner(const QwParams &prm, int tpp_log2, cudaStream_t st, qw_plan *plan)
{
    if (tpp_log2 == 3) return launch_packed_horner_t<3>(prm, st, plan);
    if (tpp_log2 == 4) return launch_packed_horner_t<4>(prm, st, plan);
    return launch_packed_horner_t<5>(prm, st, plan);
}

template <int R, int MINB>
cudaError_t launch_packed(const QwParams &prm, int tpp_log2, cudaStream_t st, qw_plan *plan)
{
    const int groups = 32 >> tpp_log2;
    if (plan) {
        cudaFuncAttributes fa;
        cudaError_t e = cudaFuncGetAttributes(&fa, rk4_cycle_packed<R, MINB, false>);
        if (e != cudaSuccess) return e;
        int nb = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, rk4_cycle_packed<R, MINB, false>, 32, 0);
        if (e != cudaSuccess) return e;
        plan->path = 4;
        plan->sites_per_thread = R;
        plan->threads_per_block = 32;
        plan->blocks_per_sm = nb;
        plan->regs_per_thread = fa.numRegs;
        plan->smem_bytes = (int)fa.sharedSizeBytes;
        plan->steps_per_pass = groups;          // points per warp
        return cudaSuccess;
    }
    const int64_t blocks = packed_warps(prm, groups);
    if (prm.step_size > 0.0)
        rk4_cycle_packed<R, MINB, true><<<(unsigned)blocks, 32, 0, st>>>(prm, tpp_log2);
    else
        rk4_cycle_packed<R, MINB, false><<<(unsigned)blocks, 32, 0, st>>>(prm, tpp_log2);
    return cudaGetLastError();
}

}  // namespace

static bool packed_fits(int64_t n, int R, int max_tpp, int *tpp_log2)
{
    if (n % R) return false;
    const int64_t tpp = n / R;
    if (tpp < 1 || tpp > max_tpp || (tpp & (tpp - 1))) return false;
    int l = 0;
    while ((1 << l) < tpp) ++l;
    *tpp_log2 = l;
    return true;
}

// (R, log2 TPP) for the packed kernel, or R = 0 when n does not factor as R * 2^k.
//  * enough points to fill the machine: throughput order R = 8, 16, then smaller, with at
//    least two points per warp.  Measured: N = 64, R = 8 vs 16: 1.83 vs 1.90 ms (C2);
//    N = 256, R = 16 packed vs R = 8 block-per-point: 5.03e11 vs 4.71e11 site-updates/s (C4).
//  * a handful of points (C1: one): the smallest R, i.e. as many lanes per point as possible,
//    which minimises the latency of a stage.
// the batch size the geometry is chosen for
static int64_t packed_batch(const QwParams &prm)
{
    return (prm.flags & QW_FLAG_FIXED_GEOMETRY) ? (int64_t)1 << 40 : prm.n_points;
}

int qw_packed_geometry(int64_t n, int64_t n_points, int max_r, int *tpp_log2)
{
    const int sms = qw_sm_count();
    const int cand[9] = {8, 16, 7, 6, 5, 4, 3, 2, 1};
    for (int i = 0; i < 9; ++i) {
        const int R = cand[i];
        int l;
        if (R > max_r || !packed_fits(n, R, 16, &l)) continue;
        const int64_t per_warp = 32 >> l;
        if ((n_points + per_warp - 1) / per_warp >= 2 * (int64_t)sms) { *tpp_log2 = l; return R; }
    }
    const int small[9] = {1, 2, 3, 4, 5, 6, 7, 8, 16};
    for (int i = 0; i < 9; ++i)
        if (small[i] <= max_r && packed_fits(n, small[i], 32, tpp_log2)) return small[i];
    return 0;
}

// sites per lane of the mirror-reduced kernel (one point per warp), or 0: the half ring
// n/2 + 1 = L (R - 1) + nfull must fit 32 lanes (and L >= R, so that 1 <= nfull <= L).  The
// fewest sites per lane that fit fill the most lanes: N = 600 is 28 lanes of 11 sites (86 % of
// the slots) instead of 18 lanes of 17 (55 %).  QW_FLAG_NO_PACK: 9 or 17 only (A/B).
int qw_mirror_sites_per_lane(const QwParams &prm)
{
    if (!(prm.flags & QW_FLAG_MIRROR) || prm.topology != QW_TOPO_CYCLE) return 0;
    const int64_t m = prm.n / 2 + 1;
    if (m < 81) return 0;
    const int cand[5] = {9, 11, 13, 15, 17};
    for (int i = 0; i < 5; ++i) {
        const int R = cand[i];
        if ((prm.flags & QW_FLAG_NO_PACK) && R != 9 && R != 17) continue;
        if ((m + R - 1) / R <= 32) return R;
    }
    return 0;
}

// Half rings of at most 16 lanes: 2 or 4 points per warp (rk4_cycle_mirror_packed).  Returns the
// sites per lane (17 or 9) and the lanes per point (8 or 16) of the geometry that fills the most
// lanes, or 0: one point per warp fills nearly as many, the step rule gives the points of a warp
// no common step count (qw_packed_supported), the batch does not fill the machine twice (a few
// points: one warp each is the faster way; not under QW_FLAG_FIXED_GEOMETRY), QW_FLAG_NO_PACK.
// A point of L lanes holds L (R - 1) + 1 .. L R sites.
static int mirror_packed_geometry(const QwParams &prm, int *llog2)
{
    const int r1 = qw_mirror_sites_per_lane(prm);
    if (r1 == 0 || (prm.flags & (QW_FLAG_NO_PACK | QW_FLAG_OCC4))) return 0;
    const bool uniform = prm.step_size == 0.0 && prm.n_steps == nullptr;
    const bool columns = prm.step_size > 0.0 && prm.time_array != nullptr &&
                         prm.order == nullptr && prm.n_time > 0;
    if (!(uniform || columns)) return 0;
    const int64_t m = prm.n / 2 + 1;
    // score = the lanes' slots that hold a site (x 32), weighted: 17 sites per lane run about
    // 10 % faster per filled slot than 9 (tools/ab_large_n.py, N = 160 .. 286; 13 and 15 in
    // between), and one point per warp (32 steps per table refill, 9 sites per lane in this
    // range) gets the same bonus.  13 and 15 sites per lane fill the gaps between the half rings
    // that fit 17-site lanes (16 L + 1 .. 17 L sites).
    const int cand[7][2] = {{17, 3}, {15, 3}, {13, 3}, {17, 4}, {15, 4}, {13, 4}, {9, 4}};
    int best = 0;
    int64_t best_score = m * 110 / r1;
    for (int i = 0; i < 7; ++i) {
        const int R = cand[i][0], lp = 1 << cand[i][1], g = 32 / lp;
        if (R != 9 && (prm.flags & QW_FLAG_R8)) continue;          // A/B
        const int64_t L = (m + R - 1) / R;
        if (L > lp || m < L * (R - 1) + 1) continue;
        const int64_t score = g * m * (100 + (R - 9) * 10 / 8) / R;
        if (score > best_score) { best_score = score; best = R; *llog2 = cand[i][1]; }
    }
    if (best == 0) return 0;
    const int64_t g = 32 >> *llog2;
    if ((packed_batch(prm) + g - 1) / g < 2 * (int64_t)qw_sm_count()) return 0;
    return best;
}

cudaError_t qw_launch_mirror(const QwParams &prm, cudaStream_t st, qw_plan *plan)
{
    int llog2 = 0;
    switch (mirror_packed_geometry(prm, &llog2)) {
    case 17:
        if (llog2 == 3) return launch_mirror_packed<17, 12, 3>(prm, st, plan);
        return launch_mirror_packed<17, 12, 4>(prm, st, plan);
    case 15:
        if (llog2 == 3) return launch_mirror_packed<15, 12, 3>(prm, st, plan);
        return launch_mirror_packed<15, 12, 4>(prm, st, plan);
    case 13:
        if (llog2 == 3) return launch_mirror_packed<13, 12, 3>(prm, st, plan);
        return launch_mirror_packed<13, 12, 4>(prm, st, plan);
    case 9: return launch_mirror_packed<9, 12, 4>(prm, st, plan);
    default: break;
    }
    switch (qw_mirror_sites_per_lane(prm)) {
    case 17:
        if (prm.flags & QW_FLAG_OCC4) return launch_mirror<17, 8>(prm, st, plan);   // A/B: 196 registers
        return launch_mirror<17, 12>(prm, st, plan);
    case 15: return launch_mirror<15, 12>(prm, st, plan);
    case 13: return launch_mirror<13, 12>(prm, st, plan);
    case 11: return launch_mirror<11, 12>(prm, st, plan);
    case 9: return launch_mirror<9, 12>(prm, st, plan);
    default: return cudaErrorInvalidConfiguration;
    }
}

bool qw_packed_supported(const QwParams &prm)
{
    int l;
    // one common step count per warp: the n_steps rule, or a fixed step size on a heatmap
    // (grid mode), where a warp takes points of one T column
    const bool uniform = prm.step_size == 0.0 && prm.n_steps == nullptr;
    const bool columns = prm.step_size > 0.0 && prm.time_array != nullptr &&
                         prm.order == nullptr && prm.n_time > 0;
    if (!(uniform || columns)) return false;
    if (prm.topology == QW_TOPO_COMPLETE)      // small complete graphs: any-dimension layout only
        return prm.n < 256 && packed_any_r(prm.n) > 0 && !(prm.flags & QW_FLAG_EXACT_SUM);
    if (!(prm.topology == QW_TOPO_CYCLE && prm.n >= 3)) return false;
    return qw_packed_geometry(prm.n, packed_batch(prm), (prm.flags & QW_FLAG_R8) ? 8 : 16, &l) > 0 ||
           packed_any_r(prm.n) > 0;
}

// nested form: N = 128 or 256 (16 sites per lane, 8 or 16 lanes per point), a batch that fills
// the machine, the n_steps rule (qw_packed_supported)
static bool packed_horner(const QwParams &prm, int *tpp_log2)
{
    if (prm.flags & (QW_FLAG_NO_HORNER | QW_FLAG_R8)) return false;
    if (prm.n != 128 && prm.n != 256 && prm.n != 512) return false;
    *tpp_log2 = prm.n == 128 ? 3 : (prm.n == 256 ? 4 : 5);
    // enough warps to fill the machine twice, as for the throughput geometries above;
    // smaller batches keep the latency-oriented stage-form geometries
    const int64_t per_warp = 32 >> *tpp_log2;
    const int sms = qw_sm_count();
    return (packed_batch(prm) + per_warp - 1) / per_warp >= 2 * (int64_t)sms;
}

// half ring under QW_FLAG_MIRROR in the packed layout: rings of 6 .. 159 sites (from 160 on the
// one-warp-per-point kernel rk4_cycle_mirror applies and is selected before this path)
static bool packed_mirror(const QwParams &prm)
{
    return (prm.flags & QW_FLAG_MIRROR) && prm.topology == QW_TOPO_CYCLE && prm.n >= 6 &&
           prm.n < 160 && packed_any_r(prm.n / 2 + 1) > 0;
}

// Which packed kernel a problem runs, as a small descriptor: also the key by which
// qw_launch_packed_multi groups problems into one grid.
struct PackedClass {
    int kind;        // 0 none, 1 R 2^k ring, 2 any ring, 3 any ring / half ring (mirror),
                     // 4 small complete graph, 5 nested form
    int R;
    int a, b;        // kind 1, 5: log2(lanes per point), - ; kind 2, 3, 4: lanes per point, lanes with R sites
    int groups;      // points per warp
    bool column;     // fixed step size on a heatmap (a warp takes rows of one T column)
};

static PackedClass packed_class(const QwParams &prm)
{
    PackedClass c = {0, 0, 0, 0, 0, prm.step_size > 0.0};
    int l = 0;
    if (prm.topology == QW_TOPO_COMPLETE) {
        c.R = packed_any_r(prm.n);
        if (c.R == 0) return c;
        c.kind = 4;
        c.a = (int)((prm.n + c.R - 1) / c.R);
        c.b = (int)(prm.n - (int64_t)c.a * (c.R - 1));
        c.groups = 32 / c.a;
        return c;
    }
    if (packed_mirror(prm)) {
        const int64_t m = prm.n / 2 + 1;
        // (8 sites per lane wherever the deal works out, e.g. 36 sites on five lanes, was tried:
        // more points per warp, but 423 instead of 299 ns per step and the sweep of
        // BCB_latest.py:309-320 ends with its longest columns: 44.2 ms instead of 40.6)
        c.kind = 3;
        c.R = packed_any_r(m);
        c.a = (int)((m + c.R - 1) / c.R);
        c.b = (int)(m - (int64_t)c.a * (c.R - 1));
        c.groups = 32 / c.a;
        return c;
    }
    if (packed_horner(prm, &l)) {
        c.kind = 5; c.R = 16; c.a = l; c.groups = 32 >> l;
        return c;
    }
    const int R = qw_packed_geometry(prm.n, packed_batch(prm), (prm.flags & QW_FLAG_R8) ? 8 : 16, &l);
    if (R > 0) {
        c.kind = 1; c.R = R; c.a = l; c.groups = 32 >> l;
        return c;
    }
    c.R = packed_any_r(prm.n);
    if (c.R == 0) return c;
    c.kind = 2;
    c.a = (int)((prm.n + c.R - 1) / c.R);
    c.b = (int)(prm.n - (int64_t)c.a * (c.R - 1));
    c.groups = 32 / c.a;
    return c;
}

cudaError_t qw_launch_packed(const QwParams &prm, cudaStream_t st, qw_plan *plan)
{
    const PackedClass c = packed_class(prm);
    const int l = c.a;
    switch (c.kind) {
    case 4:
        switch (c.R) {
        case 8: return launch_packed_complete<8, 12>(prm, st, plan);
        case 4: return launch_packed_complete<4, 16>(prm, st, plan);
        case 2: return launch_packed_complete<2, 16>(prm, st, plan);
        default: return cudaErrorInvalidConfiguration;
        }
    case 3:
        switch (c.R) {
        case 8: return launch_packed_any<8, 12, true>(prm, st, plan);
        case 4: return launch_packed_any<4, 16, true>(prm, st, plan);
        case 2: return launch_packed_any<2, 16, true>(prm, st, plan);
        default: return cudaErrorInvalidConfiguration;
        }
    case 5: return launch_packed_horner(prm, l, st, plan);
    case 2:                          // not R * 2^k: R or R-1 sites per lane
        switch (c.R) {
        case 8: return launch_packed_any<8, 12, false>(prm, st, plan);
        case 4: return launch_packed_any<4, 16, false>(prm, st, plan);
        case 2: return launch_packed_any<2, 16, false>(prm, st, plan);
        default: return cudaErrorInvalidConfiguration;
        }
    case 1: break;
    default: return cudaErrorInvalidConfiguration;
    }
    switch (c.R) {
    case 16: return launch_packed<16, 8>(prm, l, st, plan);
    case 8:
        if (prm.flags & QW_FLAG_OCC4) return launch_packed<8, 16>(prm, l, st, plan);   // A/B
        return launch_packed<8, 12>(prm, l, st, plan);
    case 7: return launch_packed<7, 12>(prm, l, st, plan);
    case 6: return launch_packed<6, 12>(prm, l, st, plan);
    case 5: return launch_packed<5, 16>(prm, l, st, plan);
    case 4: return launch_packed<4, 16>(prm, l, st, plan);
    case 3: return launch_packed<3, 16>(prm, l, st, plan);
    case 2: return launch_packed<2, 16>(prm, l, st, plan);
    case 1: return launch_packed<1, 16>(prm, l, st, plan);
    default: return cudaErrorInvalidConfiguration;
    }
}

// ---------------------------------------------------------------------------------------------
// Many problems, one grid.  Problems whose packed kernel coincides (same kind, sites per lane
// and step rule) share a launch: their warps are laid end to end, longest problems first, and
// every warp finds its problem in a job table (PackedJob).  Each problem keeps the geometry its
// own launch would use, so the results are bitwise those of separate launches.  Returns, per
// problem, whether it was taken (false: the caller launches it on its own).
// Multi-capable: kind 1 with 8 sites per lane, kinds 2 and 3 (the classes of BASELINE config 2
// and of the reference's production sweep, BCB_latest.py:309-320).
static bool packed_multi_capable(const PackedClass &c)
{
    return c.kind == 1 || c.kind == 2 || c.kind == 3;
}

static long long packed_multi_key(const PackedClass &c)
{
    return (long long)c.kind * 1000 + c.R * 10 + (c.column ? 1 : 0);
}

// The library's own streams for concurrent class grids (per device, created on first use).
struct MultiStreams {
    static constexpr int N = 4;
    std::mutex mutex;
    cudaStream_t aux[N] = {};
    cudaEvent_t fork = nullptr, join[N] = {};
    bool ok = false;
};

static MultiStreams *multi_streams()
{
    static MultiStreams table[64];
    static std::mutex create;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    MultiStreams &m = table[dev];
    std::lock_guard<std::mutex> lock(create);
    if (!m.ok) {
        bool good = cudaEventCreateWithFlags(&m.fork, cudaEventDisableTiming) == cudaSuccess;
        for (int i = 0; i < MultiStreams::N && good; ++i)
            good = cudaStreamCreateWithFlags(&m.aux[i], cudaStreamNonBlocking) == cudaSuccess &&
                   cudaEventCreateWithFlags(&m.join[i], cudaEventDisableTiming) == cudaSuccess;
        if (!good) { cudaGetLastError(); return nullptr; }
        m.ok = true;
    }
    return &m;
}

cudaError_t qw_launch_packed_multi(const QwParams *prms, int n, unsigned char *taken,
                                   cudaStream_t st, int64_t *launches)
{
    struct Item { int idx; PackedClass c; long long key; double work; };
    std::vector<Item> items;
    for (int i = 0; i < n; ++i) {
        taken[i] = 0;
        if (prms[i].n_points <= 0 || !qw_packed_supported(prms[i])) continue;
        if (prms[i].flags & QW_FLAG_OCC4) continue;
        const PackedClass c = packed_class(prms[i]);
        if (!packed_multi_capable(c)) continue;
        // a warp's work ~ sites per lane x its step count; a rough figure for the ordering only
        double steps = (double)prms[i].n_steps_uniform;
        if (c.column) steps = 1.0;            // refined per job below is not worth a device read
        items.push_back({i, c, packed_multi_key(c), (double)prms[i].n * steps});
    }
    std::stable_sort(items.begin(), items.end(),
                     [](const Item &x, const Item &y) { return x.key < y.key; });
    // one grid per kernel class; the grids run CONCURRENTLY (fork / join on the library's own
    // streams): a class of short problems is a single latency-bound wave (the production sweep
    // of BCB_latest.py:309-320: rings of 7 .. 29 sites, 6.9 ms at 32 % of the FP64 pipe) that
    // hides behind the big class instead of queueing in front of it.  Small classes are issued
    // first, the largest one last and on the caller's stream.
    struct Group { size_t b, e; double work; };
    std::vector<Group> groups;
    for (size_t b = 0; b < items.size();) {
        size_t e = b;
        while (e < items.size() && items[e].key == items[b].key) ++e;
        if (e - b >= 2) {
            double work = 0.0;
            for (size_t k = b; k < e; ++k)
                work += (double)prms[items[k].idx].n * (double)prms[items[k].idx].n_points;
            groups.push_back({b, e, work});
        }
        b = e;
    }
    std::stable_sort(groups.begin(), groups.end(),
                     [](const Group &x, const Group &y) { return x.work < y.work; });
    MultiStreams *ms = groups.size() > 1 ? multi_streams() : nullptr;
    std::unique_lock<std::mutex> lock;
    if (ms) {
        lock = std::unique_lock<std::mutex>(ms->mutex);
        if (cudaEventRecord(ms->fork, st) != cudaSuccess) { cudaGetLastError(); ms = nullptr; }
    }
    cudaError_t err = cudaSuccess;
    std::vector<int> used;
    for (size_t gi = 0; gi < groups.size() && err == cudaSuccess; ++gi) {
        const size_t b = groups[gi].b, e = groups[gi].e;
        cudaStream_t gst = st;
        if (ms && gi + 1 < groups.size()) {            // not the largest: an auxiliary stream
            const int a = (int)(gi % MultiStreams::N);
            gst = ms->aux[a];
            err = cudaStreamWaitEvent(gst, ms->fork, 0);
            if (err != cudaSuccess) break;
            if (std::find(used.begin(), used.end(), a) == used.end()) used.push_back(a);
        }
        {
            cudaStream_t st = gst;                     // the group's stream from here on
            std::vector<Item> grp(items.begin() + b, items.begin() + e);
            std::stable_sort(grp.begin(), grp.end(),
                             [](const Item &x, const Item &y) { return x.work > y.work; });
            // Fixed step size on heatmaps: a warp's length is its column's T, so the grid walks
            // the problems column rank by column rank -- every problem's longest column first,
            // then every second-longest ... -- which is close to longest-first over the whole grid
            // (the production sweep: makespan 27.3 k steps instead of 30.9 k, ideal 24.7 k).
            std::vector<PackedJob> jobs;
            int64_t w = 0;
            int64_t max_cols = 1;
            if (grp[0].c.column)
                for (size_t k = 0; k < grp.size(); ++k)
                    max_cols = std::max<int64_t>(max_cols, prms[grp[k].idx].n_time);
            for (int64_t rank = 0; rank < max_cols; ++rank) {
                for (size_t k = 0; k < grp.size(); ++k) {
                    const QwParams &pk = prms[grp[k].idx];
                    const int64_t total = packed_warps(pk, grp[k].c.groups);
                    int64_t first = 0, count = total;
                    if (grp[k].c.column) {
                        if (rank >= pk.n_time) continue;
                        count = total / pk.n_time;         // warps per column
                        first = rank * count;
                    }
                    PackedJob j;
                    memset(&j, 0, sizeof(j));
                    j.prm = pk;
                    j.a = grp[k].c.a;
                    j.b = grp[k].c.b;
                    j.warp_begin = w;
                    j.warp_offset = first;
                    w += count;
                    j.warp_end = w;
                    jobs.push_back(j);
                }
            }
            if (w > 0x7fffffffLL) continue;                    // leave them to single launches
            PackedJob *dj = nullptr;
            err = qw_scratch_alloc((void **)&dj, jobs.size() * sizeof(PackedJob), st);
            if (err != cudaSuccess) break;
            err = cudaMemcpyAsync(dj, jobs.data(), jobs.size() * sizeof(PackedJob),
                                  cudaMemcpyHostToDevice, st);
            const PackedClass &c = grp[0].c;
            const int nj = (int)jobs.size();
            const unsigned grid = (unsigned)w;
            if (err == cudaSuccess) {
#define QW_PM(KERNEL) KERNEL<<<grid, 32, 0, st>>>(dj, nj)
                if (c.kind == 1) {
#define QW_PM1(RR, MB)                                                                      \
                    case RR:                                                                \
                        if (c.column) QW_PM((rk4_cycle_packed_multi<RR, MB, true>));        \
                        else QW_PM((rk4_cycle_packed_multi<RR, MB, false>));                \
                        break;
                    switch (c.R) {
                        QW_PM1(16, 8) QW_PM1(8, 12) QW_PM1(7, 12) QW_PM1(6, 12) QW_PM1(5, 16)
                        QW_PM1(4, 16) QW_PM1(3, 16) QW_PM1(2, 16) QW_PM1(1, 16)
                    default: break;
                    }
#undef QW_PM1
                } else if (c.kind == 2) {
                    if (c.R == 8) QW_PM((rk4_cycle_packed_any_multi<8, 12, false>));
                    else if (c.R == 4) QW_PM((rk4_cycle_packed_any_multi<4, 16, false>));
                    else QW_PM((rk4_cycle_packed_any_multi<2, 16, false>));
                } else {
                    if (c.R == 8) QW_PM((rk4_cycle_packed_any_multi<8, 12, true>));
                    else if (c.R == 4) QW_PM((rk4_cycle_packed_any_multi<4, QW_ANY4_MINB, true>));
                    else QW_PM((rk4_cycle_packed_any_multi<2, QW_ANY2_MINB, true>));
                }
#undef QW_PM
                err = cudaGetLastError();
                ++*launches;
            }
            qw_scratch_free(dj, st);
            if (err == cudaSuccess)
                for (size_t k = 0; k < grp.size(); ++k) taken[grp[k].idx] = 1;
        }
    }
    for (int a : used) {                               // join: the caller's stream waits for all
        cudaError_t e2 = cudaEventRecord(ms->join[a], ms->aux[a]);
        if (e2 == cudaSuccess) e2 = cudaStreamWaitEvent(st, ms->join[a], 0);
        if (e2 != cudaSuccess && err == cudaSuccess) err = e2;
    }
    return err;
}
