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
#include "qw_stage.cuh"

namespace {

struct PackedShared {
    double2 fac[3 * QW_CHUNK];       // (1 - g, -g) at tau_n, tau_n + 1/(2S), tau_n + 1/S
};

template <int R, int MINB>
__global__ void __launch_bounds__(32, MINB) rk4_cycle_packed(const QwParams prm, const int tpp_log2)
{
    __shared__ PackedShared sh;
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x;
    const int tpp = 1 << tpp_log2, groups = 32 >> tpp_log2;
    const int li = lane & (tpp - 1), grp = lane >> tpp_log2;
    int64_t b = (int64_t)blockIdx.x * groups + grp;
    const bool valid = b < prm.n_points;
    if (!valid) b = prm.n_points - 1;            // idle groups shadow the last point
    const QwPoint pt = qw_point(prm, b);
    const int64_t S = prm.n_steps_uniform;
    const bool own = (li == 0);
    const int left = (li == 0) ? lane + tpp - 1 : lane - 1;
    const int right = (li == tpp - 1) ? lane - tpp + 1 : lane + 1;

    // per-lane stage weights: A = wa * a(t), B = wb * a(t), DA = da * d(t)/dfac, ...
    const double h = pt.h;                       // 0 when T == 0: the state never moves
    const double gl = prm.gamma_on == QW_GAMMA_ON_LAPLACIAN ? pt.gamma : 1.0;
    const double gd = prm.gamma_on == QW_GAMMA_ON_ORACLE ? pt.gamma : 1.0;
    const double a_half = 0.5 * h * gl, a_full = h * gl, a_6 = h / 6.0 * gl, a_3 = h / 3.0 * gl;
    const double d_half = 0.5 * h * gd, d_full = h * gd, d_6 = h / 6.0 * gd, d_3 = h / 3.0 * gd;
    const bool is_static = prm.schedule == QW_SCHED_STATIC;

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
            const double tau = ((double)(n0 + st) + 0.5 * (double)k) * inv_s;
            const double g = is_static ? 1.0 : qw_schedule(tau, prm.schedule);
            sh.fac[i] = is_static ? make_double2(1.0, -1.0) : make_double2(1.0 - g, -g);
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

template <int R, int MINB>
cudaError_t launch_packed(const QwParams &prm, int tpp_log2, cudaStream_t st, qw_plan *plan)
{
    const int groups = 32 >> tpp_log2;
    if (plan) {
        cudaFuncAttributes fa;
        cudaError_t e = cudaFuncGetAttributes(&fa, rk4_cycle_packed<R, MINB>);
        if (e != cudaSuccess) return e;
        int nb = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, rk4_cycle_packed<R, MINB>, 32, 0);
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
    const int64_t blocks = (prm.n_points + groups - 1) / groups;
    rk4_cycle_packed<R, MINB><<<(unsigned)blocks, 32, 0, st>>>(prm, tpp_log2);
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
int qw_packed_geometry(int64_t n, int64_t n_points, int max_r, int *tpp_log2)
{
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess ||
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
            sms <= 0)
            sms = 148;
    }
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

bool qw_packed_supported(const QwParams &prm)
{
    int l;
    return prm.topology == QW_TOPO_CYCLE && prm.step_size == 0.0 && prm.n_steps == nullptr &&
           prm.n >= 3 &&
           qw_packed_geometry(prm.n, prm.n_points, (prm.flags & QW_FLAG_R8) ? 8 : 16, &l) > 0;
}

cudaError_t qw_launch_packed(const QwParams &prm, cudaStream_t st, qw_plan *plan)
{
    int l = 0;
    const int R = qw_packed_geometry(prm.n, prm.n_points, (prm.flags & QW_FLAG_R8) ? 8 : 16, &l);
    switch (R) {
    case 16: return launch_packed<16, 8>(prm, l, st, plan);
    case 8: return launch_packed<8, 12>(prm, l, st, plan);
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
