// Global-memory fixed-step RK4: any N, both graphs.  One persistent thread block per
// point, state (y, two stage buffers, accumulator) in a per-block scratch slice that the
// L1/L2 keep close.  This is the catch-all below the register-resident kernels
// (qw_resident.cu) and the tile-fused streaming kernel (qw_stream.cu): odd N above 1024,
// complete graphs above 4096 vertices, and the QW_FLAG_FORCE_GENERIC test path.
// Same arithmetic as the reference path it replaces (BCB.py:28-99, 163-203 with the step
// rule of QW_Search.py:149-164); vertices keep their physical labels here.
#include "qw_common.cuh"

namespace {

constexpr int GEN_THREADS = 512;

struct GenTables {
    double2 ab[4 * QW_CHUNK];
    double2 d[4 * QW_CHUNK];
    double2 ad[3 * QW_CHUNK];
    double red[64];
};

// x -> (out = base + A k, acc (+)= B k) with k = -i (a L x + d x_w e_w) folded into the
// weights; returns this thread's partial sum of `out` (needed by the complete graph).
template <int STAGE>
__device__ __forceinline__ double2 gen_stage(const QwParams &prm, const double2 *__restrict__ x,
                                             double2 *__restrict__ y, double2 *__restrict__ out,
                                             double2 *__restrict__ acc, const double2 ab,
                                             const double2 dd, const double2 sum)
{
    const int64_t n = prm.n, w = prm.target;
    const bool cyc = prm.topology == QW_TOPO_CYCLE;
    const double nd = (double)n;
    double sr = 0.0, si = 0.0;
    for (int64_t j = threadIdx.x; j < n; j += blockDim.x) {
        const double2 c = x[j];
        double vr, vi;
        if (cyc) {
            const double2 l = x[j == 0 ? n - 1 : j - 1];
            const double2 r = x[j == n - 1 ? 0 : j + 1];
            vr = fma(2.0, c.x, -(l.x + r.x));
            vi = fma(2.0, c.y, -(l.y + r.y));
        } else {
            vr = fma(nd, c.x, -sum.x);
            vi = fma(nd, c.y, -sum.y);
        }
        const double2 yy = y[j];
        double2 a = (STAGE == 0) ? yy : acc[j];
        a.x = fma(ab.y, vi, a.x);
        a.y = fma(-ab.y, vr, a.y);
        double2 o = make_double2(fma(ab.x, vi, yy.x), fma(-ab.x, vr, yy.y));
        if (j == w) {
            a.x = fma(dd.y, c.y, a.x);
            a.y = fma(-dd.y, c.x, a.y);
            o.x = fma(dd.x, c.y, o.x);
            o.y = fma(-dd.x, c.x, o.y);
        }
        if (STAGE < 3) { acc[j] = a; out[j] = o; sr += o.x; si += o.y; }
        else           { out[j] = a; sr += a.x; si += a.y; }     // out == y buffer of step n+1
    }
    return make_double2(sr, si);
}

__global__ void __launch_bounds__(GEN_THREADS, 1) rk4_generic(const QwParams prm)
{
    __shared__ GenTables tb;
    const int64_t n = prm.n;
    const bool need_sum = prm.topology == QW_TOPO_COMPLETE;
    double2 *base = prm.scratch + (size_t)blockIdx.x * 5 * (size_t)n;
    double2 *ya = base, *yb = base + n, *ua = base + 2 * n, *ub = base + 3 * n,
            *acc = base + 4 * n;

    for (int64_t b = blockIdx.x; b < prm.n_points; b += gridDim.x) {
        const QwPoint pt = qw_point(prm, b);
        const double amp = 1.0 / sqrt((double)n);
        double2 *y = ya, *ynew = yb;
        double lr = 0.0;
        for (int64_t j = threadIdx.x; j < n; j += blockDim.x) {
            y[j] = make_double2(amp, 0.0);
            lr += amp;
        }
        double2 sy = make_double2(0.0, 0.0);
        if (need_sum) sy = qw_block_sum2(lr, 0.0, tb.red);
        __syncthreads();

        for (int64_t n0 = 0; n0 < pt.steps; n0 += QW_CHUNK) {
            qw_fill_table(tb.ab, tb.d, tb.ad, pt, n0, prm.schedule, prm.gamma_on);
            const int cnt = (int)((pt.steps - n0 < QW_CHUNK) ? (pt.steps - n0) : QW_CHUNK);
            for (int s = 0; s < cnt; ++s) {
                double2 part, sx = sy;
                part = gen_stage<0>(prm, y, y, ua, acc, tb.ab[4 * s], tb.d[4 * s], sx);
                if (need_sum) sx = qw_block_sum2(part.x, part.y, tb.red);
                __syncthreads();
                part = gen_stage<1>(prm, ua, y, ub, acc, tb.ab[4 * s + 1], tb.d[4 * s + 1], sx);
                if (need_sum) sx = qw_block_sum2(part.x, part.y, tb.red);
                __syncthreads();
                part = gen_stage<2>(prm, ub, y, ua, acc, tb.ab[4 * s + 2], tb.d[4 * s + 2], sx);
                if (need_sum) sx = qw_block_sum2(part.x, part.y, tb.red);
                __syncthreads();
                part = gen_stage<3>(prm, ua, y, ynew, acc, tb.ab[4 * s + 3], tb.d[4 * s + 3], sx);
                if (need_sum) sy = qw_block_sum2(part.x, part.y, tb.red);
                __syncthreads();
                double2 *tmp = y; y = ynew; ynew = tmp;
            }
        }
        double nrm = 0.0;
        for (int64_t j = threadIdx.x; j < n; j += blockDim.x) {
            const double2 v = y[j];
            nrm = fma(v.x, v.x, fma(v.y, v.y, nrm));
            if (prm.psi) prm.psi[(size_t)pt.q * (size_t)n + j] = v;
        }
        const double2 tot = qw_block_sum2(nrm, 0.0, tb.red);
        if (threadIdx.x == 0) {
            const double2 v = y[prm.target];
            prm.p[pt.q] = fma(v.x, v.x, v.y * v.y) / tot.x;
            if (prm.norm) prm.norm[pt.q] = tot.x;
        }
        __syncthreads();
    }
}

}  // namespace

// Blocks to launch and scratch bytes needed for the generic kernel on this device.
cudaError_t qw_generic_geometry(const QwParams &prm, int *blocks, size_t *scratch_bytes,
                                qw_plan *plan)
{
    int dev = 0, sms = 0, occ = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return e;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, rk4_generic, GEN_THREADS, 0);
    if (e != cudaSuccess) return e;
    if (occ < 1) occ = 1;
    int64_t nb = (int64_t)sms * occ;
    if (nb > prm.n_points) nb = prm.n_points;
    if (nb < 1) nb = 1;
    // five complex vectors per running block: no more blocks than three quarters of the free
    // memory hold (at least one; if even one does not fit, the caller reports it as unsupported)
    const size_t per_block = 5 * (size_t)prm.n * sizeof(double2);
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) {
        const int64_t fit = (int64_t)((free_b / 4 * 3) / per_block);
        if (fit < 1) return cudaErrorMemoryAllocation;
        if (nb > fit) nb = fit;
    }
    *blocks = (int)nb;
    *scratch_bytes = (size_t)nb * per_block;
    if (plan) {
        cudaFuncAttributes fa;
        e = cudaFuncGetAttributes(&fa, rk4_generic);
        if (e != cudaSuccess) return e;
        plan->path = 2;
        plan->sites_per_thread = (int)((prm.n + GEN_THREADS - 1) / GEN_THREADS);
        plan->threads_per_block = GEN_THREADS;
        plan->blocks_per_sm = occ;
        plan->regs_per_thread = fa.numRegs;
        plan->smem_bytes = (int)fa.sharedSizeBytes;
    }
    return cudaSuccess;
}

cudaError_t qw_launch_generic(const QwParams &prm, int blocks, cudaStream_t st)
{
    rk4_generic<<<blocks, GEN_THREADS, 0, st>>>(prm);
    return cudaGetLastError();
}
