// Complete graph under QW_FLAG_MIRROR: the walk never leaves a two-dimensional subspace.
//
// The reference starts every walk from the uniform superposition (BCB.py:61-66,
// QW_Search.py:141-147) and the complete graph's Laplacian and the oracle projector are
// invariant under every permutation of the unmarked vertices, so psi_j(t) has one common value
// o(t) on all j != w.  With a = psi_w and m = N - 1,
//
//     (L psi)_w = m (a - o),      (L psi)_j = o - a   (j != w),
//
// and the reference's RK4 step (BCB.py:163-190 with the step rule of QW_Search.py:149-164)
// closes on the pair (a, o): one thread integrates one grid point, O(S) work instead of
// O(S N).  p = |a|^2 / norm, norm = |a|^2 + m |o|^2, psi (if requested) is expanded on the way
// out.  Same quantities as the full kernels to rounding (the tests bound |dp| by 1e-12).
// Opt-in like the cycle graph's mirror reduction: it assumes the uniform initial state, which
// is the only one the interface has.
#include "qw_common.cuh"

namespace {

constexpr int SYM_THREADS = 128;

struct Pair { double ar, ai, or_, oi; };

// k = -i (alpha L x + d x_w e_w) on the pair
__device__ __forceinline__ Pair sym_rhs(const Pair &x, const double2 ad, const double m)
{
    const double dr = x.ar - x.or_, di = x.ai - x.oi;
    const double hwr = fma(ad.x * m, dr, ad.y * x.ar), hwi = fma(ad.x * m, di, ad.y * x.ai);
    const double hor = -ad.x * dr, hoi = -ad.x * di;
    return Pair{hwi, -hwr, hoi, -hor};
}

__device__ __forceinline__ Pair sym_axpy(const Pair &y, const double c, const Pair &k)
{
    return Pair{fma(c, k.ar, y.ar), fma(c, k.ai, y.ai), fma(c, k.or_, y.or_), fma(c, k.oi, y.oi)};
}

__global__ void __launch_bounds__(SYM_THREADS) rk4_complete_symmetric(const QwParams prm)
{
    const int64_t b = (int64_t)blockIdx.x * SYM_THREADS + threadIdx.x;
    const bool active = b < prm.n_points;
    const double m = (double)(prm.n - 1);
    const double amp = 1.0 / sqrt((double)prm.n);
    Pair y{amp, 0.0, amp, 0.0};
    int64_t q = 0;
    if (active) {
        const QwPoint pt = qw_point(prm, b);
        q = pt.q;
        const double h = pt.h;
        for (int64_t s = 0; s < pt.steps; ++s) {
            const double t = (double)s * h;
            const double2 c1 = qw_operator_scalars(t / pt.T, pt.gamma, prm.schedule, prm.gamma_on);
            const double2 c2 = qw_operator_scalars((t + 0.5 * h) / pt.T, pt.gamma, prm.schedule,
                                                   prm.gamma_on);
            const double2 c4 = qw_operator_scalars((t + h) / pt.T, pt.gamma, prm.schedule,
                                                   prm.gamma_on);
            const Pair k1 = sym_rhs(y, c1, m);
            const Pair k2 = sym_rhs(sym_axpy(y, 0.5 * h, k1), c2, m);
            const Pair k3 = sym_rhs(sym_axpy(y, 0.5 * h, k2), c2, m);
            const Pair k4 = sym_rhs(sym_axpy(y, h, k3), c4, m);
            Pair acc = sym_axpy(y, h / 6.0, k1);
            acc = sym_axpy(acc, h / 3.0, k2);
            acc = sym_axpy(acc, h / 3.0, k3);
            y = sym_axpy(acc, h / 6.0, k4);
        }
        const double pw = fma(y.ar, y.ar, y.ai * y.ai);
        const double nrm = fma(m, fma(y.or_, y.or_, y.oi * y.oi), pw);
        prm.p[q] = pw / nrm;
        if (prm.norm) prm.norm[q] = nrm;
    }
    if (prm.psi) {                       // warp-cooperative expansion, coalesced
        const unsigned lane = threadIdx.x & 31u;
#pragma unroll 1
        for (int l = 0; l < 32; ++l) {
            if (!__shfl_sync(0xffffffffu, (int)active, l)) continue;
            const int64_t ql = __shfl_sync(0xffffffffu, q, l);
            const double2 a = make_double2(__shfl_sync(0xffffffffu, y.ar, l),
                                           __shfl_sync(0xffffffffu, y.ai, l));
            const double2 o = make_double2(__shfl_sync(0xffffffffu, y.or_, l),
                                           __shfl_sync(0xffffffffu, y.oi, l));
            double2 *dst = prm.psi + (size_t)ql * (size_t)prm.n;
            for (int64_t j = lane; j < prm.n; j += 32) dst[j] = (j == prm.target) ? a : o;
        }
    }
}

}  // namespace

bool qw_symmetric_supported(const QwParams &prm)
{
    return (prm.flags & QW_FLAG_MIRROR) && prm.topology == QW_TOPO_COMPLETE;
}

cudaError_t qw_launch_symmetric(const QwParams &prm, cudaStream_t st, qw_plan *plan)
{
    if (plan) {
        cudaFuncAttributes fa;
        cudaError_t e = cudaFuncGetAttributes(&fa, rk4_complete_symmetric);
        if (e != cudaSuccess) return e;
        int occ = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, rk4_complete_symmetric,
                                                          SYM_THREADS, 0);
        if (e != cudaSuccess) return e;
        plan->path = 10;
        plan->sites_per_thread = 2;
        plan->threads_per_block = SYM_THREADS;
        plan->blocks_per_sm = occ;
        plan->regs_per_thread = fa.numRegs;
        plan->smem_bytes = (int)fa.sharedSizeBytes;
        return cudaSuccess;
    }
    const int64_t blocks = (prm.n_points + SYM_THREADS - 1) / SYM_THREADS;
    rk4_complete_symmetric<<<(unsigned)blocks, SYM_THREADS, 0, st>>>(prm);
    return cudaGetLastError();
}
