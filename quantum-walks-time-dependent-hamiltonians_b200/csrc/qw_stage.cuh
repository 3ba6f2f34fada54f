// One RK4 stage on the R sites a thread keeps in registers; shared by the resident kernels
// (qw_resident.cu) and the tile-fused streaming kernel (qw_stream.cu).
#pragma once

#include "qw_common.cuh"

template <int R>
struct State {
    double yr[R], yi[R];   // psi_n
    double ur[R], ui[R];   // stage input
    double ar[R], ai[R];   // RK4 accumulator
};

// One RK4 stage on a thread's R sites.  lap(x) is supplied by the caller through the
// neighbour values (cycle) or the global sum (complete):
//   v      = lap(x)                                   (unscaled Laplacian)
//   k      = -i v  ->  k.re = v.im, k.im = -v.re
//   STAGE 0: x = y;  u = y + A k;  acc = y + B k
//   STAGE 1,2: x = u;  u = y + A k;  acc += B k
//   STAGE 3: x = u;  y = acc + B k
// `own`: this thread holds the marked vertex in slot 0; D = (DA, DB) weights of d(t).
template <int R, int STAGE, bool CYCLE>
__device__ __forceinline__ void rk_stage(State<R> &s, const double A, const double B,
                                         const double2 hl, const double2 hr,
                                         const double2 sum, const double nd,
                                         const bool own, const double2 D)
{
    double pr = hl.x, pi = hl.y;
    const double x0r = (STAGE == 0) ? s.yr[0] : s.ur[0];
    const double x0i = (STAGE == 0) ? s.yi[0] : s.ui[0];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const double cr = (STAGE == 0) ? s.yr[r] : s.ur[r];
        const double ci = (STAGE == 0) ? s.yi[r] : s.ui[r];
        double vr, vi;
        if (CYCLE) {
            const double nr = (r + 1 < R) ? ((STAGE == 0) ? s.yr[r + 1] : s.ur[r + 1]) : hr.x;
            const double ni = (r + 1 < R) ? ((STAGE == 0) ? s.yi[r + 1] : s.ui[r + 1]) : hr.y;
            vr = fma(2.0, cr, -(pr + nr));
            vi = fma(2.0, ci, -(pi + ni));
            pr = cr;
            pi = ci;
        } else {
            vr = fma(nd, cr, -sum.x);
            vi = fma(nd, ci, -sum.y);
        }
        if (STAGE == 0) {
            s.ar[r] = fma(B, vi, s.yr[r]);
            s.ai[r] = fma(-B, vr, s.yi[r]);
        } else if (STAGE < 3) {
            s.ar[r] = fma(B, vi, s.ar[r]);
            s.ai[r] = fma(-B, vr, s.ai[r]);
        }
        if (STAGE < 3) {
            s.ur[r] = fma(A, vi, s.yr[r]);
            s.ui[r] = fma(-A, vr, s.yi[r]);
        } else {
            s.yr[r] = fma(B, vi, s.ar[r]);
            s.yi[r] = fma(-B, vr, s.ai[r]);
        }
    }
    if (own) {   // oracle term d(t) x_w on the marked vertex
        if (STAGE < 3) {
            s.ur[0] = fma(D.x, x0i, s.ur[0]);
            s.ui[0] = fma(-D.x, x0r, s.ui[0]);
            s.ar[0] = fma(D.y, x0i, s.ar[0]);
            s.ai[0] = fma(-D.y, x0r, s.ai[0]);
        } else {
            s.yr[0] = fma(D.y, x0i, s.yr[0]);
            s.yi[0] = fma(-D.y, x0r, s.yi[0]);
        }
    }
}
