// Cycle-graph resident kernel, second generation: edge sites first, early publication.
//
// Per stage a thread (1) waits until every neighbour has published the edge values of the
// current stage input, (2) updates its two edge sites, (3) publishes their new values for
// the next stage and arrives on the mbarrier, and only then (4) updates its interior
// sites.  The interior work (6 of 8 sites) therefore separates a warp's arrival from its
// next wait, so warps of a block no longer run in lock step as they do with one
// __syncthreads() per stage.  The oracle term lives in a copy of the stage code that only
// warp 0 executes (template OWN), instead of predicated-off FP64 instructions in all warps.
#pragma once

#include "qw_stage.cuh"

__device__ __forceinline__ uint32_t qw_smem_addr(const void *p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void qw_mbar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(qw_smem_addr(bar)), "r"(count));
}

__device__ __forceinline__ void qw_mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(qw_smem_addr(bar)) : "memory");
}

__device__ __forceinline__ void qw_mbar_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t done;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n" : "=r"(done) : "r"(qw_smem_addr(bar)), "r"(parity) : "memory");
    } while (!done);
}

// update of one site given the (old) stage input at the site and its two neighbours
template <int STAGE>
__device__ __forceinline__ void qw_site(const double lr, const double li, const double cr,
                                        const double ci, const double rr, const double ri,
                                        const double A, const double B, const double y_r,
                                        const double y_i, double &u_r, double &u_i, double &a_r,
                                        double &a_i, double &yo_r, double &yo_i)
{
    const double vr = fma(2.0, cr, -(lr + rr));
    const double vi = fma(2.0, ci, -(li + ri));
    if (STAGE == 0) {
        a_r = fma(B, vi, y_r);
        a_i = fma(-B, vr, y_i);
    } else if (STAGE < 3) {
        a_r = fma(B, vi, a_r);
        a_i = fma(-B, vr, a_i);
    }
    if (STAGE < 3) {
        u_r = fma(A, vi, y_r);
        u_i = fma(-A, vr, y_i);
    } else {
        yo_r = fma(B, vi, a_r);
        yo_i = fma(-B, vr, a_i);
    }
}

#define QW_X_R(s, r) ((STAGE == 0) ? (s).yr[r] : (s).ur[r])
#define QW_X_I(s, r) ((STAGE == 0) ? (s).yi[r] : (s).ui[r])

// Edge sites (slot 0 and slot R-1).  o0 / oL return the old stage-input values of the two
// edge slots, which the interior pass still needs as neighbours.
template <int R, int STAGE, bool OWN>
__device__ __forceinline__ void qw_edges(State<R> &s, const double A, const double B,
                                         const double2 hl, const double2 hr, const double2 D,
                                         double2 &o0, double2 &oL)
{
    o0 = make_double2(QW_X_R(s, 0), QW_X_I(s, 0));
    oL = make_double2(QW_X_R(s, R - 1), QW_X_I(s, R - 1));
    const double r0r = (R > 1) ? QW_X_R(s, (R > 1 ? 1 : 0)) : hr.x;
    const double r0i = (R > 1) ? QW_X_I(s, (R > 1 ? 1 : 0)) : hr.y;
    const double lLr = (R > 2) ? QW_X_R(s, (R > 2 ? R - 2 : 0)) : o0.x;
    const double lLi = (R > 2) ? QW_X_I(s, (R > 2 ? R - 2 : 0)) : o0.y;
    qw_site<STAGE>(hl.x, hl.y, o0.x, o0.y, r0r, r0i, A, B, s.yr[0], s.yi[0], s.ur[0], s.ui[0],
                   s.ar[0], s.ai[0], s.yr[0], s.yi[0]);
    if (R > 1)
        qw_site<STAGE>(lLr, lLi, oL.x, oL.y, hr.x, hr.y, A, B, s.yr[R - 1], s.yi[R - 1],
                       s.ur[R - 1], s.ui[R - 1], s.ar[R - 1], s.ai[R - 1], s.yr[R - 1],
                       s.yi[R - 1]);
    if (OWN && threadIdx.x == 0) {       // oracle term d(t) x_w: marked vertex = site 0
        if (STAGE < 3) {
            s.ur[0] = fma(D.x, o0.y, s.ur[0]);
            s.ui[0] = fma(-D.x, o0.x, s.ui[0]);
            s.ar[0] = fma(D.y, o0.y, s.ar[0]);
            s.ai[0] = fma(-D.y, o0.x, s.ai[0]);
        } else {
            s.yr[0] = fma(D.y, o0.y, s.yr[0]);
            s.yi[0] = fma(-D.y, o0.x, s.yi[0]);
        }
    }
}

template <int R, int STAGE>
__device__ __forceinline__ void qw_interior(State<R> &s, const double A, const double B,
                                            const double2 o0, const double2 oL)
{
    double pr = o0.x, pi = o0.y;
#pragma unroll
    for (int r = 1; r < R - 1; ++r) {
        const double cr = QW_X_R(s, r), ci = QW_X_I(s, r);
        const double nr = (r + 1 < R - 1) ? QW_X_R(s, (r + 1 < R ? r + 1 : r)) : oL.x;
        const double ni = (r + 1 < R - 1) ? QW_X_I(s, (r + 1 < R ? r + 1 : r)) : oL.y;
        qw_site<STAGE>(pr, pi, cr, ci, nr, ni, A, B, s.yr[r], s.yi[r], s.ur[r], s.ui[r], s.ar[r],
                       s.ai[r], s.yr[r], s.yi[r]);
        pr = cr;
        pi = ci;
    }
}

struct CycleShared {
    double2 ab[4 * QW_CHUNK];
    double2 d[4 * QW_CHUNK];
    double2 ad[3 * QW_CHUNK];
    double red[2 * 2 * 32];
    uint64_t mbar;
};

// `cnt` RK4 steps with the coefficient table already in shared memory
// SPLIT: mbarrier arrive / wait (true) or one __syncthreads() per stage (false)
template <int R, bool OWN, bool SPLIT>
__device__ __forceinline__ void qw_cycle_steps(State<R> &s, CycleShared &sh, double2 *first,
                                               double2 *last, const int nt, const int t,
                                               const int tl, const int tr, const int cnt,
                                               int &par, uint32_t &phase)
{
    for (int sidx = 0; sidx < cnt; ++sidx) {
#define QW_V2_STAGE(K)                                                                  \
        {                                                                               \
            if (SPLIT) { qw_mbar_wait(&sh.mbar, phase & 1u); ++phase; }             \
            else __syncthreads();                                                       \
            const double2 hl = last[par * nt + tl];                                     \
            const double2 hr = first[par * nt + tr];                                    \
            const double2 ab = sh.ab[4 * sidx + K];                                     \
            double2 dd = make_double2(0.0, 0.0), o0, oL;                                \
            if (OWN && t == 0) dd = sh.d[4 * sidx + K];                                 \
            qw_edges<R, K, OWN>(s, ab.x, ab.y, hl, hr, dd, o0, oL);                     \
            par ^= 1;                                                                   \
            if (K < 3) {                                                                \
                first[par * nt + t] = make_double2(s.ur[0], s.ui[0]);                   \
                last[par * nt + t] = make_double2(s.ur[R - 1], s.ui[R - 1]);            \
            } else {                                                                    \
                first[par * nt + t] = make_double2(s.yr[0], s.yi[0]);                   \
                last[par * nt + t] = make_double2(s.yr[R - 1], s.yi[R - 1]);            \
            }                                                                           \
            if (SPLIT) {                                                                \
                __syncwarp();                                                           \
                if ((t & 31) == 0) qw_mbar_arrive(&sh.mbar);                            \
            }                                                                           \
            qw_interior<R, K>(s, ab.x, ab.y, o0, oL);                                   \
        }
        QW_V2_STAGE(0)
        QW_V2_STAGE(1)
        QW_V2_STAGE(2)
        QW_V2_STAGE(3)
#undef QW_V2_STAGE
    }
}
