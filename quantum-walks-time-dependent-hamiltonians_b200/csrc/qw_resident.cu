// Register-resident fixed-step RK4 for one (T, gamma) point per thread block (sm_100a).
//
// Replaces solve_schrodinger_equation + evaluate_probability (BCB.py:177-203) with the
// fixed-step rule of QW_Search.py:149-164.  psi, the current stage input and the RK4
// accumulator live in registers for the whole integration: each thread owns R consecutive
// sites of the ring (cycle graph) or of the vertex list (complete graph).  The graph is
// relabelled so that the marked vertex is site 0 of thread 0 (both graphs are vertex
// transitive and the initial state is uniform), which makes the oracle term a single
// predicated update with static register indices.
//
// Cycle graph: the only data that crosses threads per stage is one complex value on each
//   side of a thread's segment; it goes through a double-buffered shared-memory mailbox,
//   one __syncthreads() per stage.
// Complete graph: (L x)_j = N x_j - sum(x).  sum(x) of the step's first stage is a block
//   reduction; for the later stages it follows exactly from column sums of L being zero:
//   sum(y + c k) = sum(y) - i c d x_w, so thread 0 publishes it one stage ahead.
//
// FP64 work per site-update: 22 DFMA + 8 DADD (cycle), the algorithmic 52 flops of
// SURVEY.md section 8d.
#include "qw_common.cuh"
#include "qw_stage.cuh"
#include "qw_cycle_v2.cuh"

namespace {

struct Tables {
    double2 ab[4 * QW_CHUNK];
    double2 d[4 * QW_CHUNK];
    double2 ad[3 * QW_CHUNK];
    double red[2 * 2 * 32];      // double-buffered warp partials
    double2 bcast[2];            // complete graph: sum(x) of the coming stage
};

// all-threads block sum of a complex value; one barrier; buf alternates between calls
__device__ __forceinline__ double2 block_sum_all(double2 v, double *red, int buf)
{
    v.x = qw_warp_sum(v.x);
    v.y = qw_warp_sum(v.y);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nwarps = (blockDim.x + 31) >> 5;
    double *slot = red + buf * 64;
    if (lane == 0) { slot[2 * warp] = v.x; slot[2 * warp + 1] = v.y; }
    __syncthreads();
    double2 w = make_double2(0.0, 0.0);
    if (lane < nwarps) w = make_double2(slot[2 * lane], slot[2 * lane + 1]);
    w.x = qw_warp_sum(w.x);
    w.y = qw_warp_sum(w.y);
    return w;
}

template <int R>
__device__ __forceinline__ void epilogue(const State<R> &s, const QwParams &prm,
                                         const QwPoint &pt, double *red, const int t)
{
    double nrm = 0.0;
#pragma unroll
    for (int r = 0; r < R; ++r) nrm = fma(s.yr[r], s.yr[r], fma(s.yi[r], s.yi[r], nrm));
    double pw = (t == 0) ? fma(s.yr[0], s.yr[0], s.yi[0] * s.yi[0]) : 0.0;
    __syncthreads();
    double2 tot = block_sum_all(make_double2(nrm, pw), red, 0);
    if (t == 0) {
        prm.p[pt.q] = tot.y / tot.x;
        if (prm.norm) prm.norm[pt.q] = tot.x;
    }
    if (prm.psi && t < prm.nact) {      // undo the relabelling: logical l -> (l + w) mod n
        double2 *out = prm.psi + (size_t)pt.q * (size_t)prm.n;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            int64_t j = (int64_t)t * R + r + prm.target;
            if (j >= prm.n) j -= prm.n;
            out[j] = make_double2(s.yr[r], s.yi[r]);
        }
    }
}

template <int R, int MAXNT, int MINB>
__global__ void __launch_bounds__(MAXNT, MINB) rk4_cycle_resident(const QwParams prm)
{
    extern __shared__ double2 mbox[];         // [2][2][blockDim.x]: parity, first/last
    __shared__ Tables tb;
    const int t = threadIdx.x, nt = blockDim.x, nact = prm.nact;
    const QwPoint pt = qw_point(prm, blockIdx.x);
    const bool active = t < nact;
    const bool own = (t == 0);
    // ring neighbours among the active threads; idle threads talk to themselves (zeros)
    const int tl = active ? (t == 0 ? nact - 1 : t - 1) : t;
    const int tr = active ? (t == nact - 1 ? 0 : t + 1) : t;

    State<R> s;
    const double amp = active ? 1.0 / sqrt((double)prm.n) : 0.0;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        s.yr[r] = amp; s.yi[r] = 0.0;
        s.ur[r] = 0.0; s.ui[r] = 0.0; s.ar[r] = 0.0; s.ai[r] = 0.0;
    }
    int par = 0;
    double2 *first = mbox, *last = mbox + 2 * nt;      // [parity][thread]
    first[t] = make_double2(s.yr[0], s.yi[0]);
    last[t] = make_double2(s.yr[R - 1], s.yi[R - 1]);

    const double2 zero2 = make_double2(0.0, 0.0);
    for (int64_t n0 = 0; n0 < pt.steps; n0 += QW_CHUNK) {
        qw_fill_table(tb.ab, tb.d, tb.ad, pt, n0, prm.schedule, prm.gamma_on);
        const int cnt = (int)((pt.steps - n0 < QW_CHUNK) ? (pt.steps - n0) : QW_CHUNK);
        for (int sidx = 0; sidx < cnt; ++sidx) {
#define QW_CYCLE_STAGE(K)                                                               \
            {                                                                           \
                __syncthreads();                                                        \
                const double2 hl = last[par * nt + tl];                                 \
                const double2 hr = first[par * nt + tr];                                \
                const double2 ab = tb.ab[4 * sidx + K];                                 \
                const double2 dd = own ? tb.d[4 * sidx + K] : zero2;                    \
                rk_stage<R, K, true>(s, ab.x, ab.y, hl, hr, zero2, 0.0, own, dd);       \
                par ^= 1;                                                               \
                if (K < 3) {                                                            \
                    first[par * nt + t] = make_double2(s.ur[0], s.ui[0]);               \
                    last[par * nt + t] = make_double2(s.ur[R - 1], s.ui[R - 1]);        \
                } else {                                                                \
                    first[par * nt + t] = make_double2(s.yr[0], s.yi[0]);               \
                    last[par * nt + t] = make_double2(s.yr[R - 1], s.yi[R - 1]);        \
                }                                                                       \
            }
            QW_CYCLE_STAGE(0)
            QW_CYCLE_STAGE(1)
            QW_CYCLE_STAGE(2)
            QW_CYCLE_STAGE(3)
#undef QW_CYCLE_STAGE
        }
    }
    epilogue<R>(s, prm, pt, tb.red, t);
}

// Second-generation cycle kernel (qw_cycle_v2.cuh): edge sites first, early publication,
// mbarrier arrive / wait instead of a block barrier per stage, oracle term in warp 0 only.
template <int R, int MAXNT, int MINB, bool SPLIT>
__global__ void __launch_bounds__(MAXNT, MINB) rk4_cycle_resident_v2(const QwParams prm)
{
    extern __shared__ double2 mbox[];         // [2][2][blockDim.x]: parity, first/last
    __shared__ CycleShared sh;
    const int t = threadIdx.x, nt = blockDim.x, nact = prm.nact;
    const QwPoint pt = qw_point(prm, blockIdx.x);
    const bool active = t < nact;
    const int tl = active ? (t == 0 ? nact - 1 : t - 1) : t;
    const int tr = active ? (t == nact - 1 ? 0 : t + 1) : t;

    State<R> s;
    const double amp = active ? 1.0 / sqrt((double)prm.n) : 0.0;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        s.yr[r] = amp; s.yi[r] = 0.0;
        s.ur[r] = 0.0; s.ui[r] = 0.0; s.ar[r] = 0.0; s.ai[r] = 0.0;
    }
    if (SPLIT) {
        if (t == 0) qw_mbar_init(&sh.mbar, nt >> 5);
        __syncthreads();
    }
    int par = 0;
    uint32_t phase = 0;
    double2 *first = mbox, *last = mbox + 2 * nt;
    first[t] = make_double2(s.yr[0], s.yi[0]);
    last[t] = make_double2(s.yr[R - 1], s.yi[R - 1]);
    if (SPLIT) {
        __syncwarp();
        if ((t & 31) == 0) qw_mbar_arrive(&sh.mbar);
    }

    for (int64_t n0 = 0; n0 < pt.steps; n0 += QW_CHUNK) {
        qw_fill_table(sh.ab, sh.d, sh.ad, pt, n0, prm.schedule, prm.gamma_on);
        const int cnt = (int)((pt.steps - n0 < QW_CHUNK) ? (pt.steps - n0) : QW_CHUNK);
        if (SPLIT) {
            if (t < 32) qw_cycle_steps<R, true, SPLIT>(s, sh, first, last, nt, t, tl, tr, cnt, par, phase);
            else qw_cycle_steps<R, false, SPLIT>(s, sh, first, last, nt, t, tl, tr, cnt, par, phase);
        } else {
            qw_cycle_steps<R, true, SPLIT>(s, sh, first, last, nt, t, tl, tr, cnt, par, phase);
        }
    }
    epilogue<R>(s, prm, pt, sh.red, t);
}

template <int R>
__device__ __forceinline__ double2 local_sum_y(const State<R> &s)
{
    double a = 0.0, b = 0.0;
#pragma unroll
    for (int r = 0; r < R; ++r) { a += s.yr[r]; b += s.yi[r]; }
    return make_double2(a, b);
}

template <int R>
__device__ __forceinline__ double2 local_sum_u(const State<R> &s)
{
    double a = 0.0, b = 0.0;
#pragma unroll
    for (int r = 0; r < R; ++r) { a += s.ur[r]; b += s.ui[r]; }
    return make_double2(a, b);
}

template <int R, int MAXNT, int MINB, bool EXACT>
__global__ void __launch_bounds__(MAXNT, MINB) rk4_complete_resident(const QwParams prm)
{
    __shared__ Tables tb;
    const int t = threadIdx.x, nact = prm.nact;
    const QwPoint pt = qw_point(prm, blockIdx.x);
    const bool active = t < nact;
    const bool own = (t == 0);
    const double nd = active ? (double)prm.n : 0.0;   // idle threads stay at zero

    State<R> s;
    const double amp = active ? 1.0 / sqrt((double)prm.n) : 0.0;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        s.yr[r] = amp; s.yi[r] = 0.0;
        s.ur[r] = 0.0; s.ui[r] = 0.0; s.ar[r] = 0.0; s.ai[r] = 0.0;
    }
    int rbuf = 0;
    const double2 zero2 = make_double2(0.0, 0.0);
    double2 sy = block_sum_all(local_sum_y<R>(s), tb.red, rbuf);   // sum(y), every thread
    rbuf ^= 1;

    for (int64_t n0 = 0; n0 < pt.steps; n0 += QW_CHUNK) {
        qw_fill_table(tb.ab, tb.d, tb.ad, pt, n0, prm.schedule, prm.gamma_on);
        const int cnt = (int)((pt.steps - n0 < QW_CHUNK) ? (pt.steps - n0) : QW_CHUNK);
        for (int sidx = 0; sidx < cnt; ++sidx) {
#define QW_COMPLETE_STAGE(K)                                                            \
            {                                                                           \
                double2 sx;                                                             \
                if (K == 0) sx = sy;                                                    \
                else if (EXACT) { sx = block_sum_all(local_sum_u<R>(s), tb.red, rbuf); rbuf ^= 1; } \
                else { __syncthreads(); sx = tb.bcast[K & 1]; }                         \
                if (!active) sx = zero2;                                                \
                const double2 ab = tb.ab[4 * sidx + K];                                 \
                const double2 dd = own ? tb.d[4 * sidx + K] : zero2;                    \
                if (!EXACT && own && K < 3) {                                           \
                    /* sum of the next stage input: sum(y) - i (c d) x_w */             \
                    const double xr = (K == 0) ? s.yr[0] : s.ur[0];                     \
                    const double xi = (K == 0) ? s.yi[0] : s.ui[0];                     \
                    tb.bcast[(K + 1) & 1] = make_double2(fma(dd.x, xi, sy.x),           \
                                                         fma(-dd.x, xr, sy.y));         \
                }                                                                       \
                rk_stage<R, K, false>(s, ab.x, ab.y, zero2, zero2, sx, nd, own, dd);    \
            }
            QW_COMPLETE_STAGE(0)
            QW_COMPLETE_STAGE(1)
            QW_COMPLETE_STAGE(2)
            QW_COMPLETE_STAGE(3)
#undef QW_COMPLETE_STAGE
            sy = block_sum_all(local_sum_y<R>(s), tb.red, rbuf);
            rbuf ^= 1;
        }
    }
    epilogue<R>(s, prm, pt, tb.red, t);
}

// Second-generation complete-graph kernel.  No block reduction sits on the critical path:
//   * thread 0 carries an estimate S of sum(y) and advances it with the exact identity
//     sum(y + c k) = sum(y) - i c d x_w (column sums of L vanish); it publishes the sum every
//     stage needs one stage ahead (16 B broadcast, one __syncthreads per stage);
//   * at the end of each step every warp leaves the direct sum of its part of y_{n+1} in
//     shared memory (local adds + one warp butterfly, nobody waits for it); during step n+1
//     warp 0 adds the partials and thread 0 folds the difference to its estimate into the
//     sum of step n+2.  The estimate is therefore re-anchored to a direct summation every
//     step with a lag of one step: rounding cannot accumulate, and the reduction latency is
//     hidden.
//   * the oracle term is a separate instantiation of the stage code run by warp 0 only.
template <int R, int MAXNT, int MINB>
__global__ void __launch_bounds__(MAXNT, MINB) rk4_complete_resident_v2(const QwParams prm)
{
    __shared__ Tables tb;
    const int t = threadIdx.x, nact = prm.nact;
    const int warp = t >> 5, lane = t & 31, nwarps = (blockDim.x + 31) >> 5;
    const QwPoint pt = qw_point(prm, blockIdx.x);
    const bool active = t < nact;
    const double nd = active ? (double)prm.n : 0.0;   // idle threads stay at zero
    const double2 zero2 = make_double2(0.0, 0.0);

    State<R> s;
    const double amp = active ? 1.0 / sqrt((double)prm.n) : 0.0;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        s.yr[r] = amp; s.yi[r] = 0.0;
        s.ur[r] = 0.0; s.ui[r] = 0.0; s.ar[r] = 0.0; s.ai[r] = 0.0;
    }
    // warp partials of sum(y): tb.red[parity][2*warp .. 2*warp+1]
    int spar = 0;
    {
        double2 loc = local_sum_y<R>(s);
        loc.x = qw_warp_sum(loc.x);
        loc.y = qw_warp_sum(loc.y);
        if (lane == 0) { tb.red[2 * warp] = loc.x; tb.red[2 * warp + 1] = loc.y; }
    }
    __syncthreads();
    double2 shat = zero2, delta = zero2, sinc = zero2;      // meaningful in thread 0 only
    if (t == 0) {
        for (int w = 0; w < nwarps; ++w) { shat.x += tb.red[2 * w]; shat.y += tb.red[2 * w + 1]; }
        tb.bcast[0] = shat;
    }

    for (int64_t n0 = 0; n0 < pt.steps; n0 += QW_CHUNK) {
        qw_fill_table(tb.ab, tb.d, tb.ad, pt, n0, prm.schedule, prm.gamma_on);
        const int cnt = (int)((pt.steps - n0 < QW_CHUNK) ? (pt.steps - n0) : QW_CHUNK);
        for (int sidx = 0; sidx < cnt; ++sidx) {
#define QW_COMPLETE2_STAGE(K)                                                           \
            {                                                                           \
                __syncthreads();                                                        \
                double2 sx = tb.bcast[K & 1];                                           \
                if (!active) sx = zero2;                                                \
                const double2 ab = tb.ab[4 * sidx + K];                                 \
                if (warp == 0) {                                                        \
                    const double2 dd = (t == 0) ? tb.d[4 * sidx + K] : zero2;           \
                    if (K == 0) {                                                       \
                        /* direct sum of this step's y from the warp partials */        \
                        double2 dsum = zero2;                                           \
                        if (lane < nwarps)                                              \
                            dsum = make_double2(tb.red[spar * 64 + 2 * lane],           \
                                                tb.red[spar * 64 + 2 * lane + 1]);      \
                        dsum.x = qw_warp_sum(dsum.x);                                   \
                        dsum.y = qw_warp_sum(dsum.y);                                   \
                        delta = make_double2(dsum.x - shat.x, dsum.y - shat.y);         \
                        sinc = zero2;                                                   \
                    }                                                                   \
                    if (t == 0) {                                                       \
                        const double xr = (K == 0) ? s.yr[0] : s.ur[0];                 \
                        const double xi = (K == 0) ? s.yi[0] : s.ui[0];                 \
                        sinc.x = fma(dd.y, xi, sinc.x);                                 \
                        sinc.y = fma(-dd.y, xr, sinc.y);                                \
                        if (K < 3) {                                                    \
                            tb.bcast[(K + 1) & 1] = make_double2(fma(dd.x, xi, shat.x), \
                                                                 fma(-dd.x, xr, shat.y)); \
                        } else {                                                        \
                            shat.x += sinc.x + delta.x;                                 \
                            shat.y += sinc.y + delta.y;                                 \
                            tb.bcast[0] = shat;                                         \
                        }                                                               \
                    }                                                                   \
                    rk_stage<R, K, false>(s, ab.x, ab.y, zero2, zero2, sx, nd, t == 0, dd); \
                } else {                                                                \
                    rk_stage<R, K, false>(s, ab.x, ab.y, zero2, zero2, sx, nd, false, zero2); \
                }                                                                       \
            }
            QW_COMPLETE2_STAGE(0)
            QW_COMPLETE2_STAGE(1)
            QW_COMPLETE2_STAGE(2)
            QW_COMPLETE2_STAGE(3)
#undef QW_COMPLETE2_STAGE
            // leave this warp's part of sum(y_{n+1}) for warp 0 to pick up next step
            spar ^= 1;
            double2 loc = local_sum_y<R>(s);
            loc.x = qw_warp_sum(loc.x);
            loc.y = qw_warp_sum(loc.y);
            if (lane == 0) {
                tb.red[spar * 64 + 2 * warp] = loc.x;
                tb.red[spar * 64 + 2 * warp + 1] = loc.y;
            }
        }
    }
    __syncthreads();
    epilogue<R>(s, prm, pt, tb.red, t);
}

// Barrier-free variant of the complete-graph kernel (QW_FLAG_COMPLETE_FREE; A/B record): no
// block barrier inside a 32-step chunk.  Measured on C3 (N = 4096): 4.22e11 site-updates/s
// at R = 16, the same as the one-barrier-per-stage kernel above (4.23e11), so that one stays
// the default -- with 8 warps per SM the limit is FP64 dependency latency, not the barrier.
//
// The only value a stage needs from outside its thread is sum(x) of the stage input, and
// that follows exactly from the marked vertex alone (column sums of L vanish):
//     sum(y + c k) = sum(y) - i c d x_w,     sum(y_{n+1}) = sum(y_n) - i sum_m b_m d_m x_{m,w}.
// Thread 0 owns the marked vertex, so warp 0 is the producer of all 128 stage sums of a chunk
// and writes each into its own slot of a shared table one stage before it is needed, followed
// by a release of the `published` counter; the other warps acquire the counter (a bounded
// spin on shared memory that is almost always satisfied at the first read) instead of meeting
// at a __syncthreads().  Warps of a block are thereby decoupled: nobody waits for the slowest
// warp four times per step, only for warp 0 if it falls behind.
// sum(y) is carried in double-double by thread 0 (no rounding drift from the running
// accumulation) and re-anchored to a direct block reduction of the state at every chunk
// start, where the coefficient-table refill synchronises the block anyway.
struct FreeShared {
    double2 sums[4 * QW_CHUNK + 1];     // sums[g]: sum of the input of stage g of the chunk
    double2 hi, lo, inc;                // thread 0: double-double sum(y), per-step increment
    int published;                      // sums[0 .. published-1] are valid
    int failed;
};

__device__ __forceinline__ int ld_acquire_shared(const int *p)
{
    int v;
    asm volatile("ld.acquire.cta.shared.s32 %0, [%1];" : "=r"(v)
                 : "r"((uint32_t)__cvta_generic_to_shared(p)) : "memory");
    return v;
}

__device__ __forceinline__ void st_release_shared(int *p, int v)
{
    asm volatile("st.release.cta.shared.s32 [%0], %1;" ::"r"(
                     (uint32_t)__cvta_generic_to_shared(p)), "r"(v) : "memory");
}

// blocking: spin (bounded) until sums[g] has been published
__device__ __forceinline__ double2 acquire_sum(FreeShared &fs, int g)
{
    int spins = 0;
    while (ld_acquire_shared(&fs.published) <= g) {           // bounded: safety net only
        if (*((volatile int *)&fs.failed)) break;
        if (++spins > (1 << 22)) { fs.failed = 1; break; }
    }
    return fs.sums[g];
}

// non-blocking: sums[g] if it is already there (ok = true); issued one stage ahead so that
// the shared-memory latency hides behind the current stage's arithmetic
__device__ __forceinline__ double2 try_sum(FreeShared &fs, int g, bool &ok)
{
    ok = ld_acquire_shared(&fs.published) > g;
    return fs.sums[g];
}

template <int R, int MAXNT, int MINB>
__global__ void __launch_bounds__(MAXNT, MINB) rk4_complete_resident_v3(const QwParams prm)
{
    __shared__ Tables tb;
    __shared__ FreeShared fs;
    const int t = threadIdx.x, nact = prm.nact;
    const int warp = t >> 5;
    const QwPoint pt = qw_point(prm, blockIdx.x);
    const bool active = t < nact;
    const double nd = active ? (double)prm.n : 0.0;   // idle threads stay at zero
    const double2 zero2 = make_double2(0.0, 0.0);

    State<R> s;
    const double amp = active ? 1.0 / sqrt((double)prm.n) : 0.0;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        s.yr[r] = amp; s.yi[r] = 0.0;
        s.ur[r] = 0.0; s.ui[r] = 0.0; s.ar[r] = 0.0; s.ai[r] = 0.0;
    }
    if (t == 0) fs.failed = 0;
    int rbuf = 0;

    for (int64_t n0 = 0; n0 < pt.steps; n0 += QW_CHUNK) {
        __syncthreads();                               // every warp has left the last chunk
        const double2 sy = block_sum_all(local_sum_y<R>(s), tb.red, rbuf);    // re-anchor
        rbuf ^= 1;
        if (t == 0) {
            fs.hi = sy; fs.lo = zero2;
            fs.sums[0] = sy;
            fs.published = 1;
        }
        qw_fill_table(tb.ab, tb.d, tb.ad, pt, n0, prm.schedule, prm.gamma_on);
        const int cnt = (int)((pt.steps - n0 < QW_CHUNK) ? (pt.steps - n0) : QW_CHUNK);
        bool pf_ok = false;                            // prefetched sum of the coming stage
        double2 pf = zero2;
        for (int sidx = 0; sidx < cnt; ++sidx) {
#define QW_COMPLETE3_STAGE(K)                                                           \
            {                                                                           \
                const int g = 4 * sidx + K;                                             \
                double2 sx = pf_ok ? pf : acquire_sum(fs, g);                           \
                if (!active) sx = zero2;                                                \
                const double2 ab = tb.ab[g];                                            \
                if (warp == 0) {                                                        \
                    const double2 dd = (t == 0) ? tb.d[g] : zero2;                      \
                    if (t == 0) {                                                       \
                        const double xr = (K == 0) ? s.yr[0] : s.ur[0];                 \
                        const double xi = (K == 0) ? s.yi[0] : s.ui[0];                 \
                        double2 inc = (K == 0) ? zero2 : fs.inc;                        \
                        inc.x = fma(dd.y, xi, inc.x);                                   \
                        inc.y = fma(-dd.y, xr, inc.y);                                  \
                        double2 hi = fs.hi;                                             \
                        double2 nxt;                                                    \
                        if (K < 3) {                                                    \
                            fs.inc = inc;                                               \
                            nxt = make_double2(fma(dd.x, xi, hi.x), fma(-dd.x, xr, hi.y)); \
                        } else {          /* sum(y_{n+1}) = (hi + lo) + inc, TwoSum */   \
                            double2 lo = fs.lo;                                         \
                            const double sxr = hi.x + inc.x, bxr = sxr - hi.x;          \
                            lo.x += (hi.x - (sxr - bxr)) + (inc.x - bxr);               \
                            const double sxi = hi.y + inc.y, bxi = sxi - hi.y;          \
                            lo.y += (hi.y - (sxi - bxi)) + (inc.y - bxi);               \
                            hi = make_double2(sxr + lo.x, sxi + lo.y);                  \
                            lo = make_double2(lo.x - (hi.x - sxr), lo.y - (hi.y - sxi)); \
                            fs.hi = hi; fs.lo = lo;                                     \
                            nxt = hi;                                                   \
                        }                                                               \
                        fs.sums[g + 1] = nxt;                                           \
                        st_release_shared(&fs.published, g + 2);                        \
                    }                                                                   \
                    __syncwarp();                                                       \
                    pf = try_sum(fs, g + 1, pf_ok);                                     \
                    rk_stage<R, K, false>(s, ab.x, ab.y, zero2, zero2, sx, nd, t == 0, dd); \
                } else {                                                                \
                    pf = try_sum(fs, g + 1, pf_ok);                                     \
                    rk_stage<R, K, false>(s, ab.x, ab.y, zero2, zero2, sx, nd, false, zero2); \
                }                                                                       \
            }
            QW_COMPLETE3_STAGE(0)
            QW_COMPLETE3_STAGE(1)
            QW_COMPLETE3_STAGE(2)
            QW_COMPLETE3_STAGE(3)
#undef QW_COMPLETE3_STAGE
        }
    }
    __syncthreads();
    if (fs.failed) {                                    // spin limit hit: poison the result
#pragma unroll
        for (int r = 0; r < R; ++r) s.yr[r] = NAN;
    }
    epilogue<R>(s, prm, pt, tb.red, t);
}

template <typename K>
cudaError_t launch(K kernel, const QwParams &prm, int threads, size_t smem, cudaStream_t st,
                   qw_plan *plan)
{
    if (plan) {
        cudaFuncAttributes fa;
        cudaError_t e = cudaFuncGetAttributes(&fa, kernel);
        if (e != cudaSuccess) return e;
        int nb = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, threads, smem);
        if (e != cudaSuccess) return e;
        plan->threads_per_block = threads;
        plan->blocks_per_sm = nb;
        plan->regs_per_thread = fa.numRegs;
        plan->smem_bytes = (int)(fa.sharedSizeBytes + smem);
        return cudaSuccess;
    }
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)smem);
        if (e != cudaSuccess) return e;
    }
    kernel<<<(unsigned)prm.n_points, threads, smem, st>>>(prm);
    return cudaGetLastError();
}

template <int R>
cudaError_t dispatch_threads(const QwParams &prm, int threads, cudaStream_t st, qw_plan *plan)
{
    const bool cyc = prm.topology == QW_TOPO_CYCLE;
    const bool exact = (prm.flags & QW_FLAG_EXACT_SUM) != 0;
    const size_t smem = cyc ? (size_t)threads * 4 * sizeof(double2) : 0;
    const bool legacy = (prm.flags & QW_FLAG_CYCLE_LEGACY) != 0;
    const bool split = (prm.flags & QW_FLAG_CYCLE_SPLIT) != 0;
    // The first-generation and mbarrier variants exist for the A/B record only (R = 8).
#define QW_LAUNCH(MAXNT, MINB)                                                             \
    if constexpr (R == 8) {                                                                \
        if (cyc && legacy)                                                                 \
            return launch(rk4_cycle_resident<R, MAXNT, MINB>, prm, threads, smem, st, plan); \
        if (cyc && split)                                                                  \
            return launch(rk4_cycle_resident_v2<R, MAXNT, MINB, true>, prm, threads, smem, st, \
                          plan);                                                           \
        if (!cyc && !exact && legacy)                                                      \
            return launch(rk4_complete_resident<R, MAXNT, MINB, false>, prm, threads, smem, st, \
                          plan);                                                           \
    }                                                                                      \
    if (cyc)                                                                               \
        return launch(rk4_cycle_resident_v2<R, MAXNT, MINB, false>, prm, threads, smem, st, plan); \
    if (exact)                                                                             \
        return launch(rk4_complete_resident<R, MAXNT, MINB, true>, prm, threads, smem, st, plan); \
    if constexpr (R == 8) {                                                                \
        if (prm.flags & QW_FLAG_COMPLETE_FREE)                                             \
            return launch(rk4_complete_resident_v3<R, MAXNT, MINB>, prm, threads, smem, st, \
                          plan);                                                           \
    }                                                                                      \
    return launch(rk4_complete_resident_v2<R, MAXNT, MINB>, prm, threads, smem, st, plan);
    // <= 128 threads: 3 blocks/SM (168 registers, no spills) measured faster than 4 (128)
    if (threads <= 128 && (prm.flags & QW_FLAG_OCC4)) { QW_LAUNCH(128, 4) }
    if (threads <= 128) { QW_LAUNCH(128, 3) }
    if (threads <= 256) { QW_LAUNCH(256, 2) }
    if (threads <= 512) { QW_LAUNCH(512, 1) }
    if constexpr (R <= 2) { QW_LAUNCH(1024, 1) }      // 64-register cap: R <= 2 only
    return cudaErrorInvalidConfiguration;
#undef QW_LAUNCH
}

// R = 16: twice the sites per thread, half the threads (N = 1024 -> 2 warps per point,
// N = 4096 -> 8 warps).  Fewer exchanges and barriers per FP64 instruction and 32 independent
// chains per thread, at the price of ~250 registers.  Measured +3 % (cycle N = 1024) and
// +17 % (complete N = 4096) over R = 8; default wherever N % 16 == 0 and N/16 >= 32
// (QW_FLAG_R8 keeps R = 8 for A/B runs).
cudaError_t dispatch_r16(const QwParams &prm, int threads, cudaStream_t st, qw_plan *plan)
{
    const bool cyc = prm.topology == QW_TOPO_CYCLE;
    const size_t smem = cyc ? (size_t)threads * 4 * sizeof(double2) : 0;
    if (cyc) {
        if (threads <= 64)
            return launch(rk4_cycle_resident_v2<16, 64, 4, false>, prm, threads, smem, st, plan);
        if (threads <= 128)
            return launch(rk4_cycle_resident_v2<16, 128, 2, false>, prm, threads, smem, st, plan);
        if (threads <= 256)
            return launch(rk4_cycle_resident_v2<16, 256, 1, false>, prm, threads, smem, st, plan);
    } else if (threads <= 256) {
        if (prm.flags & QW_FLAG_COMPLETE_FREE)
            return launch(rk4_complete_resident_v3<16, 256, 1>, prm, threads, smem, st, plan);
        return launch(rk4_complete_resident_v2<16, 256, 1>, prm, threads, smem, st, plan);
    }
    return cudaErrorInvalidConfiguration;
}

}  // namespace

// Sites per thread for the resident path, or 0 when n does not fit it.  R*threads == n
// exactly; the largest R that still fills a warp wins (more sites per thread = fewer
// cross-thread exchanges per FP64 instruction).  Register budget: R = 8 and 4 need up to
// 128 registers, so at most 512 threads; R <= 2 fits the 64-register cap of 1024 threads.
int qw_resident_sites_per_thread(int64_t n)
{
    if (n < 32) return n >= 1 ? 1 : 0;
    const int cand[4] = {8, 4, 2, 1};
    for (int i = 0; i < 4; ++i) {
        const int R = cand[i];
        if (n % R == 0 && n / R >= 32 && n / R <= (R >= 4 ? 512 : 1024)) return R;
    }
    return 0;
}

cudaError_t qw_launch_resident(QwParams prm, cudaStream_t st, qw_plan *plan)
{
    int R = qw_resident_sites_per_thread(prm.n);
    if (!(prm.flags & QW_FLAG_R8) && prm.n % 16 == 0 && prm.n / 16 >= 32 && prm.n / 16 <= 256)
        R = 16;
    prm.nact = (int)(prm.n / R);
    const int threads = ((prm.nact + 31) / 32) * 32;
    if (plan) {
        plan->path = prm.topology == QW_TOPO_CYCLE ? 0 : 1;
        plan->sites_per_thread = R;
    }
    switch (R) {
    case 16: return dispatch_r16(prm, threads, st, plan);
    case 8: return dispatch_threads<8>(prm, threads, st, plan);
    case 4: return dispatch_threads<4>(prm, threads, st, plan);
    case 2: return dispatch_threads<2>(prm, threads, st, plan);
    default: return dispatch_threads<1>(prm, threads, st, plan);
    }
}
