// Tile-fused HBM-streaming RK4 for rings too large for one SM's registers (cycle graph,
// N = 65 536 in BASELINE config 5), sm_100a.
//
// psi of every point in flight lives in HBM (two buffers, ping-pong).  One launch ("pass")
// advances every point by KB RK4 steps.  Work unit = one window of W = 1024 consecutive
// sites of one point (64 threads x 16 sites, or 128 x 8): load it (cp.async, coalesced 16 B per lane, into a padded shared
// buffer), run the register-resident stage code of qw_resident.cu on it with open ends,
// write back the W - 2*HALO central sites (HALO = 4*KB: every stage invalidates one more
// site at each open end, so all four stages of KB steps are fused per window).  Thread
// blocks are persistent and software-pipelined: while a block integrates window i its
// cp.async loads for the next windows are already in flight into the other slots of a
// shared-memory ring, so HBM latency is not exposed once per window.
//
// Algorithmic traffic: one read and one write of psi per pass = 32 B per site per pass,
// i.e. 32/KB bytes per site-update.  KB = 1 is the plain "read psi_n, write psi_n+1"
// streaming roofline of SURVEY.md section 8d; KB = 4 (default) is temporal blocking, which
// makes the path FP64-bound again.
//
// Replaces the same reference path as the resident kernels (BCB.py:163-203 with the step
// rule of QW_Search.py:149-164); vertices keep their physical labels, so the marked vertex
// can sit in any slot of any window (including the redundant halo copies).
#include <type_traits>

#include "qw_stage.cuh"
#include "qw_horner_level.cuh"

namespace {

// Geometry for SR sites per thread; windows are 1024 sites either way.
//   SR = 16: 64 threads (2 warps) per block, 4 blocks/SM -- the geometry of the fastest
//            resident kernel: a stage barrier involves two warps only;
//   SR = 8:  128 threads, 3 blocks/SM.
template <int SR>
struct Geo {
    static constexpr int NT = SR == 16 ? 64 : 128;     // threads per block
    static constexpr int W = SR * NT;                  // window
    static constexpr int PAD = W + W / SR;             // one pad element per thread segment
    static constexpr int BLOCKS_PER_SM = SR == 16 ? 4 : 3;
    // ring of shared window buffers: SR = 8 keeps two windows in flight behind the one being
    // integrated (HBM-bound KB = 1 passes); SR = 16 one (4 blocks/SM must fit 228 KB)
    static constexpr int DEPTH = SR == 16 ? 2 : 3;
    static __device__ __forceinline__ int pad(int e) { return e + e / SR; }
};

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(
                     (uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int PENDING>
__device__ __forceinline__ void cp_async_wait_pending()
{
    asm volatile("cp.async.wait_group %0;" ::"n"(PENDING) : "memory");
}

// Per point and per pass, computed once by stream_pass_tables instead of once per window:
// the step plan of the point and the stage-weighted operator scalars of the KB steps.
constexpr int MAXKB = 8;               // most steps fused per pass
struct PassEntry {
    double2 ab[4 * MAXKB];             // (A, B) per stage
    double2 d[4 * MAXKB];              // (DA, DB) per stage
    int32_t kb;                        // steps this point still takes in this pass
    int32_t pad[3];
};

// The same for the nested (Horner) form of the step (qw_horner_level.cuh): r1..r4 and the ten
// correction coefficients over (y_w, s1, s2, s3) per step.
struct HPassEntry {
    double rho[4 * MAXKB];
    double2 c[10 * MAXKB];
    int32_t kb;
    int32_t pad[3];
};

template <int SR, bool HORNER>
struct StreamShared {
    using Entry = typename std::conditional<HORNER, HPassEntry, PassEntry>::type;
    double2 xbuf[Geo<SR>::DEPTH][Geo<SR>::PAD];   // window ring (load target, store staging)
    Entry tab[Geo<SR>::DEPTH];            // the point's pass entry travels with its window
    double2 mbox[4 * Geo<SR>::NT];        // [parity][first/last][thread]
};

__global__ void stream_pass_tables_horner(const QwParams prm, int64_t count, int64_t step0,
                                          int kbmax, HPassEntry *__restrict__ tab)
{
    const int64_t local = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (local >= count) return;
    const QwPoint pt = qw_point(prm, prm.point_offset + local);
    const int64_t todo = pt.steps - step0;
    const int kb = todo < 0 ? 0 : (todo > kbmax ? kbmax : (int)todo);
    HPassEntry &e = tab[local];
    e.kb = kb;
    const double h = pt.h;
    for (int s = 0; s < kb; ++s) {
        const double t0 = (double)(step0 + s) * h;
        const double2 c0 = qw_operator_scalars(t0 / pt.T, pt.gamma, prm.schedule, prm.gamma_on);
        const double2 c1 = qw_operator_scalars((t0 + 0.5 * h) / pt.T, pt.gamma, prm.schedule,
                                               prm.gamma_on);
        const double2 c2 = qw_operator_scalars((t0 + h) / pt.T, pt.gamma, prm.schedule,
                                               prm.gamma_on);
        HornerCoef hc;
        horner_coefficients(hc, h, c0, c1, c2, 2.0, 6.0);
        horner_cycle_sum_basis(hc.c);
        for (int k = 0; k < 4; ++k) e.rho[4 * s + k] = hc.rho[k];
        for (int k = 0; k < 10; ++k) e.c[10 * s + k] = hc.c[k];
    }
}

__global__ void stream_pass_tables(const QwParams prm, int64_t count, int64_t step0, int kbmax,
                                   PassEntry *__restrict__ tab)
{
    const int64_t local = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (local >= count) return;
    const QwPoint pt = qw_point(prm, prm.point_offset + local);
    const int64_t todo = pt.steps - step0;
    const int kb = todo < 0 ? 0 : (todo > kbmax ? kbmax : (int)todo);
    PassEntry &e = tab[local];
    e.kb = kb;
    const double h = pt.h;
    for (int s = 0; s < kb; ++s) {
        const double t0 = (double)(step0 + s) * h;
        const double2 c0 = qw_operator_scalars(t0 / pt.T, pt.gamma, prm.schedule, prm.gamma_on);
        const double2 c1 = qw_operator_scalars((t0 + 0.5 * h) / pt.T, pt.gamma, prm.schedule,
                                               prm.gamma_on);
        const double2 c2 = qw_operator_scalars((t0 + h) / pt.T, pt.gamma, prm.schedule,
                                               prm.gamma_on);
        e.ab[4 * s + 0] = make_double2(0.5 * h * c0.x, h / 6.0 * c0.x);
        e.d[4 * s + 0] = make_double2(0.5 * h * c0.y, h / 6.0 * c0.y);
        e.ab[4 * s + 1] = make_double2(0.5 * h * c1.x, h / 3.0 * c1.x);
        e.d[4 * s + 1] = make_double2(0.5 * h * c1.y, h / 3.0 * c1.y);
        e.ab[4 * s + 2] = make_double2(h * c1.x, h / 3.0 * c1.x);
        e.d[4 * s + 2] = make_double2(h * c1.y, h / 3.0 * c1.y);
        e.ab[4 * s + 3] = make_double2(0.0 * c2.x, h / 6.0 * c2.x);
        e.d[4 * s + 3] = make_double2(0.0 * c2.y, h / 6.0 * c2.y);
    }
}

template <int SR, int STAGE>
__device__ __forceinline__ void read_slot(const State<SR> &s, int slot, double &xr, double &xi)
{
#pragma unroll
    for (int r = 0; r < SR; ++r)
        if (r == slot) {
            xr = (STAGE == 0) ? s.yr[r] : s.ur[r];
            xi = (STAGE == 0) ? s.yi[r] : s.ui[r];
        }
}

// oracle term d(t) x_w applied to whichever slot holds the marked vertex
template <int SR, int STAGE>
__device__ __forceinline__ void fix_slot(State<SR> &s, int slot, double2 D, double xr, double xi)
{
#pragma unroll
    for (int r = 0; r < SR; ++r)
        if (r == slot) {
            if (STAGE < 3) {
                s.ur[r] = fma(D.x, xi, s.ur[r]);
                s.ui[r] = fma(-D.x, xr, s.ui[r]);
                s.ar[r] = fma(D.y, xi, s.ar[r]);
                s.ai[r] = fma(-D.y, xr, s.ai[r]);
            } else {
                s.yr[r] = fma(D.y, xi, s.yr[r]);
                s.yi[r] = fma(-D.y, xr, s.yi[r]);
            }
        }
}

template <int SR, int STAGE>
__device__ __forceinline__ void fix_slot0(State<SR> &s, double2 D, double xr, double xi)
{
    if (STAGE < 3) {
        s.ur[0] = fma(D.x, xi, s.ur[0]);
        s.ui[0] = fma(-D.x, xr, s.ui[0]);
        s.ar[0] = fma(D.y, xi, s.ar[0]);
        s.ai[0] = fma(-D.y, xr, s.ai[0]);
    } else {
        s.yr[0] = fma(D.y, xi, s.yr[0]);
        s.yi[0] = fma(-D.y, xr, s.yi[0]);
    }
}

// issue the loads of window `idx` (point idx / tiles, tile idx % tiles) and of its point's
// pass entry into ring slot `slot`; one cp.async group
// HALO: sites invalidated at each open end by KB steps (4 per step).  In the ALIGNED variant
// it is rounded up to a multiple of SR so that every window starts on a thread-segment
// boundary of the (rotated) ring: the marked vertex, stored at logical site 0, is then always
// slot 0 of some thread and the oracle term needs no slot search.
template <int KB, int SR, bool ALIGNED>
struct Win {
    static constexpr int HALO = ALIGNED ? ((4 * KB + SR - 1) / SR) * SR : 4 * KB;
    static constexpr int TILE = Geo<SR>::W - 2 * HALO;
};

template <int KB, int SR, bool ALIGNED, bool HORNER>
__device__ __forceinline__ void prefetch_window(StreamShared<SR, HORNER> &sh, int slot,
                                                const double2 *in,
                                                const typename StreamShared<SR, HORNER>::Entry *tab,
                                                int64_t n, uint32_t idx, uint32_t tiles, int t,
                                                int64_t mirror_m = 0, bool even = true)
{
    using Entry = typename StreamShared<SR, HORNER>::Entry;
    static_assert(sizeof(Entry) % 16 == 0 && KB <= MAXKB, "entry copy");
    constexpr int HALO = Win<KB, SR, ALIGNED>::HALO, TILE = Win<KB, SR, ALIGNED>::TILE,
                  SNT = Geo<SR>::NT;
    const uint32_t local = idx / tiles, tile = idx - local * tiles;
    const double2 *src = in + (int64_t)local * n;
    double2 *xb = sh.xbuf[slot];
    if (mirror_m) {
        // half ring behind 3 ghost sites: stored index i <-> chain site c = i - 3; the halo
        // beyond either end is the mirror image (c < 0 -> -c; c > M-1 -> about the antipode)
        const int64_t first = (int64_t)tile * TILE - HALO;
#pragma unroll
        for (int i = 0; i < SR; ++i) {
            const int e = i * SNT + t;
            int64_t c = first + e - 3;
            if (c < 0) c = -c;
            if (c > mirror_m - 1) c = 2 * (mirror_m - 1) + (even ? 0 : 1) - c;
            cp_async16(&xb[Geo<SR>::pad(e)], &src[c + 3]);
        }
    } else {
    int64_t base = ((int64_t)tile * TILE - HALO) % n;
    if (base < 0) base += n;
#pragma unroll
    for (int i = 0; i < SR; ++i) {
        const int e = i * SNT + t;
        int64_t j = base + e;
        if (j >= n) { j -= n; if (j >= n) j %= n; }
        cp_async16(&xb[Geo<SR>::pad(e)], &src[j]);
    }
    }
    // the part of the entry the KB steps of this pass use (kb and the tables' heads)
    for (int i = t; i < (int)(sizeof(Entry) / 16); i += SNT)
        cp_async16(reinterpret_cast<char *>(&sh.tab[slot]) + 16 * i,
                   reinterpret_cast<const char *>(tab + local) + 16 * i);
    cp_async_commit();
}

// HORNER (ALIGNED only): the nested form of the step, 24 instead of 30 FP64 instructions per
// site-update; the ring is then rotated so that the marked vertex is logical site 3, i.e. slot
// 3 of a thread whose segment holds the whole window w-3..w+3 of the oracle corrections.
// resident blocks per SM.  (The nested form needs two instead of three complex vectors in
// registers and would fit five blocks of the 64 x 16 geometry at 168 registers: measured
// 3.64e11 vs 4.01e11 site-updates/s at KB = 4, so it stays at four.)
template <int SR, bool HORNER>
constexpr int stream_blocks_per_sm() { return Geo<SR>::BLOCKS_PER_SM; }

template <int KB, int SR, bool ALIGNED, bool HORNER>
__global__ void __launch_bounds__(Geo<SR>::NT, stream_blocks_per_sm<SR, HORNER>())
rk4_cycle_stream(const QwParams prm, const double2 *__restrict__ in, double2 *__restrict__ out,
                 const typename StreamShared<SR, HORNER>::Entry *__restrict__ tab, uint32_t tiles,
                 uint32_t total)
{
    static_assert(!HORNER || (ALIGNED && SR >= 2 * HW0 + 1), "nested form needs the aligned layout");
    using Shared = StreamShared<SR, HORNER>;
    using Entry = typename Shared::Entry;
    using Regs = typename std::conditional<HORNER, HState<SR>, State<SR>>::type;
    constexpr int HALO = Win<KB, SR, ALIGNED>::HALO;
    constexpr int TILE = Win<KB, SR, ALIGNED>::TILE;
    constexpr int SNT = Geo<SR>::NT;
    constexpr int SDEPTH = Geo<SR>::DEPTH;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Shared &sh = *reinterpret_cast<Shared *>(smem_raw);

    const int t = threadIdx.x;
    const int64_t mirror_m = HORNER ? prm.mirror_m : 0;
    const bool even = (prm.n & 1) == 0;
    const int64_t n = mirror_m ? mirror_m + 3 : prm.n;       // stored sites per point
    const int tl = t == 0 ? 0 : t - 1, tr = t == SNT - 1 ? SNT - 1 : t + 1;   // open ends
    const double2 zero2 = make_double2(0.0, 0.0);
    double2 *fst = sh.mbox, *lst = sh.mbox + 2 * SNT;
    const uint32_t stride = gridDim.x;

    // ring of SDEPTH shared buffers: windows i+1 .. i+SDEPTH-1 are in flight while window i
    // is being integrated.  One cp.async group per ring slot (empty past the end) keeps
    // wait_group's count uniform.
    uint32_t idx = blockIdx.x;
#pragma unroll
    for (int k = 0; k < SDEPTH - 1; ++k) {
        const uint64_t w = (uint64_t)idx + (uint64_t)k * stride;
        if (w < total) prefetch_window<KB, SR, ALIGNED, HORNER>(sh, k, in, tab, n, (uint32_t)w, tiles, t, mirror_m, even);
        else cp_async_commit();
    }

    int slot = 0;
    for (; idx < total; idx += stride, slot = (slot + 1 == SDEPTH) ? 0 : slot + 1) {
        const uint32_t local = idx / tiles, tile = idx - local * tiles;
        const int64_t first = (int64_t)tile * TILE - HALO;     // window start (may be < 0)

        cp_async_wait_pending<SDEPTH - 2>();
        __syncthreads();               // window idx has landed; the slot freed last round is free
        const Entry &pe = sh.tab[slot];
        const int kb = pe.kb;
        Regs s;
#pragma unroll
        for (int r = 0; r < SR; ++r) {
            const double2 v = sh.xbuf[slot][Geo<SR>::pad(t * SR + r)];
            s.yr[r] = v.x; s.yi[r] = v.y;
            if constexpr (HORNER) { s.zr[r] = 0.0; s.zi[r] = 0.0; }
            else { s.ur[r] = 0.0; s.ui[r] = 0.0; s.ar[r] = 0.0; s.ai[r] = 0.0; }
        }
        {
            const uint64_t w = (uint64_t)idx + (uint64_t)(SDEPTH - 1) * stride;
            const int ps = (slot + SDEPTH - 1) % SDEPTH;
            if (w < total) prefetch_window<KB, SR, ALIGNED, HORNER>(sh, ps, in, tab, n, (uint32_t)w, tiles, t, mirror_m, even);
            else cp_async_commit();
        }

        if (kb > 0) {
            // slot of the marked vertex in this thread's segment, or -1 (windows repeat mod n)
            // (the marked vertex is logical site target - rot: 0 in the ALIGNED layout)
            int64_t wrel = (prm.target - prm.stream_rot) - first - (int64_t)t * SR;
            if (mirror_m) {                      // the marked vertex is stored index 3, once
                wrel = HW0 - first - (int64_t)t * SR;
                if (wrel < 0) wrel = SR;
            } else {
                if (wrel < 0) { wrel += n; if (wrel < 0) wrel = wrel % n + n; }
                if (wrel >= n) wrel -= n;
            }
            const int own_slot = wrel < SR ? (int)wrel : -1;     // ALIGNED: 0 or -1
            const bool warp_owns = __ballot_sync(0xffffffffu, own_slot >= 0) != 0u;
            int par = 0;
            fst[t] = make_double2(s.yr[0], s.yi[0]);
            lst[t] = make_double2(s.yr[SR - 1], s.yi[SR - 1]);
            __syncthreads();                   // edge values published
            if constexpr (HORNER) {
              const bool own = own_slot == HW0;
              double2 b0 = zero2, b1 = zero2, b2 = zero2, b3 = zero2;
              for (int sidx = 0; sidx < kb; ++sidx) {
#define QW_STREAM_LEVEL(LEVEL, CBASE, NTERMS)                                           \
                {                                                                       \
                    if (LEVEL < 4 || sidx > 0) __syncthreads();                         \
                    const double2 hl = lst[par * SNT + tl];                             \
                    const double2 hr = fst[par * SNT + tr];                             \
                    const double rho = pe.rho[4 * sidx + LEVEL - 1];                    \
                    double2 o0, oL;                                                     \
                    if (LEVEL == 4 && warp_owns && own) {                               \
                        b0 = make_double2(s.yr[HW0], s.yi[HW0]);                        \
                        b1 = make_double2(s.yr[HW0 - 1] + s.yr[HW0 + 1], s.yi[HW0 - 1] + s.yi[HW0 + 1]); \
                        b2 = make_double2(s.yr[HW0 - 2] + s.yr[HW0 + 2], s.yi[HW0 - 2] + s.yi[HW0 + 2]); \
                        b3 = make_double2(s.yr[HW0 - 3] + s.yr[HW0 + 3], s.yi[HW0 - 3] + s.yi[HW0 + 3]); \
                    }                                                                   \
                    hedges<SR, LEVEL>(s, rho, hl, hr, o0, oL);                          \
                    hinterior<SR, LEVEL>(s, rho, o0, oL);                               \
                    if (warp_owns && own) {        /* oracle correction sigma_LEVEL */  \
                        const double2 *c = pe.c + 10 * sidx + CBASE;                    \
                        double2 sg = cmul2(c[0], b0);                                   \
                        if (NTERMS > 1) sg = cmul2_acc(c[1], b1, sg);                   \
                        if (NTERMS > 2) sg = cmul2_acc(c[2], b2, sg);                   \
                        if (NTERMS > 3) sg = cmul2_acc(c[3], b3, sg);                   \
                        if (LEVEL > 1) { s.zr[HW0] += sg.x; s.zi[HW0] += sg.y; }        \
                        else { s.yr[HW0] += sg.x; s.yi[HW0] += sg.y; }                  \
                    }                                                                   \
                    par ^= 1;                                                           \
                    if (LEVEL > 1) {                                                    \
                        fst[par * SNT + t] = make_double2(s.zr[0], s.zi[0]);            \
                        lst[par * SNT + t] = make_double2(s.zr[SR - 1], s.zi[SR - 1]);  \
                    } else {                                                            \
                        fst[par * SNT + t] = make_double2(s.yr[0], s.yi[0]);            \
                        lst[par * SNT + t] = make_double2(s.yr[SR - 1], s.yi[SR - 1]);  \
                    }                                                                   \
                }
                QW_STREAM_LEVEL(4, 0, 1)
                QW_STREAM_LEVEL(3, 1, 2)
                QW_STREAM_LEVEL(2, 3, 3)
                QW_STREAM_LEVEL(1, 6, 4)
#undef QW_STREAM_LEVEL
              }
            } else {
            for (int sidx = 0; sidx < kb; ++sidx) {
#define QW_STREAM_STAGE(K)                                                              \
                {                                                                       \
                    if (K > 0 || sidx > 0) __syncthreads();                             \
                    const double2 hl = lst[par * SNT + tl];                             \
                    const double2 hr = fst[par * SNT + tr];                             \
                    const double2 ab = pe.ab[4 * sidx + K];                             \
                    double xr = 0.0, xi = 0.0;                                          \
                    if (ALIGNED) {                                                      \
                        xr = (K == 0) ? s.yr[0] : s.ur[0];                              \
                        xi = (K == 0) ? s.yi[0] : s.ui[0];                              \
                    } else if (warp_owns) read_slot<SR, K>(s, own_slot, xr, xi);        \
                    rk_stage<SR, K, true>(s, ab.x, ab.y, hl, hr, zero2, 0.0, false, zero2); \
                    if (warp_owns) {                                                    \
                        if (ALIGNED) {                                                  \
                            if (own_slot == 0) fix_slot0<SR, K>(s, pe.d[4 * sidx + K], xr, xi); \
                        } else fix_slot<SR, K>(s, own_slot, pe.d[4 * sidx + K], xr, xi); \
                    }                                                                   \
                    par ^= 1;                                                           \
                    if (K < 3) {                                                        \
                        fst[par * SNT + t] = make_double2(s.ur[0], s.ui[0]);            \
                        lst[par * SNT + t] = make_double2(s.ur[SR - 1], s.ui[SR - 1]);  \
                    } else {                                                            \
                        fst[par * SNT + t] = make_double2(s.yr[0], s.yi[0]);            \
                        lst[par * SNT + t] = make_double2(s.yr[SR - 1], s.yi[SR - 1]);  \
                    }                                                                   \
                }
                QW_STREAM_STAGE(0)
                QW_STREAM_STAGE(1)
                QW_STREAM_STAGE(2)
                QW_STREAM_STAGE(3)
#undef QW_STREAM_STAGE
            }
            }
        }

        // registers -> padded shared memory -> coalesced store of the central TILE sites
        // (xbuf[slot] was consumed into registers by every thread before the barriers above,
        // or, for kb == 0, is only rewritten with the values just read from it)
        if (kb == 0) __syncthreads();
#pragma unroll
        for (int r = 0; r < SR; ++r)
            sh.xbuf[slot][Geo<SR>::pad(t * SR + r)] = make_double2(s.yr[r], s.yi[r]);
        __syncthreads();
        double2 *dst = out + (int64_t)local * n;
#pragma unroll
        for (int i = 0; i < SR; ++i) {
            const int e = i * SNT + t;
            const int64_t j = first + e;
            if (e >= HALO && e < HALO + TILE && j < n) dst[j] = sh.xbuf[slot][Geo<SR>::pad(e)];
        }
    }
    cp_async_wait_pending<0>();
}

// ---------------------------------------------------------------------------------------------
// Blackwell data movement for the same pass: 1-D bulk copies (cp.async.bulk, the TMA engine;
// SASS UBLKCP) with mbarrier transaction counts instead of per-lane cp.async / LDS / STG.
//
// The TMA engine is request-rate bound (a first version with one 128 / 256 byte copy per thread
// segment ran at 4.0 TB/s where the cp.async kernel reaches 5.7), so a window moves in CHUNKS of
// eight thread segments (2 KB with 16 sites per thread, 1 KB with 8): NCH = NT / 8 copies per
// window and direction, issued by the first NCH lanes of warp 0, one chunk each.
//   layout  chunk c lies at c * (8 SR + 1) sites of the slot: one pad element between chunks.
//           Thread t works on segment seg = (t % NCH) * 8 + t / NCH, so the eight lanes of a
//           quarter warp read from eight different chunks and their 16-byte loads fall into
//           eight different bank groups ((chunk + r) mod 8): conflict-free without a pad per
//           segment, which a bulk copy could not skip.  The halo mailbox is indexed by segment
//           (padded the same way).
//   load    lane c arrives on the slot's mbarrier with expect_tx(chunk bytes) and issues the
//           copy HBM -> slot (two copies if the chunk straddles the end of the ring); the
//           barrier completes when all NCH lanes have armed it and all bytes have landed.
//   store   after the KB steps every thread writes its sites back into its place of the same
//           slot (generic proxy) and fences to the async proxy; one block barrier; lane c
//           issues the copy slot -> HBM of the central part of its chunk.
//   reuse   at the start of the next window, once the block holds that window in registers,
//           lane c waits until its store has been READ out of shared memory
//           (cp.async.bulk.wait_group.read) and re-arms the slot with its chunk of the window
//           DEPTH - 1 ahead.
// Per window and thread this leaves SR shared loads, SR shared stores and one barrier outside
// the arithmetic; the copies of the neighbouring windows run on the TMA engine underneath it.
// The point's pass entry travels the same way (lane 0, one more bulk copy) into a ring one
// deeper than the window ring, because the slowest thread may still read the entry of the
// previous window when lane 0 re-arms a slot.
// Nested form, aligned layout, full ring, n >= one window (the other cases keep cp.async).
template <int SR>
struct TmaGeo {
    static constexpr int DEPTH = SR == 16 ? 2 : 3;
    static constexpr int NT = Geo<SR>::NT;
    static constexpr int NCH = NT / 8;                 // chunks per window
    static constexpr int CHS = 8 * SR;                 // sites per chunk
    static constexpr int SLOT = NCH * (CHS + 1);       // double2 per slot
    static constexpr int MB = NT + NT / 8;             // mailbox entries per array
    static __device__ __forceinline__ int seg(int t) { return (t % NCH) * 8 + t / NCH; }
    static __device__ __forceinline__ int place(int sg) { return (sg / 8) * (CHS + 1) + (sg % 8) * SR; }
    static __device__ __forceinline__ int mb(int sg) { return sg + sg / 8; }
};

template <int SR>
struct TmaShared {
    double2 xbuf[TmaGeo<SR>::DEPTH][TmaGeo<SR>::SLOT];
    HPassEntry tab[TmaGeo<SR>::DEPTH + 1];
    double2 mbox[4 * TmaGeo<SR>::MB];      // [parity][first/last][segment, padded]
    unsigned long long full[TmaGeo<SR>::DEPTH];
};

__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(unsigned long long *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "QW_MBAR_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra QW_MBAR_DONE;\n"
        "bra QW_MBAR_WAIT;\n"
        "QW_MBAR_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_load(void *smem_dst, const void *gmem_src, uint32_t bytes,
                                          unsigned long long *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_store(void *gmem_dst, const void *smem_src, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst),
                 "r"(smem_u32(smem_src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// lane c (< NCH) of warp 0: arm the slot of local iteration kk with chunk c of window w
template <int KB, int SR>
__device__ __forceinline__ void tma_issue_chunk(TmaShared<SR> &sh, const uint32_t kk,
                                                const double2 *in, const HPassEntry *tab,
                                                const int64_t n, const uint32_t w,
                                                const uint32_t tiles, const int c)
{
    using G = TmaGeo<SR>;
    constexpr int HALO = Win<KB, SR, true>::HALO, TILE = Win<KB, SR, true>::TILE;
    constexpr uint32_t CHB = G::CHS * sizeof(double2);
    const uint32_t local = w / tiles, tile = w - local * tiles;
    int64_t j = (int64_t)tile * TILE - HALO + (int64_t)c * G::CHS;   // first site of the chunk
    if (j < 0) j += n;                                               // n >= W: one wrap at most
    if (j >= n) j -= n;
    const uint32_t xs = kk % G::DEPTH, ts = kk % (G::DEPTH + 1);
    unsigned long long *bar = &sh.full[xs];
    double2 *dst = &sh.xbuf[xs][c * (G::CHS + 1)];
    const double2 *src = in + (int64_t)local * n;
    mbar_arrive_expect_tx(bar, c == 0 ? CHB + (uint32_t)sizeof(HPassEntry) : CHB);
    const int64_t head = n - j;                                      // sites before the ring ends
    if (head >= G::CHS) {
        bulk_load(dst, src + j, CHB, bar);
    } else {                                                         // multiples of SR sites
        bulk_load(dst, src + j, (uint32_t)head * (uint32_t)sizeof(double2), bar);
        bulk_load(dst + head, src, CHB - (uint32_t)head * (uint32_t)sizeof(double2), bar);
    }
    if (c == 0) bulk_load(&sh.tab[ts], tab + local, (uint32_t)sizeof(HPassEntry), bar);
}

// kb steps of the nested form on a thread's segment with open ends.  OWN: this WARP holds the
// marked vertex of the window (one warp in about 2 x 69 at N = 65536); only its code contains the
// oracle corrections (thread-level, `own` selects the thread), so the common path is free of
// divergence handling -- with the branches in every level (first version) the step loop spent
// 7 % of its samples on VOTEU / BSSY / BSYNC and could not be scheduled across them.
template <int SR, bool OWN>
__device__ __forceinline__ void tma_steps(HState<SR> &s, const HPassEntry &pe, const int kb,
                                          double2 *fst, double2 *lst, const int MB, const int mt,
                                          const int ml, const int mr, const bool own)
{
    const double2 zero2 = make_double2(0.0, 0.0);
    double2 b0 = zero2, b1 = zero2, b2 = zero2, b3 = zero2;
    for (int sidx = 0; sidx < kb; ++sidx) {
#define QW_TMA_LEVEL(LEVEL, CBASE, NTERMS)                                              \
        {                                                                               \
            constexpr int par = (4 - LEVEL) & 1;                                        \
            if (LEVEL < 4 || sidx > 0) __syncthreads();                                 \
            const double2 hl = lst[par * MB + ml];                                      \
            const double2 hr = fst[par * MB + mr];                                      \
            const double rho = pe.rho[4 * sidx + LEVEL - 1];                            \
            double2 o0, oL;                                                             \
            if (OWN && LEVEL == 4 && own) {                                             \
                b0 = make_double2(s.yr[HW0], s.yi[HW0]);                                \
                b1 = make_double2(s.yr[HW0 - 1] + s.yr[HW0 + 1], s.yi[HW0 - 1] + s.yi[HW0 + 1]); \
                b2 = make_double2(s.yr[HW0 - 2] + s.yr[HW0 + 2], s.yi[HW0 - 2] + s.yi[HW0 + 2]); \
                b3 = make_double2(s.yr[HW0 - 3] + s.yr[HW0 + 3], s.yi[HW0 - 3] + s.yi[HW0 + 3]); \
            }                                                                           \
            hedges<SR, LEVEL>(s, rho, hl, hr, o0, oL);                                  \
            hinterior<SR, LEVEL>(s, rho, o0, oL);                                       \
            if (OWN && own) {                  /* oracle correction sigma_LEVEL */      \
                const double2 *c = pe.c + 10 * sidx + CBASE;                            \
                double2 sgm = cmul2(c[0], b0);                                          \
                if (NTERMS > 1) sgm = cmul2_acc(c[1], b1, sgm);                         \
                if (NTERMS > 2) sgm = cmul2_acc(c[2], b2, sgm);                         \
                if (NTERMS > 3) sgm = cmul2_acc(c[3], b3, sgm);                         \
                if (LEVEL > 1) { s.zr[HW0] += sgm.x; s.zi[HW0] += sgm.y; }              \
                else { s.yr[HW0] += sgm.x; s.yi[HW0] += sgm.y; }                        \
            }                                                                           \
            if (LEVEL > 1) {                                                            \
                fst[(par ^ 1) * MB + mt] = make_double2(s.zr[0], s.zi[0]);              \
                lst[(par ^ 1) * MB + mt] = make_double2(s.zr[SR - 1], s.zi[SR - 1]);    \
            } else {                                                                    \
                fst[(par ^ 1) * MB + mt] = make_double2(s.yr[0], s.yi[0]);              \
                lst[(par ^ 1) * MB + mt] = make_double2(s.yr[SR - 1], s.yi[SR - 1]);    \
            }                                                                           \
        }
        QW_TMA_LEVEL(4, 0, 1)
        QW_TMA_LEVEL(3, 1, 2)
        QW_TMA_LEVEL(2, 3, 3)
        QW_TMA_LEVEL(1, 6, 4)
#undef QW_TMA_LEVEL
    }
}

template <int KB, int SR>
__global__ void __launch_bounds__(Geo<SR>::NT, Geo<SR>::BLOCKS_PER_SM)
rk4_cycle_stream_tma(const QwParams prm, const double2 *__restrict__ in, double2 *__restrict__ out,
                     const HPassEntry *__restrict__ tab, uint32_t tiles, uint32_t total)
{
    using G = TmaGeo<SR>;
    static_assert(SR >= 2 * HW0 + 1 && sizeof(HPassEntry) % 16 == 0 && KB <= MAXKB, "layout");
    constexpr int HALO = Win<KB, SR, true>::HALO;
    constexpr int TILE = Win<KB, SR, true>::TILE;
    constexpr int SNT = Geo<SR>::NT;
    constexpr int DEPTH = G::DEPTH;
    constexpr int MB = G::MB;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    TmaShared<SR> &sh = *reinterpret_cast<TmaShared<SR> *>(smem_raw);

    const int t = threadIdx.x;
    const int sg = G::seg(t);                          // the thread's segment of every window
    const int64_t n = prm.n;
    const int ml = G::mb(sg == 0 ? 0 : sg - 1), mr = G::mb(sg == SNT - 1 ? SNT - 1 : sg + 1);
    const int mt = G::mb(sg);                          // open ends: the edge segments see themselves
    double2 *fst = sh.mbox, *lst = sh.mbox + 2 * MB;
    const uint32_t stride = gridDim.x;
    const bool issuer = t < G::NCH;                    // lane c moves chunk c
    // window idx = local * tiles + tile, advanced without divisions
    const uint32_t dl = stride / tiles, dt = stride - dl * tiles;
    uint32_t local = blockIdx.x / tiles, tile = blockIdx.x - local * tiles;

    if (t == 0) {
#pragma unroll
        for (int k = 0; k < DEPTH; ++k) mbar_init(&sh.full[k], G::NCH);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (issuer) {
#pragma unroll
        for (int k = 0; k < DEPTH - 1; ++k) {
            const uint64_t w = (uint64_t)blockIdx.x + (uint64_t)k * stride;
            if (w < total) tma_issue_chunk<KB, SR>(sh, (uint32_t)k, in, tab, n, (uint32_t)w, tiles, t);
        }
    }

    uint32_t kk = 0;
    for (uint32_t idx = blockIdx.x; idx < total; idx += stride, ++kk) {
        const int64_t first = (int64_t)tile * TILE - HALO;     // window start (may be < 0)
        const uint32_t xs = kk % DEPTH, ts = kk % (DEPTH + 1);
        double2 *seg = &sh.xbuf[xs][G::place(sg)];

        mbar_wait(&sh.full[xs], (kk / DEPTH) & 1u);            // window idx and its entry landed
        HState<SR> s;
#pragma unroll
        for (int r = 0; r < SR; ++r) {
            const double2 v = seg[r];
            s.yr[r] = v.x; s.yi[r] = v.y;
            s.zr[r] = 0.0; s.zi[r] = 0.0;
        }
        const HPassEntry &pe = sh.tab[ts];
        const int kb = pe.kb;
        fst[mt] = make_double2(s.yr[0], s.yi[0]);
        lst[mt] = make_double2(s.yr[SR - 1], s.yi[SR - 1]);
        // the slot of the previous window: its stores have left shared memory -> re-arm it
        if (issuer) {
            const uint64_t w = (uint64_t)idx + (uint64_t)(DEPTH - 1) * stride;
            if (kk > 0) bulk_wait_read0();
            if (w < total)
                tma_issue_chunk<KB, SR>(sh, kk + DEPTH - 1, in, tab, n, (uint32_t)w, tiles, t);
        }
        __syncthreads();                       // edge values published

        if (kb > 0) {
            // the marked vertex is logical site HW0 of the rotated ring: slot HW0 of the thread
            // whose segment starts at a multiple of n
            int64_t wrel = HW0 - first - (int64_t)sg * SR;
            if (wrel < 0) wrel += n;
            if (wrel >= n) wrel -= n;
            const bool own = wrel == HW0;
            if (__ballot_sync(0xffffffffu, own) != 0u)
                tma_steps<SR, true>(s, pe, kb, fst, lst, MB, mt, ml, mr, own);
            else
                tma_steps<SR, false>(s, pe, kb, fst, lst, MB, mt, ml, mr, false);
        }

        // registers -> the thread's place in the slot; then lane c: central part of chunk c -> HBM
#pragma unroll
        for (int r = 0; r < SR; ++r) seg[r] = make_double2(s.yr[r], s.yi[r]);
        fence_proxy_async();
        __syncthreads();
        if (issuer) {
            int e0 = t * G::CHS, e1 = e0 + G::CHS;                 // window positions of the chunk
            if (e0 < HALO) e0 = HALO;
            if (e1 > HALO + TILE) e1 = HALO + TILE;
            if (first + e1 > n) e1 = (int)(n - first);             // last tile: the ring ends
            if (e1 > e0)
                bulk_store(out + (int64_t)local * n + first + e0,
                           &sh.xbuf[xs][t * (G::CHS + 1) + (e0 - t * G::CHS)],
                           (uint32_t)(e1 - e0) * (uint32_t)sizeof(double2));
            bulk_commit();
        }
        local += dl;
        tile += dt;
        if (tile >= tiles) { tile -= tiles; ++local; }
    }
    if (issuer) bulk_wait0();
}

__global__ void stream_init(double2 *buf, int64_t total, double amp)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride)
        buf[i] = make_double2(amp, 0.0);
}

// largest step count among the points of a chunk (needed only when it varies per point)
__global__ void stream_max_steps(const QwParams prm, int64_t count, unsigned long long *out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const QwPoint pt = qw_point(prm, prm.point_offset + i);
    atomicMax(out, (unsigned long long)pt.steps);
}

// p, ||psi||^2 and optional psi(T) copy of every point of the chunk
__global__ void __launch_bounds__(512) stream_finish(const QwParams prm, const double2 *__restrict__ buf)
{
    __shared__ double red[64];
    const int64_t local = blockIdx.x, n = prm.n;
    const QwPoint pt = qw_point(prm, prm.point_offset + local);
    if (prm.mirror_m) {
        // half ring behind 3 ghosts: every chain site counts twice in the norm except the
        // marked vertex and (even n) the antipode; psi is written with both mirror images
        const int64_t m = prm.mirror_m;
        const bool even = (n & 1) == 0;
        const double2 *src = buf + local * (m + 3) + 3;
        double nrm = 0.0;
        for (int64_t c = threadIdx.x; c < m; c += blockDim.x) {
            const double2 v = src[c];
            const double wgt = (c == 0 || (even && c == m - 1)) ? 1.0 : 2.0;
            nrm = fma(wgt * v.x, v.x, fma(wgt * v.y, v.y, nrm));
            if (prm.psi) {
                int64_t a = prm.target + c, b = prm.target - c;
                if (a >= n) a -= n;
                if (b < 0) b += n;
                prm.psi[(size_t)pt.q * (size_t)n + a] = v;
                prm.psi[(size_t)pt.q * (size_t)n + b] = v;
            }
        }
        const double2 tot = qw_block_sum2(nrm, 0.0, red);
        if (threadIdx.x == 0) {
            const double2 v = src[0];
            prm.p[pt.q] = fma(v.x, v.x, v.y * v.y) / tot.x;
            if (prm.norm) prm.norm[pt.q] = tot.x;
        }
        return;
    }
    const double2 *src = buf + local * n;
    double nrm = 0.0;
    for (int64_t j = threadIdx.x; j < n; j += blockDim.x) {
        const double2 v = src[j];
        nrm = fma(v.x, v.x, fma(v.y, v.y, nrm));
        if (prm.psi) {                          // undo the rotation: physical (l + rot) mod n
            int64_t jp = j + prm.stream_rot;
            if (jp >= n) jp -= n;
            prm.psi[(size_t)pt.q * (size_t)n + jp] = v;
        }
    }
    const double2 tot = qw_block_sum2(nrm, 0.0, red);
    if (threadIdx.x == 0) {
        int64_t lw = prm.target - prm.stream_rot;          // logical site of the marked vertex
        if (lw < 0) lw += n;
        const double2 v = src[lw];
        prm.p[pt.q] = fma(v.x, v.x, v.y * v.y) / tot.x;
        if (prm.norm) prm.norm[pt.q] = tot.x;
    }
}

// Temporal blocking (KB >= 2) is FP64-bound: 64 x 16 blocks.  KB = 1 is HBM-bound and wants
// the deeper prefetch ring of the 128 x 8 geometry; so do QW_FLAG_R8 and short rings.
int stream_sr(const QwParams &prm)
{
    const bool kb1 = (prm.flags & QW_FLAG_STREAM_KB1) && !(prm.flags & QW_FLAG_STREAM_KB2);
    return ((prm.flags & QW_FLAG_R8) || kb1 || prm.n < 4096) ? 8 : 16;
}

// RK4 steps fused per pass.  The load / unload phases of a window cost about as much as 1.75
// steps (measured, 64 x 16 geometry), so deeper temporal blocking keeps paying while the halo
// (4 sites per step and side, rounded to the thread segment) stays a few per cent of the
// window: 8 steps (halo 32 of 1024) by default, 4 in the 128 x 8 geometry.
int stream_kb(const QwParams &prm)
{
    if ((prm.flags & QW_FLAG_STREAM_KB1) && (prm.flags & QW_FLAG_STREAM_KB2)) return 4;
    if (prm.flags & QW_FLAG_STREAM_KB1) return 1;
    if (prm.flags & QW_FLAG_STREAM_KB2) return 2;
    return stream_sr(prm) == 16 ? 8 : 4;
}

int g_sm_count = 0;

int sm_count()
{
    if (g_sm_count == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev);
        if (g_sm_count <= 0) g_sm_count = 148;
    }
    return g_sm_count;
}

template <int KB, int SR, bool ALIGNED, bool HORNER>
cudaError_t configure()
{
    static QwOncePerDevice configured;
    if (!configured.needed()) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(rk4_cycle_stream<KB, SR, ALIGNED, HORNER>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)sizeof(StreamShared<SR, HORNER>));
    if (e == cudaSuccess) configured.mark();
    return e;
}

template <int KB, int SR>
cudaError_t configure_tma()
{
    static QwOncePerDevice configured;
    if (!configured.needed()) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(rk4_cycle_stream_tma<KB, SR>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)sizeof(TmaShared<SR>));
    if (e == cudaSuccess) configured.mark();
    return e;
}

// bulk-copy (TMA) data movement: nested form on the aligned full ring; QW_FLAG_CYCLE_LEGACY
// keeps the cp.async kernel for A/B runs
bool stream_tma(const QwParams &prm, int sr)
{
    return !(prm.flags & (QW_FLAG_NO_HORNER | QW_FLAG_CYCLE_LEGACY)) && !prm.mirror_m &&
           prm.n % sr == 0 && prm.n >= 1024 && !((prm.flags & QW_FLAG_MIRROR) && prm.n >= 8192);
}

template <int KB, int SR>
cudaError_t launch_pass_tma(const QwParams &prm, const double2 *in, double2 *out, void *tab,
                            int64_t step0, int64_t points, cudaStream_t st)
{
    constexpr int TILE = Win<KB, SR, true>::TILE;
    HPassEntry *etab = static_cast<HPassEntry *>(tab);
    stream_pass_tables_horner<<<(unsigned)((points + 127) / 128), 128, 0, st>>>(prm, points, step0,
                                                                               KB, etab);
    const uint32_t tiles = (uint32_t)((prm.n + TILE - 1) / TILE);
    const uint64_t total = (uint64_t)points * tiles;
    uint32_t grid = (uint32_t)(sm_count() * Geo<SR>::BLOCKS_PER_SM);      // persistent
    if (total < grid) grid = (uint32_t)total;
    cudaError_t e = configure_tma<KB, SR>();
    if (e != cudaSuccess) return e;
    rk4_cycle_stream_tma<KB, SR><<<grid, Geo<SR>::NT, sizeof(TmaShared<SR>), st>>>(
        prm, in, out, etab, tiles, (uint32_t)total);
    return cudaGetLastError();
}

template <int KB, int SR, bool ALIGNED, bool HORNER>
cudaError_t launch_pass_a(const QwParams &prm, const double2 *in, double2 *out, void *tab,
                          int64_t step0, int64_t points, cudaStream_t st)
{
    using Entry = typename StreamShared<SR, HORNER>::Entry;
    constexpr int TILE = Win<KB, SR, ALIGNED>::TILE;
    Entry *etab = static_cast<Entry *>(tab);
    if constexpr (HORNER)
        stream_pass_tables_horner<<<(unsigned)((points + 127) / 128), 128, 0, st>>>(
            prm, points, step0, KB, etab);
    else
        stream_pass_tables<<<(unsigned)((points + 127) / 128), 128, 0, st>>>(prm, points, step0,
                                                                             KB, etab);
    const int64_t len = prm.mirror_m ? prm.mirror_m + 3 : prm.n;      // stored sites per point
    const uint32_t tiles = (uint32_t)((len + TILE - 1) / TILE);
    const uint64_t total = (uint64_t)points * tiles;
    uint32_t grid = (uint32_t)(sm_count() * stream_blocks_per_sm<SR, HORNER>());   // persistent
    if (total < grid) grid = (uint32_t)total;
    cudaError_t e = configure<KB, SR, ALIGNED, HORNER>();
    if (e != cudaSuccess) return e;
    rk4_cycle_stream<KB, SR, ALIGNED, HORNER>
        <<<grid, Geo<SR>::NT, sizeof(StreamShared<SR, HORNER>), st>>>(prm, in, out, etab, tiles,
                                                                      (uint32_t)total);
    return cudaGetLastError();
}

// nested form wherever the aligned layout applies (n a multiple of the sites per thread)
bool stream_mirror(const QwParams &prm)
{
    // one reflection per window end must suffice: the half ring is longer than a window
    return (prm.flags & QW_FLAG_MIRROR) && !(prm.flags & QW_FLAG_NO_HORNER) && prm.n >= 8192;
}

bool stream_horner(const QwParams &prm, int sr)
{
    if (stream_mirror(prm)) return true;         // no wrap: every window start is aligned
    return prm.n % sr == 0 && !(prm.flags & QW_FLAG_NO_HORNER);
}

template <int KB, int SR>
cudaError_t launch_pass(const QwParams &prm, const double2 *in, double2 *out, void *tab,
                        int64_t step0, int64_t points, cudaStream_t st)
{
    if (stream_tma(prm, SR)) return launch_pass_tma<KB, SR>(prm, in, out, tab, step0, points, st);
    if (stream_horner(prm, SR))
        return launch_pass_a<KB, SR, true, true>(prm, in, out, tab, step0, points, st);
    return prm.n % SR == 0
               ? launch_pass_a<KB, SR, true, false>(prm, in, out, tab, step0, points, st)
               : launch_pass_a<KB, SR, false, false>(prm, in, out, tab, step0, points, st);
}

template <int KB, int SR>
cudaError_t plan_of(qw_plan *plan, bool horner, int tma)
{
    cudaFuncAttributes fa;
    int occ = 0;
    cudaError_t e;
    size_t smem;
    if (tma == 1) {
        smem = sizeof(TmaShared<SR>);
        e = configure_tma<KB, SR>();
        if (e == cudaSuccess) e = cudaFuncGetAttributes(&fa, rk4_cycle_stream_tma<KB, SR>);
        if (e == cudaSuccess)
            e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(
                &occ, rk4_cycle_stream_tma<KB, SR>, Geo<SR>::NT, smem);
    } else if (horner) {
        smem = sizeof(StreamShared<SR, true>);
        e = configure<KB, SR, true, true>();
        if (e == cudaSuccess) e = cudaFuncGetAttributes(&fa, rk4_cycle_stream<KB, SR, true, true>);
        if (e == cudaSuccess)
            e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(
                &occ, rk4_cycle_stream<KB, SR, true, true>, Geo<SR>::NT, smem);
    } else {
        smem = sizeof(StreamShared<SR, false>);
        e = configure<KB, SR, true, false>();
        if (e == cudaSuccess) e = cudaFuncGetAttributes(&fa, rk4_cycle_stream<KB, SR, true, false>);
        if (e == cudaSuccess)
            e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(
                &occ, rk4_cycle_stream<KB, SR, true, false>, Geo<SR>::NT, smem);
    }
    if (e != cudaSuccess) return e;
    plan->sites_per_thread = SR;
    plan->threads_per_block = Geo<SR>::NT;
    plan->blocks_per_sm = occ;
    plan->regs_per_thread = fa.numRegs;
    plan->smem_bytes = (int)(fa.sharedSizeBytes + smem);
    plan->steps_per_pass = KB;
    return cudaSuccess;
}

#define QW_STREAM_DISPATCH(FN, ...)                                                     \
    (sr == 16 ? (kb == 1 ? FN<1, 16>(__VA_ARGS__) : kb == 2 ? FN<2, 16>(__VA_ARGS__)    \
                 : kb == 8 ? FN<8, 16>(__VA_ARGS__) : FN<4, 16>(__VA_ARGS__))           \
              : (kb == 1 ? FN<1, 8>(__VA_ARGS__) : kb == 2 ? FN<2, 8>(__VA_ARGS__)      \
                                                           : FN<4, 8>(__VA_ARGS__)))

}  // namespace

bool qw_stream_supported(const QwParams &prm)
{
    return prm.topology == QW_TOPO_CYCLE && prm.n >= 8;
}

cudaError_t qw_stream_plan(const QwParams &prm, qw_plan *plan)
{
    plan->path = 3;
    const int kb = stream_kb(prm), sr = stream_sr(prm);
    return QW_STREAM_DISPATCH(plan_of, plan, stream_horner(prm, sr),
                              stream_tma(prm, sr) ? 1 : 0);
}

cudaError_t qw_launch_stream(const QwParams &prm_in, cudaStream_t st, int64_t *launches)
{
    QwParams prm = prm_in;
    const int kb = stream_kb(prm), sr = stream_sr(prm);
    // QW_FLAG_MIRROR: the half ring w .. w + n/2 behind 3 mirrored ghost sites (the marked
    // vertex is stored index 3 = slot 3 of the first thread segment), reflecting window halos
    prm.mirror_m = stream_mirror(prm) ? prm.n / 2 + 1 : 0;
    const int64_t n = prm.mirror_m ? prm.mirror_m + 3 : prm.n;        // stored sites per point
    // aligned layout: the marked vertex is logical site 0 (stage form) or HW0 (nested form)
    prm.stream_rot = 0;
    if (!prm.mirror_m && n % sr == 0) {
        prm.stream_rot = prm.target - (stream_horner(prm, sr) ? HW0 : 0);
        if (prm.stream_rot < 0) prm.stream_rot += n;
    }
    // points in flight: bounded by ~8 GiB of ping-pong scratch and by 32-bit window indices
    int64_t chunk = (int64_t)(8ull << 30) / (2 * n * (int64_t)sizeof(double2));
    const int64_t tiles_max = n / (1024 - 64) + 2;
    if (chunk > (int64_t)0x7fffffff / tiles_max) chunk = (int64_t)0x7fffffff / tiles_max;
    if (chunk < 1) chunk = 1;
    if (chunk > prm.n_points) chunk = prm.n_points;
    double2 *bufA = nullptr;
    unsigned long long *dmax = nullptr;
    cudaError_t e = cudaMallocAsync((void **)&bufA, (size_t)chunk * 2 * n * sizeof(double2), st);
    if (e != cudaSuccess) return e;
    double2 *bufB = bufA + chunk * n;
    void *tab = nullptr;
    const size_t entry = sizeof(HPassEntry) > sizeof(PassEntry) ? sizeof(HPassEntry) : sizeof(PassEntry);
    e = cudaMallocAsync(&tab, (size_t)chunk * entry, st);
    if (e != cudaSuccess) { cudaFreeAsync(bufA, st); return e; }
    const bool varying = prm.step_size > 0.0 || prm.n_steps != nullptr;
    if (varying) {
        e = cudaMallocAsync((void **)&dmax, sizeof(unsigned long long), st);
        if (e != cudaSuccess) { cudaFreeAsync(bufA, st); cudaFreeAsync(tab, st); return e; }
    }
    const double amp = 1.0 / sqrt((double)prm.n);

    for (int64_t p0 = 0; p0 < prm.n_points && e == cudaSuccess; p0 += chunk) {
        const int64_t cnt = (prm.n_points - p0 < chunk) ? (prm.n_points - p0) : chunk;
        prm.point_offset = p0;
        int64_t max_steps = prm.n_steps_uniform;
        if (varying) {
            unsigned long long h = 0;
            cudaMemsetAsync(dmax, 0, sizeof(h), st);
            stream_max_steps<<<(unsigned)((cnt + 255) / 256), 256, 0, st>>>(prm, cnt, dmax);
            ++*launches;
            e = cudaMemcpyAsync(&h, dmax, sizeof(h), cudaMemcpyDeviceToHost, st);
            if (e == cudaSuccess) e = cudaStreamSynchronize(st);
            if (e != cudaSuccess) break;
            max_steps = (int64_t)h;
        }
        stream_init<<<sm_count() * 8, 256, 0, st>>>(bufA, cnt * n, amp);
        ++*launches;
        double2 *cur = bufA, *nxt = bufB;
        for (int64_t s0 = 0; s0 < max_steps && e == cudaSuccess; s0 += kb) {
            e = QW_STREAM_DISPATCH(launch_pass, prm, cur, nxt, tab, s0, cnt, st);
            *launches += 2;
            double2 *tmp = cur; cur = nxt; nxt = tmp;
        }
        if (e != cudaSuccess) break;
        stream_finish<<<(unsigned)cnt, 512, 0, st>>>(prm, cur);
        ++*launches;
        e = cudaGetLastError();
    }
    cudaFreeAsync(bufA, st);
    cudaFreeAsync(tab, st);
    if (dmax) cudaFreeAsync(dmax, st);
    return e;
}
