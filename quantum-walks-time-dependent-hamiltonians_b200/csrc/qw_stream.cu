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
    int32_t kb;
    int32_t pad[3];
    double2 c[10 * MAXKB];             // read by the one thread that holds the marked vertex
};
constexpr uint32_t HPASS_HEAD = 4 * MAXKB * sizeof(double) + 16;   // rho and kb

template <int SR, bool HORNER>
struct StreamShared {
    using Entry = typename std::conditional<HORNER, HPassEntry, PassEntry>::type;
    double2 xbuf[Geo<SR>::DEPTH][Geo<SR>::PAD];   // window ring (load target, store staging)
    Entry tab[Geo<SR>::DEPTH];            // the point's pass entry travels with its window
    double2 mbox[4 * Geo<SR>::NT];        // [parity][first/last][thread]
};

__global__ void stream_pass_tables_horner(const QwParams prm, int64_t count, int64_t step0,
                                          int kbmax, HPassEntry *__restrict__ tab,
                                          unsigned int *__restrict__ counter = nullptr)
{
    const int64_t local = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (local == 0 && counter) *counter = 0u;          // window counter of the bulk-copy kernel
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
// What shaped it (measured on the N = 65536 grid, tools/ab_stream.py, tools/probe_stream.py):
//  * one copy per thread segment (128 / 256 B) is request-rate bound in the TMA engine: 4.0 TB/s
//    where the cp.async kernel reaches 5.7;
//  * a bulk copy costs its issuing warp about 60 cycles, so with eight 2 KB chunks per window and
//    direction (a pad element between chunks for the shared-memory banks) the issue work alone
//    was 1100 cycles per window, all of it in front of the next window's arithmetic;
//  * hence the pad lives in HBM too: psi is stored with ONE PAD ELEMENT AFTER EVERY 128 SITES
//    (+0.8 % traffic).  A window -- 1024 sites plus its 8 pads -- is then contiguous in HBM and
//    in shared memory, and moves with ONE bulk load and ONE bulk store (two where it wraps
//    around the ring).
// Bank conflicts: thread t works on segment seg = (t % 8) * (NT / 8) + t / 8, so the eight lanes
// of a quarter warp read segments that lie 128 sites apart, i.e. behind 0, 1, ... 7 more pads:
// their 16-byte loads fall into eight different bank groups whatever the window's offset within
// a 128-site block.  The halo mailbox is indexed by segment and padded the same way.
//   load    one thread arrives on the slot's mbarrier with expect_tx(window + entry bytes) and
//           issues the copies; the barrier completes when all bytes have landed;
//   store   after the KB steps every thread writes its sites back into its place of the same
//           slot (generic proxy) and fences to the async proxy; one block barrier; one thread
//           issues the copy of the central part of the window to HBM;
//   ring    three slots, window kk in slot kk % 3; window kk + 2 goes into the slot window
//           kk - 1 has left, once the store of window kk - 1 has been READ out of shared memory
//           (cp.async.bulk.wait_group.read).
//           KB >= 4 (FP64-bound): thread 0 stores window kk at its end and thread 32 -- the other
//           warp, side by side -- requests window kk + 2; the store it depends on is a whole
//           window old.  KB <= 2 (HBM-bound): thread 0 requests window kk + 2 at the START of
//           window kk, as soon as the block holds window kk in registers (two windows of lead).
// Per window and thread this leaves SR shared loads, SR shared stores and two barriers outside
// the arithmetic.  The head of the point's pass entry (rho, kb: 272 B) travels with the window
// into a ring one deeper than the window ring; the correction coefficients are fetched by the
// one warp per ring that holds the marked vertex.
// Nested form, full ring, n a multiple of 128 and >= one window (other cases: cp.async kernel).
constexpr int TMA_PADQ = 128;                          // sites between two pad elements

__host__ __device__ __forceinline__ int64_t tma_padded(int64_t x)      // index of site x, x >= 0
{
    return x + x / TMA_PADQ;
}

template <int SR>
struct TmaGeo {
    static constexpr int DEPTH = 3;                    // window kk lives in slot kk % 3
    static constexpr int NT = Geo<SR>::NT;
    static constexpr int W = Geo<SR>::W;
    static constexpr int SLOT = W + W / TMA_PADQ;      // double2 per slot
    static constexpr int MB = NT + 8;                  // mailbox entries per array
    static __device__ __forceinline__ int seg(int t) { return (t % 8) * (NT / 8) + t / 8; }
    static __device__ __forceinline__ int mb(int sg) { return sg + sg / (NT / 8); }
};

struct TmaHead {                           // the first HPASS_HEAD bytes of an HPassEntry
    double rho[4 * MAXKB];
    int32_t kb;
    int32_t pad[3];
};

template <int SR>
struct TmaShared {
    double2 xbuf[TmaGeo<SR>::DEPTH][TmaGeo<SR>::SLOT];
    TmaHead tab[TmaGeo<SR>::DEPTH + 1];      // window kk's entry in tab[kk % (DEPTH + 1)]
    double2 cown[10 * MAXKB];              // correction coefficients of the window with the marked vertex
    double2 mbox[4 * TmaGeo<SR>::MB];      // [parity][first/last][segment, padded]
    unsigned long long full[TmaGeo<SR>::DEPTH];
    uint32_t widx[TmaGeo<SR>::DEPTH + 1];  // window index of iteration kk (0xffffffff: none left)
};

// smem_u32: qw_horner_level.cuh
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

__device__ __forceinline__ void mbar_arrive(unsigned long long *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ONE thread: arm the slot of local iteration kk with window w of the padded ring and the head of
// its point's pass entry; w >= total: nothing left -- complete the barrier empty-handed so that
// the block sees the end marker
template <int KB, int SR>
__device__ __forceinline__ void tma_issue_window(TmaShared<SR> &sh, const uint32_t kk,
                                                 const double2 *in, const HPassEntry *tab,
                                                 const int64_t n, const int64_t np,
                                                 const uint32_t w, const uint32_t tiles,
                                                 const uint32_t total)
{
    using G = TmaGeo<SR>;
    constexpr int HALO = Win<KB, SR, true>::HALO, TILE = Win<KB, SR, true>::TILE;
    constexpr uint32_t WB = G::SLOT * sizeof(double2);
    const uint32_t xs = kk % G::DEPTH;
    unsigned long long *bar = &sh.full[xs];
    if (w >= total) {
        sh.widx[kk % (G::DEPTH + 1)] = 0xffffffffu;
        mbar_arrive(bar);
        return;
    }
    sh.widx[kk % (G::DEPTH + 1)] = w;
    const uint32_t local = w / tiles, tile = w - local * tiles;
    double2 *dst = sh.xbuf[xs];
    const double2 *src = in + (int64_t)local * np;
    int64_t j = (int64_t)tile * TILE - HALO;                         // first site of the window
    if (j < 0) j += n;                                               // n >= W: one wrap at most
    const int64_t pj = tma_padded(j);
    mbar_arrive_expect_tx(bar, WB + HPASS_HEAD);
    bulk_load(&sh.tab[kk % (G::DEPTH + 1)], tab + local, HPASS_HEAD, bar);
    const int64_t head = np - pj;                                    // elements before the ring ends
    if (head >= G::SLOT) {
        bulk_load(dst, src + pj, WB, bar);
    } else {
        bulk_load(dst, src + pj, (uint32_t)head * (uint32_t)sizeof(double2), bar);
        bulk_load(dst + head, src, WB - (uint32_t)head * (uint32_t)sizeof(double2), bar);
    }
}

// kb steps of the nested form on a thread's segment with open ends.  OWN: this WARP holds the
// marked vertex of the window (one warp in about 2 x 69 at N = 65536); only its code contains the
// oracle corrections (thread-level, `own` selects the thread), so the common path is free of
// divergence handling -- with the branches in every level (first version) the step loop spent
// 7 % of its samples on VOTEU / BSSY / BSYNC and could not be scheduled across them.
template <int SR, bool OWN>
__device__ __forceinline__ void tma_steps(HState<SR> &s, const TmaHead &pe, const double2 *coef,
                                          const int kb, double2 *fst, double2 *lst, const int MB,
                                          const int mt, const int ml, const int mr, const bool own)
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
                const double2 *c = coef + 10 * sidx + CBASE;                            \
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

// Development probe (-DQW_PROBE): cycles thread 0 of every block spends in the phases of a window,
// summed over the launch; read with qw_debug_probe().
#ifdef QW_PROBE
__device__ unsigned long long qw_probe[8];
__device__ unsigned long long qw_probe_blocks[3 * 1024];      // smid, start ns, end ns per block
#define QW_PROBE_T(var) const long long var = clock64()
#define QW_PROBE_ADD(i, a, b) if (t == 0) atomicAdd(&qw_probe[i], (unsigned long long)((b) - (a)))
#else
#define QW_PROBE_T(var)
#define QW_PROBE_ADD(i, a, b)
#endif

template <int KB, int SR>
__global__ void __launch_bounds__(Geo<SR>::NT, Geo<SR>::BLOCKS_PER_SM)
rk4_cycle_stream_tma(const QwParams prm, const double2 *__restrict__ in, double2 *__restrict__ out,
                     const HPassEntry *__restrict__ tab, uint32_t tiles, uint32_t total,
                     unsigned int *__restrict__ next_window)
{
    using G = TmaGeo<SR>;
    static_assert(SR >= 2 * HW0 + 1 && sizeof(HPassEntry) % 16 == 0 && KB <= MAXKB, "layout");
    static_assert(TMA_PADQ % SR == 0 && Geo<SR>::W % TMA_PADQ == 0, "segments must not straddle a pad");
    constexpr int HALO = Win<KB, SR, true>::HALO;
    constexpr int TILE = Win<KB, SR, true>::TILE;
    constexpr int SNT = Geo<SR>::NT;
    constexpr int DEPTH = G::DEPTH;
    constexpr int MB = G::MB;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    TmaShared<SR> &sh = *reinterpret_cast<TmaShared<SR> *>(smem_raw);

    const int t = threadIdx.x;
    const int sg = G::seg(t);                          // the thread's segment of every window
    const int64_t n = prm.n, np = tma_padded(prm.n);   // sites / stored elements per point
    const int ml = G::mb(sg == 0 ? 0 : sg - 1), mr = G::mb(sg == SNT - 1 ? SNT - 1 : sg + 1);
    const int mt = G::mb(sg);                          // open ends: the edge segments see themselves
    double2 *fst = sh.mbox, *lst = sh.mbox + 2 * MB;

    // Windows are handed out by an atomic counter, not by a fixed stride: co-resident blocks do
    // not get equal shares of the FP64 pipe (measured: with equal window counts the blocks of one
    // launch finished between 567 and 949 us), so a static split ends with a good part of the
    // machine idle.  The loader thread draws an index one window ahead of its use, so the
    // atomic's round trip is never waited for.
    constexpr bool LATE = KB >= 4;
    const bool storer = t == 0;
    const bool loader = LATE ? t == 32 : t == 0;
#ifdef QW_PROBE
    if (t == 0 && blockIdx.x < 1024) {
        unsigned sm; unsigned long long ns;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns));
        qw_probe_blocks[3 * blockIdx.x] = sm;
        qw_probe_blocks[3 * blockIdx.x + 1] = ns;
    }
#endif
    if (t == 0) {
#pragma unroll
        for (int k = 0; k < DEPTH; ++k) mbar_init(&sh.full[k], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    uint32_t wnext = 0xffffffffu;              // loader: the window it will request next
    if (loader) {                              // windows 0 and 1 of this block
#pragma unroll
        for (int k = 0; k < DEPTH - 1; ++k)
            tma_issue_window<KB, SR>(sh, (uint32_t)k, in, tab, n, np, atomicAdd(next_window, 1u),
                                     tiles, total);
        wnext = atomicAdd(next_window, 1u);
    }

    for (uint32_t kk = 0;; ++kk) {
        const uint32_t xs = kk % DEPTH;
        QW_PROBE_T(c0);
        mbar_wait(&sh.full[xs], (kk / DEPTH) & 1u);            // window kk and its entry landed
        QW_PROBE_T(c1);
        const uint32_t idx = sh.widx[kk % (DEPTH + 1)];
        if (idx == 0xffffffffu) break;                         // no window left (block-uniform)
        const uint32_t local = idx / tiles, tile = idx - local * tiles;
        const int64_t first = (int64_t)tile * TILE - HALO;     // window start (may be < 0)
        // position of the window within its 128-site block -> where the pads sit in the slot
        const int fm = (int)((first + TMA_PADQ) % TMA_PADQ);   // first >= -HALO > -128
        double2 *seg = &sh.xbuf[xs][sg * SR + (fm + sg * SR) / TMA_PADQ];
        HState<SR> s;
#pragma unroll
        for (int r = 0; r < SR; ++r) {
            const double2 v = seg[r];
            s.yr[r] = v.x; s.yi[r] = v.y;
            s.zr[r] = 0.0; s.zi[r] = 0.0;
        }
        const TmaHead &pe = sh.tab[kk % (DEPTH + 1)];
        const int kb = pe.kb;
        fst[mt] = make_double2(s.yr[0], s.yi[0]);
        lst[mt] = make_double2(s.yr[SR - 1], s.yi[SR - 1]);
        if (!LATE && loader) {
            if (kk > 0) bulk_wait_read0();     // my store of window kk - 1 has left its slot
            tma_issue_window<KB, SR>(sh, kk + DEPTH - 1, in, tab, n, np, wnext, tiles, total);
            if (wnext < total) wnext = atomicAdd(next_window, 1u);
        }
        // the marked vertex is logical site HW0 of the rotated ring: slot HW0 of the thread whose
        // segment starts at a multiple of n
        int64_t wrel = HW0 - first - (int64_t)sg * SR;
        if (wrel < 0) wrel += n;
        if (wrel >= n) wrel -= n;
        const bool own = wrel == HW0;
        const bool warp_owns = __ballot_sync(0xffffffffu, own) != 0u;
        if (warp_owns && kb > 0) {             // its correction coefficients: HBM -> shared memory
            const double2 *cg = tab[local].c;
            for (int i = t & 31; i < 10 * kb; i += 32) sh.cown[i] = cg[i];
        }
        __syncthreads();                       // edge values (and coefficients) published
        QW_PROBE_T(c2);

        if (kb > 0) {
            if (warp_owns)
                tma_steps<SR, true>(s, pe, sh.cown, kb, fst, lst, MB, mt, ml, mr, own);
            else
                tma_steps<SR, false>(s, pe, sh.cown, kb, fst, lst, MB, mt, ml, mr, false);
        }

        // registers -> the thread's place in the slot; then thread 0: central part -> HBM, and
        // (LATE) thread 32: the request for window kk + 2 into the slot window kk - 1 has left
        // (thread 0 makes sure of that before the barrier: its store of window kk - 1 is a
        // window old)
        QW_PROBE_T(c3);
#pragma unroll
        for (int r = 0; r < SR; ++r) seg[r] = make_double2(s.yr[r], s.yi[r]);
        fence_proxy_async();
        if (LATE && storer) bulk_wait_read0();
        __syncthreads();
        QW_PROBE_T(c4);
        if (storer) {
            int64_t a = first + HALO, b = first + HALO + TILE;     // sites [a, b) of the ring
            if (b > n) b = n;                                      // last tile: the ring ends
            if (b > a) {
                // padded positions relative to the window start; floor division for first < 0
                const int64_t base = first + (first >= 0 ? first / TMA_PADQ : -1);
                const int64_t pa = tma_padded(a), pb = tma_padded(b);
                bulk_store(out + (int64_t)local * np + pa, sh.xbuf[xs] + (pa - base),
                           (uint32_t)(pb - pa) * (uint32_t)sizeof(double2));
            }
            bulk_commit();
        }
        if (LATE && loader) {
            tma_issue_window<KB, SR>(sh, kk + DEPTH - 1, in, tab, n, np, wnext, tiles, total);
            if (wnext < total) wnext = atomicAdd(next_window, 1u);
        }
        QW_PROBE_T(c5);
        QW_PROBE_ADD(0, c0, c1);
        QW_PROBE_ADD(1, c1, c2);
        QW_PROBE_ADD(2, c2, c3);
        QW_PROBE_ADD(3, c3, c4);
        QW_PROBE_ADD(4, c4, c5);
        QW_PROBE_ADD(5, 0, 1);
    }
    if (storer) bulk_wait0();
#ifdef QW_PROBE
    if (t == 0 && blockIdx.x < 1024) {
        unsigned long long ns;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns));
        qw_probe_blocks[3 * blockIdx.x + 2] = ns;
    }
#endif
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
    const int64_t pq = prm.stream_pad;                     // one pad element after every pq sites
    const double2 *src = buf + local * (pq ? n + n / pq : n);
    double nrm = 0.0;
    for (int64_t j = threadIdx.x; j < n; j += blockDim.x) {
        const double2 v = src[pq ? j + j / pq : j];
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
        const double2 v = src[pq ? lw + lw / pq : lw];
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

int sm_count() { return qw_sm_count(); }

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
           prm.n % TMA_PADQ == 0 && prm.n >= 1024 && sr > 0 &&
           !((prm.flags & QW_FLAG_MIRROR) && prm.n >= 8192);
}

template <int KB, int SR>
cudaError_t launch_pass_tma(const QwParams &prm, const double2 *in, double2 *out, void *tab,
                            int64_t step0, int64_t points, cudaStream_t st)
{
    // the window counter of this pass sits behind the pass entries (zeroed by the table kernel)
    unsigned int *counter = reinterpret_cast<unsigned int *>(
        static_cast<char *>(tab) + (size_t)points * sizeof(HPassEntry));
    constexpr int TILE = Win<KB, SR, true>::TILE;
    HPassEntry *etab = static_cast<HPassEntry *>(tab);
    stream_pass_tables_horner<<<(unsigned)((points + 127) / 128), 128, 0, st>>>(prm, points, step0,
                                                                               KB, etab, counter);
    const uint32_t tiles = (uint32_t)((prm.n + TILE - 1) / TILE);
    const uint64_t total = (uint64_t)points * tiles;
    uint32_t grid = (uint32_t)(sm_count() * Geo<SR>::BLOCKS_PER_SM);      // persistent
    if (total < grid) grid = (uint32_t)total;
    cudaError_t e = configure_tma<KB, SR>();
    if (e != cudaSuccess) return e;
    rk4_cycle_stream_tma<KB, SR><<<grid, Geo<SR>::NT, sizeof(TmaShared<SR>), st>>>(
        prm, in, out, etab, tiles, (uint32_t)total, counter);
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

#ifdef QW_PROBE
extern "C" int qw_debug_probe_blocks(unsigned long long *out)
{
    return (int)cudaMemcpyFromSymbol(out, qw_probe_blocks, sizeof(unsigned long long) * 3 * 1024);
}

extern "C" int qw_debug_probe(unsigned long long *out, int reset)
{
    cudaError_t e = cudaMemcpyFromSymbol(out, qw_probe, sizeof(unsigned long long) * 8);
    if (e == cudaSuccess && reset) {
        unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        e = cudaMemcpyToSymbol(qw_probe, z, sizeof(z));
    }
    return (int)e;
}
#endif

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
    prm.stream_pad = stream_tma(prm, sr) ? TMA_PADQ : 0;
    // stored elements per point
    const int64_t n = prm.mirror_m ? prm.mirror_m + 3
                                   : (prm.stream_pad ? tma_padded(prm.n) : prm.n);
    // aligned layout: the marked vertex is logical site 0 (stage form) or HW0 (nested form)
    prm.stream_rot = 0;
    if (!prm.mirror_m && prm.n % sr == 0) {
        prm.stream_rot = prm.target - (stream_horner(prm, sr) ? HW0 : 0);
        if (prm.stream_rot < 0) prm.stream_rot += prm.n;
    }
    // points in flight: bounded by ~8 GiB of ping-pong scratch and by 32-bit window indices
    int64_t chunk = (int64_t)(8ull << 30) / (2 * n * (int64_t)sizeof(double2));
    const int64_t tiles_max = n / (1024 - 64) + 2;
    if (chunk > (int64_t)0x7fffffff / tiles_max) chunk = (int64_t)0x7fffffff / tiles_max;
    if (chunk < 1) chunk = 1;                  // one point above the 8 GiB budget: try anyway
    if (chunk > prm.n_points) chunk = prm.n_points;
    {
        size_t free_b = 0, total_b = 0;        // ... unless not even one point fits the device
        if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess &&
            (size_t)2 * (size_t)n * sizeof(double2) > free_b / 8 * 7)
            return cudaErrorMemoryAllocation;
    }
    double2 *bufA = nullptr;
    unsigned long long *dmax = nullptr;
    cudaError_t e = qw_scratch_alloc((void **)&bufA, (size_t)chunk * 2 * n * sizeof(double2), st);
    if (e != cudaSuccess) return e;
    double2 *bufB = bufA + chunk * n;
    void *tab = nullptr;
    const size_t entry = sizeof(HPassEntry) > sizeof(PassEntry) ? sizeof(HPassEntry) : sizeof(PassEntry);
    e = qw_scratch_alloc(&tab, (size_t)chunk * entry + 64, st);    // + the window counter
    if (e != cudaSuccess) { qw_scratch_free(bufA, st); return e; }
    const bool varying = prm.step_size > 0.0 || prm.n_steps != nullptr;
    if (varying) {
        e = qw_scratch_alloc((void **)&dmax, sizeof(unsigned long long), st);
        if (e != cudaSuccess) { qw_scratch_free(bufA, st); qw_scratch_free(tab, st); return e; }
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
    qw_scratch_free(bufA, st);
    qw_scratch_free(tab, st);
    if (dmax) qw_scratch_free(dmax, st);
    return e;
}
