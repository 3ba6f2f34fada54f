// Register-resident RK4 in nested (Horner) form, one (T, gamma) point per thread block.
//
// Same integrator as qw_resident.cu -- fixed-step RK4 of QW_Search.py:149-164 over the
// implicit H(t) = a(t) L + d(t)|w><w| of BCB.py:28-99 -- evaluated with fewer FP64
// instructions.  With M = -iL and Q = -i|w><w| a step is
//     y+ = y + sum_s b_s F_s u_s,  u_1 = y,  u_{s+1} = y + c_s F_s u_s,  F_s = a_s M + d_s Q.
// Without Q every F_s is a multiple of M and the step operator is the polynomial
// I + b1 M + b2 M^2 + b3 M^3 + b4 M^4, evaluated as four nested axpy's
//     z4 = y + r4 M y,  z3 = y + r3 M z4,  z2 = y + r2 M z3,  y+ = y + r1 M z2
// (r1 = b1, r2 = b2/b1, r3 = b3/b2, r4 = b4/b3): per site and level one DADD and two DFMA per
// real component, 24 FP64 instructions per site-update instead of 30, and no accumulator
// vector (two complex vectors in registers instead of three).
// The oracle term does not commute with L, but every word of the step operator that
// contains a Q ends in a multiple of M^k |w>, k = 0..3.  Adding a complex scalar sigma_l to
// the marked vertex of z4, z3, z2 and y+ (the outer M's spread it over w-3..w+3) therefore
// restores the RK4 step exactly.  The four scalars are linear in
// (y_w, (Ly)_w, (L^2 y)_w, (L^3 y)_w) with coefficients that depend on the step's operator
// scalars only; they are tabulated per step next to r1..r4 by the block's own threads
// (one thread per step of a 32-step chunk).  tests/horner_model.py is the numpy statement of
// this algebra, checked against the oracle on the CPU.  How the four scalars are evaluated and
// applied at run time -- as the ADDEND y_w + sigma_l of the level's axpy at the marked vertex,
// from dot products over the window y_{w-3} .. y_w that one warp forms with all 32 lanes, split
// over an owner and a helper warp -- is described at HornerStep (qw_horner_level.cuh) and at
// horner_steps below.
//
// Any dimension: n = nact (R - 1) + nfull sites are dealt out as R sites to the first nfull
// threads and R - 1 to the others (possible whenever n >= R^2).  The last slot of a short
// thread is a ghost: before every level it is loaded with the right neighbour's first value,
// so that the thread's last real site sees the correct neighbour in the unrolled stencil; a
// short thread publishes slot R-2 as its last value.  Two selects per level, no FP64 work.
//
// Layout: the ring is relabelled so that the marked vertex is slot HW0 = 3 of thread 0, i.e.
// thread 0 holds w-3..w+R-4 and all the oracle work is local to it.  Halo values cross
// threads through a double-buffered shared-memory mailbox, one __syncthreads() per level;
// edge slots are updated and published first.
#include "qw_horner_level.cuh"

namespace {

// CH steps per table refill, one thread per step (a refill is ~800 instructions in the lanes that
// have a step, while the other warps wait; 64 instead of 32 steps per refill measured no gain)
template <int CH>
struct HornerShared {
    HornerStep st[CH];
    double red[2 * 32];
    double2 win[4];                      // slots 0 .. 3 of the owner thread (y_{w-3} .. y_w; half
                                         // ring: y_w .. y_{w+3}), published by the owner; not [W0]
    double2 priv[32];                    // the owner WARP's y[W0], lane by lane
    double tq[8];                        // the helper warp's outputs (-, t_3, t_2, t_1)
    double2 zero;
    int owner;
};

// One thread per step of the chunk; stage times as in qw_fill_table (t_n = n h by
// multiplication, tau = t / T by division).  Caller synchronises before and after.
__device__ __forceinline__ void horner_fill_cycle(HornerStep *tab, const QwPoint &pt,
                                                  const int64_t n0, const int cnt,
                                                  const int schedule, const int gamma_on,
                                                  const bool fold4)
{
    for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
        const double t0 = (double)(n0 + i) * pt.h;
        const double2 ad1 = qw_operator_scalars(t0 / pt.T, pt.gamma, schedule, gamma_on);
        const double2 ad2 = qw_operator_scalars((t0 + 0.5 * pt.h) / pt.T, pt.gamma, schedule,
                                                gamma_on);
        const double2 ad4 = qw_operator_scalars((t0 + pt.h) / pt.T, pt.gamma, schedule, gamma_on);
        HornerCoef hc;
        horner_coefficients(hc, pt.h, ad1, ad2, ad4, 2.0, 6.0);     // L_ww = 2, (L^2)_ww = 6
        horner_cycle_sum_basis(hc.c);
        horner_step_store(tab[i], hc, fold4);
    }
}

__device__ __forceinline__ double2 hblock_sum(double2 v, double *red)
{
    v.x = qw_warp_sum(v.x);
    v.y = qw_warp_sum(v.y);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nwarps = (blockDim.x + 31) >> 5;
    if (lane == 0) { red[2 * warp] = v.x; red[2 * warp + 1] = v.y; }
    __syncthreads();
    double2 w = make_double2(0.0, 0.0);
    if (lane < nwarps) w = make_double2(red[2 * lane], red[2 * lane + 1]);
    w.x = qw_warp_sum(w.x);
    w.y = qw_warp_sum(w.y);
    return w;
}

// Physical warp slot of the calling warp.  Warps are assigned to the four SM sub-partitions
// (each with its own FP64 lanes) by slot number modulo 4.
__device__ __forceinline__ unsigned warp_slot()
{
    unsigned w;
    asm volatile("mov.u32 %0, %%warpid;" : "=r"(w));
    return w;
}

// `cnt` RK4 steps with the table in shared memory.
// ROLE of the calling warp in the oracle correction (see HornerStep).  The FP64 pipe of an SM
// sub-partition is DISPATCH bound: an FP64 warp-instruction takes two issue cycles, every other
// instruction one, and the warps of a block move in lockstep through the level barriers, so what
// counts is the instruction total of the LONGEST warp (tools/hot_probe.cu: with all the
// correction work in the owner warp the kernel lost 9 %, half of it to the imbalance).  Hence
//   HR_OWNER  publishes y_{w-3..w+3} before the step's first barrier, picks up sigma_4, t_3, t_2,
//             t_1 after the second and uses t_l = y_w + sigma_l as the addend of its marked slot;
//   HR_HELPER (another warp of the block) forms the four dot products during level 4;
//   HR_SELF   one-warp blocks: both duties, every t_l folded (level 4 included);
//   HR_PLAIN  everybody else.
// A step has four levels, so the mailbox parity of a level is a compile-time constant: level 4
// and 2 read buffer 0 and write buffer 1, level 3 and 1 the other way round.
// (A pairwise bar.arrive / bar.sync handshake between the two warps of a block instead of the
// block barrier was measured: equal at best, -6 % once the publication was pinned early.
// ptxas moves arithmetic across BAR.ARV / BAR.SYNC freely: whatever the PTX order, the interior
// arithmetic of a level ends up after the *next* wait, so an arrival is always followed at once
// by the wait for the partner's arrival of the same level -- a plain barrier.  DESIGN.md 7.)
// RAGGED: some threads hold R-1 sites (`full` false); a compile-time switch, because even two
// extra selects per level cost the exact-multiple sizes 5 % (register pressure at 168).
// MIRROR (QW_FLAG_MIRROR): the threads hold the half ring w .. w + n/2; the marked vertex is
// slot 0 of thread 0, whose left neighbour is its own slot 1; the last thread's right neighbour
// is the mirror image about the antipode (mend = 1: even n, the site before the last; mend = 2:
// odd n, the last site itself; 0: not the last thread).
// hooks for tools/hot_probe.cu (ablation timing); the library never defines them
#ifndef QW_LEVEL_BAR
#define QW_LEVEL_BAR() __syncthreads()
#endif
#ifndef QW_OCC_R16_64
#define QW_OCC_R16_64 6                  // blocks per SM of the 64-thread R = 16 cycle kernel
#endif
#ifndef QW_HCHUNK
#define QW_HCHUNK 32                     // steps per table refill in blocks of two warps and more
#endif
constexpr int HR_PLAIN = 0, HR_SELF = 1, HR_OWNER = 2, HR_HELPER = 3;

// Before the first step: the owner warp's lanes store their y[W0], the owner its window.
template <int R, bool MIRROR>
__device__ __forceinline__ void horner_first_window(const HState<R> &s, double2 *win,
                                                    double2 *priv, const int lane, const bool own)
{
    constexpr int W0 = MIRROR ? 0 : HW0;
    priv[lane] = make_double2(s.yr[W0], s.yi[W0]);
    if (own) {
#pragma unroll
        for (int r = 0; r < 4; ++r) win[r] = make_double2(s.yr[r], s.yi[r]);
    }
}

struct HornerOracle {                    // shared-memory pieces of the oracle correction
    double2 *win, *priv;
    double *tq;
    const double2 *zero;
    bool helps;                          // HR_HELPER: this warp's dot products are the ones stored
};

template <int R, int ROLE, bool RAGGED, bool MIRROR>
__device__ __forceinline__ void horner_steps(HState<R> &s, const HornerStep *tab, double2 *first,
                                             double2 *last, const int nt, const int t,
                                             const int tl, const int tr, const int cnt,
                                             const bool own, const bool full_in, const int mend,
                                             const HornerOracle &oc)
{
    constexpr int W0 = MIRROR ? 0 : HW0;
    constexpr bool OWNS = ROLE == HR_SELF || ROLE == HR_OWNER;
    constexpr bool OWNER = ROLE == HR_OWNER;
    constexpr bool HELPS = ROLE == HR_HELPER;
    constexpr int WS = OWNS ? W0 : -1;
    constexpr int PUBW = OWNS ? (MIRROR ? 4 : HW0) : 0;     // window: slots 0 .. 3 (W0 is in priv)
    const bool full = RAGGED ? full_in : true;
    const int lane = t & 31;
    // owner warp: where a lane finds the addend of its marked slot at levels 3, 2, 1 -- the owner
    // lane the helper's t_3, t_2, t_1, every other lane its own y[W0]
    const uint32_t a_priv = smem_u32(oc.priv + lane);
    const uint32_t a_t = (OWNS && own) ? smem_u32(oc.tq) : a_priv;
    const uint32_t a_ts = (OWNS && own) ? 16u : 0u;
    const uint32_t a_win = smem_u32(oc.win);
    // helper (and HR_SELF): the window site y_{w-q} of lane 8q + l
    const int q = lane >> 3;
    const uint32_t wa = horner_window_addr<MIRROR>(MIRROR ? q : HW0 - q, oc.win, oc.priv);
    for (int sidx = 0; sidx < cnt; ++sidx) {
        const HornerStep &hs = tab[sidx];
#define QW_LX_R(LV, r) ((LV == 4) ? s.yr[r] : s.zr[r])
#define QW_LX_I(LV, r) ((LV == 4) ? s.yi[r] : s.zi[r])
#define QW_HORNER_LEVEL(LEVEL)                                                              \
        {                                                                                   \
            constexpr int par = (4 - LEVEL) & 1;                                            \
            QW_LEVEL_BAR();                                                                 \
            double2 hl = last[par * nt + tl];                                               \
            double2 hr = first[par * nt + tr];                                              \
            const double rho = hs.rho[LEVEL - 1];                                           \
            double2 tw = make_double2(0.0, 0.0);    /* owner warp: addend of the marked slot */ \
            if (OWNER && LEVEL < 4) tw = lds2(a_t + (4 - LEVEL) * a_ts);                    \
            if (OWNER && LEVEL == 4) {              /* y_w + c4 y_w; c4 = 0 in the other lanes */ \
                const double2 c4 = *(own ? &hs.c4 : oc.zero);                               \
                tw = cmul2_acc(c4, make_double2(s.yr[W0], s.yi[W0]),                        \
                               make_double2(s.yr[W0], s.yi[W0]));                           \
            }                                                                               \
            double2 o0, oL;                                                                 \
            if (MIRROR) {                     /* reflecting ends of the half ring */        \
                if (t == 0) hl = make_double2(QW_LX_R(LEVEL, 1), QW_LX_I(LEVEL, 1));                            \
                if (mend == 1)                                                              \
                    hr = full ? make_double2(QW_LX_R(LEVEL, R - 2), QW_LX_I(LEVEL, R - 2))                      \
                              : make_double2(QW_LX_R(LEVEL, R - 3), QW_LX_I(LEVEL, R - 3));                     \
                if (mend == 2)                                                              \
                    hr = full ? make_double2(QW_LX_R(LEVEL, R - 1), QW_LX_I(LEVEL, R - 1))                      \
                              : make_double2(QW_LX_R(LEVEL, R - 2), QW_LX_I(LEVEL, R - 2));                     \
            }                                                                               \
            if (RAGGED && !full) {            /* ghost slot := the right neighbour's value */ \
                if (LEVEL == 4) { s.yr[R - 1] = hr.x; s.yi[R - 1] = hr.y; }                 \
                else { s.zr[R - 1] = hr.x; s.zi[R - 1] = hr.y; }                            \
            }                                                                               \
            if (ROLE == HR_SELF) {            /* the owner warp does it all: every addend from */ \
                if (LEVEL == 4) {             /* the dot product, through shared memory         */ \
                    const double v = horner_window_dot(wa, hs, lane);                       \
                    if (lane < 8) oc.tq[lane] = v;                                          \
                    __syncwarp();                                                           \
                }                                                                           \
                tw = lds2(a_t + (4 - LEVEL) * a_ts);                                        \
            }                                                                               \
            if (HELPS && LEVEL == 4) {                                                      \
                const double v = horner_window_dot(wa, hs, lane);                       \
                if (lane < 8 && oc.helps) oc.tq[lane] = v;                                  \
            }                                                                               \
            hedges<R, LEVEL, WS, PUBW>(s, rho, hl, hr, o0, oL, tw.x, tw.y, a_win, (int)own); \
            hinterior<R, LEVEL, WS, PUBW>(s, rho, o0, oL, tw.x, tw.y, a_win, (int)own);     \
            if (OWNS && LEVEL == 1) {         /* the next step's y[W0], lane by lane */     \
                sts2(a_priv, s.yr[W0], s.yi[W0]);                                           \
                if (ROLE == HR_SELF) __syncwarp();                                          \
            }                                                                               \
            {   /* publish: the last REAL site is slot R-1, or R-2 in a short thread */     \
                const double lr = (LEVEL > 1) ? (full ? s.zr[R - 1] : s.zr[R - 2])          \
                                              : (full ? s.yr[R - 1] : s.yr[R - 2]);         \
                const double li = (LEVEL > 1) ? (full ? s.zi[R - 1] : s.zi[R - 2])          \
                                              : (full ? s.yi[R - 1] : s.yi[R - 2]);         \
                first[(par ^ 1) * nt + t] = (LEVEL > 1) ? make_double2(s.zr[0], s.zi[0])    \
                                                        : make_double2(s.yr[0], s.yi[0]);   \
                last[(par ^ 1) * nt + t] = make_double2(lr, li);                            \
            }                                                                               \
        }
        QW_HORNER_LEVEL(4)
        QW_HORNER_LEVEL(3)
        QW_HORNER_LEVEL(2)
        QW_HORNER_LEVEL(1)
#undef QW_HORNER_LEVEL
#undef QW_LX_R
#undef QW_LX_I
    }
}

template <int R, int MAXNT, int MINB, bool RAGGED, bool MIRROR = false>
__global__ void __launch_bounds__(MAXNT, MINB) rk4_cycle_horner(const QwParams prm)
{
    static_assert(!MIRROR || RAGGED, "the half ring has an arbitrary number of sites");
    constexpr int W0 = MIRROR ? 0 : HW0;
    static_assert(R >= 2 * HW0 + 1, "the window w-3..w+3 must sit in one thread");
    extern __shared__ double2 mbox[];         // [2][2][MAXNT]: parity, first/last
    constexpr int CH = MAXNT >= 64 ? QW_HCHUNK : QW_CHUNK;
    __shared__ HornerShared<CH> sh;
    constexpr int NTS = MAXNT;                // mailbox stride: offsets become immediates
    const int t = threadIdx.x, nt = blockDim.x, nact = prm.nact, nfull = prm.nfull;
    const QwPoint pt = qw_point(prm, blockIdx.x);
    const bool active = t < nact;
    const bool full = !RAGGED || t < nfull || !active; // R sites (idle threads: R zeros)
    // ring neighbours among the active threads; idle threads talk to themselves (zeros)
    const int tl = active ? (t == 0 ? nact - 1 : t - 1) : t;
    const int tr = active ? (t == nact - 1 ? 0 : t + 1) : t;
    // The owner of the marked vertex is lane 0 of warp 0 -- or of another warp, rotating with
    // the block's physical warp slots: the owner warp issues ~14 % more FP64 instructions, and
    // warps go to the four SM sub-partitions by slot number modulo 4, so without the rotation
    // all owner warps of an SM would share one (4-warp blocks) or two (2-warp blocks) of them.
    int ht = 0;
    const int mend = (MIRROR && t == nact - 1) ? ((prm.n & 1) ? 2 : 1) : 0;
    if (!MIRROR && (nt == 64 || nt == 128) && nact == nt && !(prm.flags & QW_FLAG_NO_ROTATE)) {
        if (t == 0) sh.owner = 32 * (int)((warp_slot() >> 2) & (nt == 64 ? 1u : 3u));
        __syncthreads();
        ht = sh.owner;
    }
    const bool own = (t == ht);

    HState<R> s;
    const double amp = active ? 1.0 / sqrt((double)prm.n) : 0.0;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        s.yr[r] = amp; s.yi[r] = 0.0;
        s.zr[r] = 0.0; s.zi[r] = 0.0;
    }
    double2 *first = mbox, *last = mbox + 2 * NTS;     // [parity][thread]; a step starts at 0
    first[t] = make_double2(s.yr[0], s.yi[0]);
    last[t] = make_double2(s.yr[R - 1], s.yi[R - 1]);  // uniform start: any slot will do

    const int warp = t >> 5;
    const int helper = ((ht >> 5) + 1) % (nt >> 5);    // the warp that forms the dot products
    // Blocks of five warps and more: the owner warp does all of it (HR_SELF), as a one-warp block
    // does.  Measured (tools/r2_ab_sizes.sh): N = 3000 5.40e11 against 5.18e11 with the split,
    // N = 4096 6.30e11 against 6.0-6.2e11; N = 2048 (four warps) 6.47e11 against 6.64e11.
#ifdef QW_ABL_SELFISH
    const bool selfish = true;
#else
    const bool selfish = nt > 128;
#endif
    const HornerOracle oc = {sh.win, sh.priv, sh.tq, &sh.zero, warp == helper};
    if (t == 0) sh.zero = make_double2(0.0, 0.0);
    if (warp == (ht >> 5)) horner_first_window<R, MIRROR>(s, sh.win, sh.priv, t & 31, own);
    for (int64_t n0 = 0; n0 < pt.steps; n0 += CH) {
        const int cnt = (int)((pt.steps - n0 < CH) ? (pt.steps - n0) : CH);
        __syncthreads();                               // the last chunk's table is no longer read
#ifdef QW_ABL_NOREFILL
        if (n0 == 0)
#endif
        horner_fill_cycle(sh.st, pt, n0, cnt, prm.schedule, prm.gamma_on, nt == 32 || selfish);
        __syncthreads();
#define QW_HS(ROLE) horner_steps<R, ROLE, RAGGED, MIRROR>(s, sh.st, first, last, NTS, t, tl, tr, \
                                                          cnt, own, full, mend, oc)
#if defined(QW_ABL_NOOWN)
        QW_HS(HR_PLAIN);
#else
        if (nt == 32) QW_HS(HR_SELF);
        else if (selfish) { if (warp == (ht >> 5)) QW_HS(HR_SELF); else QW_HS(HR_PLAIN); }
        else if (warp == (ht >> 5)) QW_HS(HR_OWNER);
        else if (warp == helper) QW_HS(HR_HELPER);
        else QW_HS(HR_PLAIN);
#endif
#undef QW_HS
    }

    // p = |psi_w|^2 / ||psi||^2, ||psi||^2 (BCB.py:189-199, Runge_Kutta_Error.py:103)
    double nrm = 0.0;
#pragma unroll
    for (int r = 0; r < R - 1; ++r) nrm = fma(s.yr[r], s.yr[r], fma(s.yi[r], s.yi[r], nrm));
    if (full) nrm = fma(s.yr[R - 1], s.yr[R - 1], fma(s.yi[R - 1], s.yi[R - 1], nrm));
    const double pw = own ? fma(s.yr[W0], s.yr[W0], s.yi[W0] * s.yi[W0]) : 0.0;
    if (MIRROR) {          // interior sites of the half ring count twice; w and (even n) the
        nrm *= 2.0;        // antipode once
        nrm -= pw;
        if (mend == 1) {
            const double ar = full ? s.yr[R - 1] : s.yr[R - 2], ai = full ? s.yi[R - 1] : s.yi[R - 2];
            nrm -= fma(ar, ar, ai * ai);
        }
    }
    __syncthreads();
    const double2 tot = hblock_sum(make_double2(nrm, pw), sh.red);
    if (t == 0) {
        prm.p[pt.q] = tot.y / tot.x;
        if (prm.norm) prm.norm[pt.q] = tot.x;
    }
    if (prm.psi && active) {      // undo the relabelling: slot r of thread t -> site w + l,
        double2 *out = prm.psi + (size_t)pt.q * (size_t)prm.n;   // l = off(t) - off(ht) + r - HW0
        const int64_t off_t = (int64_t)t * (R - 1) + (t < nfull ? t : nfull);
        const int64_t off_h = (int64_t)ht * (R - 1) + (ht < nfull ? ht : nfull);
#pragma unroll
        for (int r = 0; r < R; ++r) {
            if (r == R - 1 && !full) continue;                    // ghost
            int64_t j = off_t - off_h + r - W0 + prm.target;
            if (j < 0) j += prm.n;
            if (j >= prm.n) j -= prm.n;
            out[j] = make_double2(s.yr[r], s.yi[r]);
            if (MIRROR) {                                         // the mirror image w - l
                int64_t jm = prm.target - (off_t + r);
                if (jm < 0) jm += prm.n;
                out[jm] = make_double2(s.yr[r], s.yi[r]);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Rings of 4097 .. 32768 sites: resident across a THREAD-BLOCK CLUSTER.  CTA r of a cluster of C
// holds sites [off_r, off_r + n_r) of the relabelled ring in registers, like the one-block kernel
// above; the two halo values that cross a CTA boundary travel through distributed shared memory.
//   * Layout of a CTA: [left edge site][compute threads: R or R - 1 sites each][right edge site].
//     The compute warps run the code of the one-block kernel unchanged (horner_steps); the two
//     EDGE SITES belong to lanes 0 and 1 of an extra COMM WARP and appear to compute thread 0 /
//     the last compute thread as ordinary mailbox neighbours.  (Two earlier versions kept the edge
//     sites in the compute threads: the cross-CTA wait and the divergent edge code then sat at the
//     end of every level of two of the eight warps, and with them of the whole block: 720 instead
//     of ~485 cycles per level, 3.9e11 site-updates/s at N = 8192.)
//   * Every CTA has, per side and mailbox parity, one 16-byte slot and one mbarrier.  Right after
//     the block barrier of a level the comm lane reads its inner neighbour from the mailbox, waits
//     for the neighbour CTA's value of the previous level (mbarrier.try_wait; it has had a whole
//     level to arrive: DSMEM latency ~215 cycles), updates its site and sends the new value on with
//     st.async: data + 16 bytes of the remote barrier's transaction count in one operation, no
//     fence (st.shared::cluster + mbarrier.arrive.release.cluster compiles to ST + MEMBAR.ALL.GPU
//     + SYNCS.ARRIVE; measured 2.5e11).  No cluster-wide barrier in the time loop.
//   * Two parities suffice: the neighbour can refill a slot only after it has received the value
//     this CTA publishes AFTER reading the slot.  The k-th read of the parity-0 slots happens at
//     level 4 / 2 of a step, of the parity-1 slots at level 3 / 1, so the mbarrier phase a level
//     waits for is a compile-time constant: 0, 0, 1, 1 for levels 4, 3, 2, 1.
// The marked vertex is slot 3 of compute thread 0 of CTA 0 (the left edge site of CTA 0 is site
// w - 4).  Every CTA builds the step tables for itself; the partial norms meet in CTA 0.
struct ClusterShared {
    HornerStep st[QW_CHUNK];
    double red[2 * 32];
    double2 win[4];                      // as in HornerShared
    double2 priv[32];
    double tq[8];
    double2 zero;
    double2 halo[2][2];                  // [left / right][parity]: written by the neighbour CTA
    unsigned long long bar[2][2];        // [left / right][parity]
    double2 csum[16];                    // CTA 0: (norm, |psi_w|^2) of every CTA
    double2 edge_out[2];                 // the comm lanes' sites at the end (norm, psi)
    int owner;
};

__device__ __forceinline__ uint32_t cl_smem(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t cl_rank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t cl_size()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cl_sync()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n"
                 "barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cl_map(uint32_t addr, uint32_t rank)
{
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
// value into the neighbour CTA's slot, signalled on ITS mbarrier
__device__ __forceinline__ void cl_send(uint32_t slot, uint32_t bar, double x, double y)
{
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f64 [%0], {%1, %2}, [%3];"
                 ::"r"(slot), "d"(x), "d"(y), "r"(bar) : "memory");
}
// The receiving lane arms its barrier for the 16 bytes (the only arrival of the phase; the bytes
// may already be there) and waits for the phase.
__device__ __forceinline__ void cl_wait(uint32_t bar, uint32_t parity)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], 16;" ::"r"(bar) : "memory");
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "QW_CL_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra QW_CL_DONE;\n"
        "bra QW_CL_WAIT;\n"
        "QW_CL_DONE:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}

// The comm warp's side of `cnt` steps: lane 0 owns the CTA's left edge site, lane 1 the right one
// (the other lanes only keep the block barriers company).  side = lane; me = mailbox index of the
// lane's site; inner = mailbox index of the compute thread next to it.
__device__ __forceinline__ void cluster_comm_steps(double &yr, double &yi, ClusterShared &sh,
                                                   double2 *first, double2 *last, const int nt,
                                                   const int side, const int me, const int inner,
                                                   const uint32_t my_bar0, const uint32_t my_bar1,
                                                   const uint32_t to_slot0, const uint32_t to_slot1,
                                                   const uint32_t to_bar0, const uint32_t to_bar1,
                                                   const int cnt, const bool lane_on)
{
    double zr = 0.0, zi = 0.0;
    for (int sidx = 0; sidx < cnt; ++sidx) {
        const HornerStep &hs = sh.st[sidx];
#define QW_COMM_LEVEL(LEVEL)                                                                \
        {                                                                                   \
            constexpr int par = (4 - LEVEL) & 1;                                            \
            constexpr uint32_t phase = (LEVEL <= 2) ? 1u : 0u;                              \
            __syncthreads();                                                                \
            if (lane_on) {                                                                  \
                /* inner neighbour: the compute thread's edge value of the level's input */ \
                const double2 in = side == 0 ? first[par * nt + inner] : last[par * nt + inner]; \
                cl_wait(par ? my_bar1 : my_bar0, phase);                                    \
                const double2 out = sh.halo[side][par];        /* the neighbour CTA's site */ \
                const double rho = hs.rho[LEVEL - 1];                                       \
                const double cr = (LEVEL == 4) ? yr : zr, ci = (LEVEL == 4) ? yi : zi;      \
                double o_r, o_i;                                                            \
                hsite(side == 0 ? out.x : in.x, side == 0 ? out.y : in.y, cr, ci,           \
                      side == 0 ? in.x : out.x, side == 0 ? in.y : out.y, rho, yr, yi, o_r, o_i); \
                if (LEVEL > 1) { zr = o_r; zi = o_i; } else { yr = o_r; yi = o_i; }         \
                cl_send(par ? to_slot0 : to_slot1, par ? to_bar0 : to_bar1, o_r, o_i);      \
                first[(par ^ 1) * nt + me] = make_double2(o_r, o_i);                        \
                last[(par ^ 1) * nt + me] = make_double2(o_r, o_i);                         \
            }                                                                               \
        }
        QW_COMM_LEVEL(4)
        QW_COMM_LEVEL(3)
        QW_COMM_LEVEL(2)
        QW_COMM_LEVEL(1)
#undef QW_COMM_LEVEL
    }
}

// blockDim = compute threads (rounded up to warps) + 32 (the comm warp)
// MAXNT / MINB: 288 threads, one CTA per SM (up to 4096 sites per CTA), or 160 threads, two CTAs
// per SM (up to 2050 sites per CTA, twice as many CTAs per cluster): two independent barrier
// domains per SM instead of one.
template <int R, bool RAGGED, int MAXNT, int MINB>
__global__ void __launch_bounds__(MAXNT, MINB) rk4_cycle_cluster(const QwParams prm)
{
    static_assert(R >= 2 * HW0 + 1, "the window w-3..w+3 must sit in one thread");
    extern __shared__ double2 mbox[];         // [2][2][blockDim.x]: parity, first/last
    __shared__ ClusterShared sh;
    const int t = threadIdx.x, nt = blockDim.x, ntc = nt - 32;       // ntc: compute threads
    const uint32_t rank = cl_rank(), csz = cl_size();
    const QwPoint pt = qw_point(prm, blockIdx.x / csz);
    // this CTA's part of the ring: nb sites, two of them edge sites
    const int64_t base = prm.n / csz, extra = prm.n - base * csz;
    const int64_t nb = base + (rank < extra ? 1 : 0);
    const int64_t off = rank * base + (rank < extra ? rank : extra);
    const int nact = (int)((nb - 2 + R - 1) / R);
    const int nfull = (int)(nb - 2 - (int64_t)nact * (R - 1));
    const bool comm = t >= ntc;
    const int lane = t & 31;
    const bool active = t < nact;
    const bool full = !RAGGED || t < nfull || !active;
    const int me_l = ntc, me_r = ntc + 1;              // mailbox indices of the two edge sites
    // neighbours: compute thread 0 / nact - 1 talk to the edge sites
    const int tl = active ? (t == 0 ? me_l : t - 1) : t;
    const int tr = active ? (t == nact - 1 ? me_r : t + 1) : t;

    if (t == 0) {
#pragma unroll
        for (int d = 0; d < 2; ++d)
#pragma unroll
            for (int p = 0; p < 2; ++p)
                asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(cl_smem(&sh.bar[d][p])) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }

    HState<R> s;
    const double amp0 = 1.0 / sqrt((double)prm.n);
    const double amp = active ? amp0 : 0.0;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        s.yr[r] = amp; s.yi[r] = 0.0;
        s.zr[r] = 0.0; s.zi[r] = 0.0;
    }
    double2 *first = mbox, *last = mbox + 2 * nt;      // [parity][thread]; a step starts at 0
    first[t] = make_double2(comm ? amp0 : s.yr[0], 0.0);
    last[t] = make_double2(comm ? amp0 : s.yr[R - 1], 0.0);   // uniform start: any slot will do
    cl_sync();                                         // every CTA's barriers exist

    // comm lane 0 / 1: the left / right edge site and its links
    const bool lane_on = comm && lane < 2;
    const int side = lane & 1;
    uint32_t my_bar0 = 0, my_bar1 = 0, to_slot0 = 0, to_slot1 = 0, to_bar0 = 0, to_bar1 = 0;
    double er = amp0, ei = 0.0;                        // the edge site (psi_n)
    if (lane_on) {
        const uint32_t nbr = side == 0 ? (rank == 0 ? csz - 1 : rank - 1)
                                       : (rank + 1 == csz ? 0 : rank + 1);
        my_bar0 = cl_smem(&sh.bar[side][0]);
        my_bar1 = cl_smem(&sh.bar[side][1]);
        // my left site is the RIGHT-hand halo of the CTA to the left, and vice versa
        to_slot0 = cl_map(cl_smem(&sh.halo[side ^ 1][0]), nbr);
        to_slot1 = cl_map(cl_smem(&sh.halo[side ^ 1][1]), nbr);
        to_bar0 = cl_map(cl_smem(&sh.bar[side ^ 1][0]), nbr);
        to_bar1 = cl_map(cl_smem(&sh.bar[side ^ 1][1]), nbr);
        if (pt.steps > 0) cl_send(to_slot0, to_bar0, er, ei);        // fill #0 of the parity-0 slots
    }

    const int warp = t >> 5;
    const bool owner_cta = rank == 0;
    const bool own = owner_cta && t == 0;
    const HornerOracle oc = {sh.win, sh.priv, sh.tq, &sh.zero, owner_cta && warp == 1};
    if (t == 0) sh.zero = make_double2(0.0, 0.0);
    if (owner_cta && warp == 0) horner_first_window<R, false>(s, sh.win, sh.priv, t & 31, own);
    for (int64_t n0 = 0; n0 < pt.steps; n0 += QW_CHUNK) {
        const int cnt = (int)((pt.steps - n0 < QW_CHUNK) ? (pt.steps - n0) : QW_CHUNK);
        __syncthreads();                               // the last chunk's table is no longer read
        horner_fill_cycle(sh.st, pt, n0, cnt, prm.schedule, prm.gamma_on, ntc > 128);
        __syncthreads();
#define QW_HS(ROLE) horner_steps<R, ROLE, RAGGED, false>(s, sh.st, first, last, nt, t, tl, tr, \
                                                         cnt, own, full, 0, oc)
        if (comm)
            cluster_comm_steps(er, ei, sh, first, last, nt, side, side == 0 ? me_l : me_r,
                               side == 0 ? 0 : nact - 1, my_bar0, my_bar1, to_slot0, to_slot1,
                               to_bar0, to_bar1, cnt, lane_on);
        else if (owner_cta && warp == 0 && ntc > 128) QW_HS(HR_SELF);   // as the one-block kernel
        else if (owner_cta && warp == 0) QW_HS(HR_OWNER);
        else if (owner_cta && warp == 1 && ntc <= 128) QW_HS(HR_HELPER);   // >= 2 compute warps
        else QW_HS(HR_PLAIN);
#undef QW_HS
    }

    // p = |psi_w|^2 / ||psi||^2, ||psi||^2: partial sums to CTA 0
    double nrm = 0.0;
    if (!comm) {
#pragma unroll
        for (int r = 0; r < R - 1; ++r) nrm = fma(s.yr[r], s.yr[r], fma(s.yi[r], s.yi[r], nrm));
        if (full) nrm = fma(s.yr[R - 1], s.yr[R - 1], fma(s.yi[R - 1], s.yi[R - 1], nrm));
    } else if (lane_on) {
        nrm = fma(er, er, ei * ei);
        sh.edge_out[side] = make_double2(er, ei);
    }
    const double pw = (owner_cta && t == 0) ? fma(s.yr[HW0], s.yr[HW0], s.yi[HW0] * s.yi[HW0]) : 0.0;
    __syncthreads();
    const double2 tot = hblock_sum(make_double2(nrm, pw), sh.red);
    if (t == 0) {
        const uint32_t dst = cl_map(cl_smem(&sh.csum[rank]), 0);
        asm volatile("st.shared::cluster.v2.f64 [%0], {%1, %2};" ::"r"(dst), "d"(tot.x), "d"(tot.y) : "memory");
    }
    cl_sync();                                         // all partial sums (and all halo traffic) done
    if (owner_cta && t == 0) {
        double nsum = 0.0, psum = 0.0;
        for (uint32_t r = 0; r < csz; ++r) { nsum += sh.csum[r].x; psum += sh.csum[r].y; }
        prm.p[pt.q] = psum / nsum;
        if (prm.norm) prm.norm[pt.q] = nsum;
    }
    if (prm.psi) {
        // undo the relabelling: logical site l of the ring (CTA 0: l = 0 is its left edge site,
        // l = 1 + HW0 the marked vertex) -> physical site w + l - 1 - HW0
        double2 *out = prm.psi + (size_t)pt.q * (size_t)prm.n;
        const int64_t shift = (int64_t)prm.target - 1 - HW0;
        if (active) {
            const int64_t off_t = off + 1 + (int64_t)t * (R - 1) + (t < nfull ? t : nfull);
#pragma unroll
            for (int r = 0; r < R; ++r) {
                if (r == R - 1 && !full) continue;                // ghost
                int64_t j = (off_t + r + shift) % prm.n;
                if (j < 0) j += prm.n;
                out[j] = make_double2(s.yr[r], s.yi[r]);
            }
        }
        if (lane_on) {
            int64_t j = (off + (side == 0 ? 0 : nb - 1) + shift) % prm.n;
            if (j < 0) j += prm.n;
            out[j] = make_double2(er, ei);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Complete graph: (L x)_j = N x_j - sum(x).  A nested level z' = y - i rho (N z - S_z) touches no
// other site, and the sum of every level input follows from the marked vertex alone because
// the column sums of L vanish: S(z_l) = S(y) + sigma_l, S(y+) = S(y) + sigma_1, with sigmas that
// are linear in (y_w, (Ly)_w) ((L^2 y)_w = N (Ly)_w, (L^3 y)_w = N^2 (Ly)_w).  The pair
// (y_w, S(y)) is therefore a closed two-variable system.  It is integrated a whole 32-step chunk
// AHEAD of the vector work (by the "scalar warp": a warp of its own above 3072 sites, the last
// worker warp below, see the two kernels) and the level means and corrections of the coming
// chunk are left in shared memory; the tables are refilled two chunks ahead.  The worker warps
// meet at ONE barrier per chunk and never wait inside it.  16 FP64 instructions per site-update
// (classic stage form: 22, one barrier per stage).
// The tracked S(y) is never re-anchored to a direct sum of the state: the difference e between
// the two obeys e+ = Phi0 e with the step polynomial at the top eigenvalue of L, |Phi0| <= 1
// inside the stability region of RK4, so rounding noise does not grow -- while a correction
// that arrives a chunk late has rotated by 32 steps' worth of phase and DOES grow (measured).
struct CHStep {                          // read by the workers
    double rho[4];
};

// The marked-vertex subsystem is LINEAR in x = (y_w, S(y)): everything the scalar warp owes the
// workers for a step, and the state after the step, is out_o = q[o][0] y_w + q[o][1] S(y).
//   o = 0..3: (S(y), S(z4), S(z3), S(z2)) / N  -- the level means the workers subtract
//   o = 4..7: sigma_4 .. sigma_1                -- the owner thread's corrections
//   o = 8, 9: y_w and S(y) after the step
// The maps are tabulated with the coefficients (lane-parallel over the steps of a chunk), so the
// only sequential work per step is one 20 x 4 real matrix-vector product spread over 20 lanes:
// 5 FP64 warp-instructions and one shuffle stage (the level-by-level form it replaces: 36
// instructions in a dependent chain 16 deep, which under the workers' FP64 load took as long
// as the workers' own step and made them wait at the chunk barrier).
struct CSMap {
    double2 q[10][2];
};

template <int MB>                        // MB = 1: one warp does both duties, scalar steps first
struct CFSharedT {
    CHStep tab[3][QW_CHUNK];             // chunk c in tab[c % 3]
    CSMap map[MB][QW_CHUNK];             // chunk c in map[c & (MB - 1)]
    double2 sums[2][QW_CHUNK][4];        // chunk c in [c & 1]: (S(y), S(z4), S(z3), S(z2)) / N per step
    double2 sig[2][QW_CHUNK][4];         // sigma_4 .. sigma_1 per step (owner thread)
    double red[2 * 32];
};
using CFShared = CFSharedT<2>;

// one step of the subsystem in complex arithmetic (the operations of the vector code on the
// marked vertex): the ten outputs listed at CSMap for the input (yw, Sy)
__device__ __forceinline__ void cmap_eval(const double *rho, const double2 *c0, const double2 *c1,
                                          const double n, const double invn, const double2 yw,
                                          const double2 Sy, double2 *out)
{
    const double2 m1 = make_double2(fma(n, yw.x, -Sy.x), fma(n, yw.y, -Sy.y));      // (Ly)_w
    double2 g[4];
#pragma unroll
    for (int m = 0; m < 4; ++m) g[m] = cmul2_acc(c1[m], m1, cmul2(c0[m], yw));
    out[0] = make_double2(Sy.x * invn, Sy.y * invn);
    double2 v = m1;
#pragma unroll
    for (int m = 0; m < 3; ++m) {                       // levels 4, 3, 2
        const double2 S = make_double2(Sy.x + g[m].x, Sy.y + g[m].y);
        out[1 + m] = make_double2(S.x * invn, S.y * invn);
        const double2 z = make_double2(fma(rho[3 - m], v.y, yw.x) + g[m].x,
                                       fma(-rho[3 - m], v.x, yw.y) + g[m].y);
        v = make_double2(fma(n, z.x, -S.x), fma(n, z.y, -S.y));
    }
#pragma unroll
    for (int m = 0; m < 4; ++m) out[4 + m] = g[m];
    out[8] = make_double2(fma(rho[0], v.y, yw.x) + g[3].x, fma(-rho[0], v.x, yw.y) + g[3].y);
    out[9] = make_double2(Sy.x + g[3].x, Sy.y + g[3].y);
}

// executed by ONE warp: step-size ratios and subsystem maps of `cnt` steps starting at step n0
__device__ __forceinline__ void horner_fill_complete(CHStep *tab, CSMap *map, const QwPoint &pt,
                                                     const int64_t n0, const int cnt,
                                                     const int schedule, const int gamma_on,
                                                     const double n, const int lane)
{
    for (int i = lane; i < cnt; i += 32) {
        const double t0 = (double)(n0 + i) * pt.h;
        const double2 ad1 = qw_operator_scalars(t0 / pt.T, pt.gamma, schedule, gamma_on);
        const double2 ad2 = qw_operator_scalars((t0 + 0.5 * pt.h) / pt.T, pt.gamma, schedule,
                                                gamma_on);
        const double2 ad4 = qw_operator_scalars((t0 + pt.h) / pt.T, pt.gamma, schedule, gamma_on);
        HornerCoef hc;
        horner_coefficients(hc, pt.h, ad1, ad2, ad4, n - 1.0, n * (n - 1.0));
        // basis (y_w, Ly_w): L^2y_w = N Ly_w, L^3y_w = N^2 Ly_w
        double2 *c = hc.c;
        c[4] = axpy2(n, c[5], c[4]);
        c[7] = axpy2(n * n, c[9], axpy2(n, c[8], c[7]));
        CHStep &o = tab[i];
#pragma unroll
        for (int k = 0; k < 4; ++k) o.rho[k] = hc.rho[k];
        const double2 zero = make_double2(0.0, 0.0), one = make_double2(1.0, 0.0);
        const double2 c0[4] = {c[0], c[1], c[3], c[6]};       // coefficient of y_w in sigma_4..1
        const double2 c1[4] = {zero, c[2], c[4], c[7]};       // coefficient of Ly_w
        double2 qa[10], qb[10];
        const double invn = 1.0 / n;
        cmap_eval(hc.rho, c0, c1, n, invn, one, zero, qa);
        cmap_eval(hc.rho, c0, c1, n, invn, zero, one, qb);
        CSMap &mp = map[i];
#pragma unroll
        for (int k = 0; k < 10; ++k) { mp.q[k][0] = qa[k]; mp.q[k][1] = qb[k]; }
    }
}

// `cnt` steps of the subsystem.  x0..x3 = (re, im) y_w, (re, im) S(y), the same in every lane.
// Lane l < 20 evaluates component l & 1 of output l >> 1; lanes 0..15 store theirs (the level
// means and the corrections of the step), lanes 16..19 hold the next state.
__device__ __forceinline__ void scalar_steps(const CSMap *map, const int cnt, double &x0,
                                             double &x1, double &x2, double &x3,
                                             double2 (*sums)[4], double2 (*sig)[4], const int lane)
{
    const unsigned full = 0xffffffffu;
    const int o = lane < 20 ? (lane >> 1) : 9, c = lane & 1;
#pragma unroll 4
    for (int sidx = 0; sidx < cnt; ++sidx) {
        const double2 P = map[sidx].q[o][0], Q = map[sidx].q[o][1];
        // re: P.x x0 - P.y x1 + Q.x x2 - Q.y x3;  im: P.y x0 + P.x x1 + Q.y x2 + Q.x x3
        const double a0 = c ? P.y : P.x, a1 = c ? P.x : -P.y;
        const double a2 = c ? Q.y : Q.x, a3 = c ? Q.x : -Q.y;
        const double r = fma(a0, x0, a1 * x1) + fma(a2, x2, a3 * x3);
        if (lane < 8) reinterpret_cast<double *>(sums[sidx])[lane] = r;
        else if (lane < 16) reinterpret_cast<double *>(sig[sidx])[lane - 8] = r;
        x0 = __shfl_sync(full, r, 16); x1 = __shfl_sync(full, r, 17);
        x2 = __shfl_sync(full, r, 18); x3 = __shfl_sync(full, r, 19);
    }
}

// one level on all R sites of a thread: out = y - i rho (N x - S), evaluated as
// y - i (rho N)(x - S/N): a two-operand DADD and a DFMA instead of two three-operand DFMAs
// (the FP64 pipe is limited by register-operand delivery: 0.43 against 0.38 warp-instructions
// per clock and sub-partition in tools/fp64_mix_probe.cu, patterns H3 / H).
// rn = rho N in a working thread, 0 in an idle one (its sites stay at zero); rng: the same for
// the last slot -- zero in a short thread, whose last slot is a ghost
template <int R, int LEVEL, bool RAGGED>
__device__ __forceinline__ void clevel(HState<R> &s, const double rn, const double rng,
                                       const double2 Sb)
{
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const bool g = RAGGED && r == R - 1;
        const double dr_ = HX_R(r) - Sb.x;
        const double di_ = HX_I(r) - Sb.y;
        HO_R(r) = fma(g ? rng : rn, di_, s.yr[r]);
        HO_I(r) = fma(-(g ? rng : rn), dr_, s.yi[r]);
    }
}

// Vector work of chunk c (cnt steps).  No barrier inside.
template <int R, bool OWN, bool RAGGED, typename SH>
__device__ __forceinline__ void complete_chunk(HState<R> &s, SH &sh, const int64_t c,
                                               const int cnt, const double nd, const double msk,
                                               const bool full)
{
    const CHStep *tab = sh.tab[c % 3];
    const int cur = (int)(c & 1);
    const double ndg = full ? nd : 0.0;                    // ghost slot stays at zero
    for (int sidx = 0; sidx < cnt; ++sidx) {
        const CHStep &hs = tab[sidx];
        // all scalars of the step up front: nothing inside the step waits for shared memory
        double2 Sl[4];
        double rl[4], rg[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            Sl[k] = sh.sums[cur][sidx][k];
            rl[k] = hs.rho[k] * nd;                        // nd = 0 in an idle thread
            rg[k] = hs.rho[k] * ndg;
        }
#define QW_CH_LEVEL(LEVEL)                                                                  \
        {                                                                                   \
            clevel<R, LEVEL, RAGGED>(s, rl[LEVEL - 1], rg[LEVEL - 1], Sl[4 - LEVEL]);       \
            if (OWN) {                                                                      \
                const double2 sg = sh.sig[cur][sidx][4 - LEVEL];                            \
                if (LEVEL > 1) {                                                            \
                    s.zr[0] = fma(msk, sg.x, s.zr[0]); s.zi[0] = fma(msk, sg.y, s.zi[0]);   \
                } else {                                                                    \
                    s.yr[0] = fma(msk, sg.x, s.yr[0]); s.yi[0] = fma(msk, sg.y, s.yi[0]);   \
                }                                                                           \
            }                                                                               \
        }
        QW_CH_LEVEL(4)
        QW_CH_LEVEL(3)
        QW_CH_LEVEL(2)
        QW_CH_LEVEL(1)
#undef QW_CH_LEVEL
    }
}

// Which part of a point a block integrates.  Unsplit: everything.  Split (complete graph above
// 4096 vertices): block b of `split` holds sites [off, off + nb) of the relabelled graph (site 0 =
// the marked vertex), nb = n / split rounded up for the first n % split blocks, dealt out to its
// threads as R sites to the first nfull and R - 1 to the others.
struct CPart {
    int64_t point;               // launch-order index of the point
    int64_t off;                 // first site of the block
    int32_t b, nact, nfull;
};

template <int R, bool SPLIT>
__device__ __forceinline__ CPart complete_part(const QwParams &prm)
{
    CPart c;
    if (!SPLIT) {
        c.point = blockIdx.x; c.off = 0; c.b = 0; c.nact = prm.nact; c.nfull = prm.nfull;
        return c;
    }
    const int64_t split = prm.split;
    c.point = blockIdx.x / split;
    c.b = (int)(blockIdx.x - c.point * split);
    const int64_t base = prm.n / split, extra = prm.n - base * split;
    const int64_t nb = base + (c.b < extra ? 1 : 0);
    c.off = c.b * base + (c.b < extra ? c.b : extra);
    c.nact = (int)((nb + R - 1) / R);
    c.nfull = (int)(nb - (int64_t)c.nact * (R - 1));
    return c;
}

// p, norm and psi(T) of a point from the workers' registers
template <int R, bool SPLIT, typename SH>
__device__ __forceinline__ void complete_epilogue(const HState<R> &s, SH &sh,
                                                  const QwParams &prm, const QwPoint &pt,
                                                  const CPart &cp, const int t, const bool active,
                                                  const bool full)
{
    double nrm = 0.0;
#pragma unroll
    for (int r = 0; r < R; ++r) nrm = fma(s.yr[r], s.yr[r], fma(s.yi[r], s.yi[r], nrm));
    const double pw = (t == 0 && cp.b == 0) ? fma(s.yr[0], s.yr[0], s.yi[0] * s.yi[0]) : 0.0;
    __syncthreads();
    const double2 tot = hblock_sum(make_double2(nrm, pw), sh.red);
    if (t == 0) {
        if (!SPLIT) {
            prm.p[pt.q] = tot.y / tot.x;
            if (prm.norm) prm.norm[pt.q] = tot.x;
        } else {
            // partial sums meet in global memory; the last block of the point to arrive adds them
            // in block order (deterministic) and writes the results
            double *part = prm.split_part + (size_t)cp.point * (size_t)prm.split * 2;
            part[2 * cp.b] = tot.x;
            part[2 * cp.b + 1] = tot.y;
            __threadfence();
            const unsigned prev = atomicAdd(&prm.split_done[cp.point], 1u);
            if (prev == (unsigned)prm.split - 1u) {
                __threadfence();
                const volatile double *vp = part;
                double nsum = 0.0, psum = 0.0;
                for (int b = 0; b < prm.split; ++b) { nsum += vp[2 * b]; psum += vp[2 * b + 1]; }
                prm.p[pt.q] = psum / nsum;
                if (prm.norm) prm.norm[pt.q] = nsum;
            }
        }
    }
    if (prm.psi && active) {      // undo the relabelling: slot r of thread t -> site w + off + off(t) + r
        double2 *out = prm.psi + (size_t)pt.q * (size_t)prm.n;
        const int64_t off_t = cp.off + (int64_t)t * (R - 1) + (t < cp.nfull ? t : cp.nfull);
#pragma unroll
        for (int r = 0; r < R; ++r) {
            if (r == R - 1 && !full) continue;                    // ghost
            int64_t j = off_t + r + prm.target;
            if (j >= prm.n) j -= prm.n;
            out[j] = make_double2(s.yr[r], s.yi[r]);
        }
    }
}

// blockDim = worker threads (n / R rounded up to warps) + 32 AUX.  AUX = 1: the last warp is the
// scalar warp and refills the tables too.  AUX = 2: the last warp only refills the tables (the
// "table warp"), the one before it is the scalar warp.
// Warps go to the four SM sub-partitions by slot number and the FP64 pipe is per sub-partition:
// with eight worker warps and both duties in warp 8 (level-by-level scalar step), sub-partition
// 0 issued 587 FP64 instructions per step against 520 on the others and everybody waited for
// it at the chunk barrier (18 % of the warp samples).
template <int R, int MAXNT, int MINB, int AUX, bool RAGGED, bool SPLIT = false>
__global__ void __launch_bounds__(MAXNT + 32 * AUX, MINB) rk4_complete_horner(const QwParams prm)
{
    static_assert(!SPLIT || RAGGED, "a block of a split point holds an arbitrary number of sites");
    __shared__ CFShared sh;
    const CPart cp = complete_part<R, SPLIT>(prm);
    const int t = threadIdx.x, nact = cp.nact;
    const int lane = t & 31, warp = t >> 5, fwarp = (blockDim.x >> 5) - 1;
    const int swarp = fwarp - (AUX - 1);
    const QwPoint pt = qw_point(prm, cp.point);
    const bool active = t < nact;
    const bool full = !RAGGED || t < cp.nfull;         // R sites; else R-1 and a ghost slot
    const double nfull = (double)prm.n;
    const double nd = active ? nfull : 0.0;

    HState<R> s;
    const double amp = active ? 1.0 / sqrt(nfull) : 0.0;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        s.yr[r] = (r == R - 1 && !full) ? 0.0 : amp; s.yi[r] = 0.0;
        s.zr[r] = 0.0; s.zi[r] = 0.0;
    }
    const double msk = (t == 0 && cp.b == 0) ? 1.0 : 0.0;
    const int64_t nchunks = (pt.steps + QW_CHUNK - 1) / QW_CHUNK;
#define QW_CNT(C) ((int)((pt.steps - (C) * QW_CHUNK) < QW_CHUNK ? (pt.steps - (C) * QW_CHUNK) : QW_CHUNK))

    double x0 = 1.0 / sqrt(nfull), x1 = 0.0, x2 = 0.0, x3 = 0.0;    // y_w, S(y) (scalar warp)
    if (nchunks > 0) {
        // prologue: S(y) by direct summation (a split point: n times the uniform amplitude, the
        // same in every block); tables of chunks 0 and 1, scalars of chunk 0
        double a = 0.0, b = 0.0;
#pragma unroll
        for (int r = 0; r < R; ++r) { a += s.yr[r]; b += s.yi[r]; }
        double2 Sy = hblock_sum(make_double2(a, b), sh.red);
        if (SPLIT) Sy = make_double2(nfull * (1.0 / sqrt(nfull)), 0.0);
        x2 = Sy.x; x3 = Sy.y;
        if (AUX == 2 && warp == fwarp && nchunks > 1)
            horner_fill_complete(sh.tab[1], sh.map[1], pt, QW_CHUNK, QW_CNT(1), prm.schedule,
                                 prm.gamma_on, nfull, lane);
        if (warp == swarp) {
            horner_fill_complete(sh.tab[0], sh.map[0], pt, 0, QW_CNT(0), prm.schedule,
                                 prm.gamma_on, nfull, lane);
            if (AUX == 1 && nchunks > 1)
                horner_fill_complete(sh.tab[1], sh.map[1], pt, QW_CHUNK, QW_CNT(1), prm.schedule,
                                     prm.gamma_on, nfull, lane);
            __syncwarp();
            scalar_steps(sh.map[0], QW_CNT(0), x0, x1, x2, x3, sh.sums[0], sh.sig[0], lane);
        }
    }
    for (int64_t c = 0; c < nchunks; ++c) {
        __syncthreads();
        if (warp >= swarp) {
            // two chunks ahead: tables; one chunk ahead: the (y_w, S) subsystem
            if (warp == fwarp && c + 2 < nchunks)
                horner_fill_complete(sh.tab[(c + 2) % 3], sh.map[c & 1], pt, (c + 2) * QW_CHUNK,
                                     QW_CNT(c + 2), prm.schedule, prm.gamma_on, nfull, lane);
            if (warp == swarp && c + 1 < nchunks) {
                const int nxt = (int)((c + 1) & 1);
                scalar_steps(sh.map[nxt], QW_CNT(c + 1), x0, x1, x2, x3, sh.sums[nxt],
                             sh.sig[nxt], lane);
            }
        } else if (warp == 0) {
            complete_chunk<R, true, RAGGED>(s, sh, c, QW_CNT(c), nd, msk, full);
        } else {
            complete_chunk<R, false, RAGGED>(s, sh, c, QW_CNT(c), nd, msk, full);
        }
    }
#undef QW_CNT

    complete_epilogue<R, SPLIT>(s, sh, prm, pt, cp, t, active, full);
}

// Blocks of at most four worker warps: no auxiliary warp.  168 registers allow three warps per
// SM sub-partition, and a dedicated scalar warp would hold a third (N <= 1024: up to a half) of
// them while doing 5 FP64 instructions per step.  Here the last worker warp advances the
// subsystem one chunk ahead before its own chunk, and another worker warp refills the tables
// (both out of line, so the register allocation of the vector loop is not affected; the
// worker's state is parked on the stack for the duration of the call, once per 32 steps).
__device__ __noinline__ void fill_complete_call(CHStep *tab, CSMap *map, const QwPoint *pt,
                                                const int64_t n0, const int cnt, const int schedule,
                                                const int gamma_on, const double n, const int lane)
{
    horner_fill_complete(tab, map, *pt, n0, cnt, schedule, gamma_on, n, lane);
}

__device__ __noinline__ void scalar_steps_call(const CSMap *map, const int cnt, double *x,
                                               double2 (*sums)[4], double2 (*sig)[4], const int lane)
{
    double x0 = x[0], x1 = x[1], x2 = x[2], x3 = x[3];
    scalar_steps(map, cnt, x0, x1, x2, x3, sums, sig, lane);
    x[0] = x0; x[1] = x1; x[2] = x2; x[3] = x3;
}

template <int R, int MAXNT, int MINB, int MB, bool RAGGED, bool SPLIT = false>
__global__ void __launch_bounds__(MAXNT, MINB) rk4_complete_folded(const QwParams prm)
{
    static_assert(!SPLIT || RAGGED, "a block of a split point holds an arbitrary number of sites");
    __shared__ CFSharedT<MB> sh;
    const CPart cp = complete_part<R, SPLIT>(prm);
    const int t = threadIdx.x, nact = cp.nact;
    const int lane = t & 31, warp = t >> 5, nw = blockDim.x >> 5;
    const int swarp = nw - 1, fwarp = nw >= 4 ? 1 : 0;      // warp 0 carries the marked vertex
    const QwPoint pt = qw_point(prm, cp.point);
    const bool active = t < nact;
    const bool full = !RAGGED || t < cp.nfull;         // R sites; else R-1 and a ghost slot
    const double nfull = (double)prm.n;
    const double nd = active ? nfull : 0.0;

    HState<R> s;
    const double amp = active ? 1.0 / sqrt(nfull) : 0.0;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        s.yr[r] = (r == R - 1 && !full) ? 0.0 : amp; s.yi[r] = 0.0;
        s.zr[r] = 0.0; s.zi[r] = 0.0;
    }
    const double msk = (t == 0 && cp.b == 0) ? 1.0 : 0.0;
    const int64_t nchunks = (pt.steps + QW_CHUNK - 1) / QW_CHUNK;
#define QW_CNT(C) ((int)((pt.steps - (C) * QW_CHUNK) < QW_CHUNK ? (pt.steps - (C) * QW_CHUNK) : QW_CHUNK))

    double x[4] = {1.0 / sqrt(nfull), 0.0, 0.0, 0.0};       // y_w, S(y) (warp `swarp`)
    if (nchunks > 0) {
        double a = 0.0, b = 0.0;
#pragma unroll
        for (int r = 0; r < R; ++r) { a += s.yr[r]; b += s.yi[r]; }
        double2 Sy = hblock_sum(make_double2(a, b), sh.red);
        if (SPLIT) Sy = make_double2(nfull * (1.0 / sqrt(nfull)), 0.0);
        x[2] = Sy.x; x[3] = Sy.y;
        if (warp == fwarp && fwarp != swarp && nchunks > 1)
            fill_complete_call(sh.tab[1], sh.map[MB - 1], &pt, QW_CHUNK, QW_CNT(1), prm.schedule,
                               prm.gamma_on, nfull, lane);
        if (warp == swarp) {
            fill_complete_call(sh.tab[0], sh.map[0], &pt, 0, QW_CNT(0), prm.schedule,
                               prm.gamma_on, nfull, lane);
            __syncwarp();
            scalar_steps_call(sh.map[0], QW_CNT(0), x, sh.sums[0], sh.sig[0], lane);
            if (fwarp == swarp && nchunks > 1) {
                __syncwarp();
                fill_complete_call(sh.tab[1], sh.map[MB - 1], &pt, QW_CHUNK, QW_CNT(1),
                                   prm.schedule, prm.gamma_on, nfull, lane);
            }
        }
    }
    for (int64_t c = 0; c < nchunks; ++c) {
        __syncthreads();
        // one chunk ahead: the (y_w, S) subsystem; two chunks ahead: tables
        if (warp == swarp && c + 1 < nchunks) {
            const int nxt = (int)((c + 1) & 1);
            scalar_steps_call(sh.map[nxt & (MB - 1)], QW_CNT(c + 1), x, sh.sums[nxt], sh.sig[nxt],
                              lane);
            if (MB == 1) __syncwarp();           // the refill below reuses the only map buffer
        }
        if (warp == fwarp && c + 2 < nchunks)
            fill_complete_call(sh.tab[(c + 2) % 3], sh.map[c & (MB - 1)], &pt, (c + 2) * QW_CHUNK,
                               QW_CNT(c + 2), prm.schedule, prm.gamma_on, nfull, lane);
        if (warp == 0) {
            complete_chunk<R, true, RAGGED>(s, sh, c, QW_CNT(c), nd, msk, full);
        } else {
            complete_chunk<R, false, RAGGED>(s, sh, c, QW_CNT(c), nd, msk, full);
        }
    }
#undef QW_CNT
    complete_epilogue<R, SPLIT>(s, sh, prm, pt, cp, t, active, full);
}

template <typename K>
cudaError_t hlaunch(K kernel, const QwParams &prm, int threads, size_t smem, cudaStream_t st,
                    qw_plan *plan)
{
    if (smem + sizeof(HornerShared<QW_HCHUNK>) > 48 * 1024) {     // static + dynamic above the default cap
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)smem);
        if (e != cudaSuccess) return e;
    }
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
    const int64_t blocks = prm.n_points * (prm.split > 1 ? prm.split : 1);
    kernel<<<(unsigned)blocks, threads, smem, st>>>(prm);
    return cudaGetLastError();
}

}  // namespace

// Sites per thread of the nested-form kernels, or 0 when n does not fit them: R in {16, 8}
// (the window w-3..w+3 has to sit in one thread), any n from 256 to 4096.
// QW_FLAG_MIRROR on rings that need more than one warp per point (above 1086 sites): the
// block holds the half ring n/2 + 1
static bool horner_mirror(const QwParams &prm)
{
    return (prm.flags & QW_FLAG_MIRROR) && prm.topology == QW_TOPO_CYCLE && prm.n > 1086;
}

// Complete graph above 4096 vertices: blocks per point (0: not split).  The nested form needs
// nothing from the other sites but the tabulated (y_w, S) subsystem, which every block advances
// for itself, so a point splits over any number of blocks without communication; the partial
// norms meet at the end (complete_epilogue).
// Blocks of two to four warps (1024 .. 2048 sites; three to six of them per SM) hide one
// another's barriers better than one block of up to 4096 sites per SM.  Among those block counts
// the cheapest wins: cost = blocks x warps per block x the measured cost of a warp in a block of
// that many warps (site-updates/s per filled lane, tools/ab_large_n.py with 2368 points: four
// warps 9.2e11, two 8.7e11, three 7.8e11); ties: the fewest blocks.  Measured against blocks of
// 2049 .. 4096 sites (QW_FLAG_OCC4 keeps that rule for A/B): N = 5000 6.1e11 -> 8.5e11 (5 x 1000
// sites in two full warps instead of 2 x 2500 in five), 8192 8.7e11 -> 9.2e11, 12289 6.6e11 ->
// 8.1e11, 16384 8.7e11 -> 9.2e11, 40000 8.5e11 -> 9.0e11 (profiles/r2_ab_complete_split_rule.txt).
int qw_horner_split(const QwParams &prm)
{
    if (prm.topology != QW_TOPO_COMPLETE || prm.n <= 4096) return 0;
    if (prm.flags & (QW_FLAG_EXACT_SUM | QW_FLAG_NO_HORNER | QW_FLAG_CYCLE_LEGACY |
                     QW_FLAG_COMPLETE_FREE | QW_FLAG_FORCE_GENERIC))
        return 0;
    int64_t split = (prm.n + 4095) / 4096;
    if (!(prm.flags & QW_FLAG_OCC4)) {
        static const int64_t warp_cost[5] = {0, 106, 106, 118, 100};
        const int64_t lo = (prm.n + 2047) / 2048, hi = (prm.n + 1023) / 1024;
        int64_t best = INT64_MAX;
        for (int64_t s = lo; s <= hi; ++s) {
            const int64_t nb = (prm.n + s - 1) / s;                // the largest block
            const int64_t warps = ((nb + 15) / 16 + 31) / 32;      // 2 .. 4
            const int64_t cost = s * warps * warp_cost[warps];
            if (cost < best) { best = cost; split = s; }
        }
    }
    if (split > 65536 || prm.n_points * split > 0x7fffffffLL) return 0;
    return (int)split;
}

// Cycle graph, 4097 .. 32768 sites: CTAs per cluster (0: not this path).  Portable cluster
// sizes only (<= 8); every CTA holds at most 256 x 16 sites.
int qw_cluster_size(const QwParams &prm)
{
    if (prm.topology != QW_TOPO_CYCLE || prm.n <= 4096 || prm.n > 8 * 4096) return 0;
    if (prm.flags & (QW_FLAG_NO_HORNER | QW_FLAG_CYCLE_LEGACY | QW_FLAG_CYCLE_SPLIT |
                     QW_FLAG_FORCE_GENERIC | QW_FLAG_FORCE_STREAM | QW_FLAG_MIRROR))
        return 0;
    // Measured (tools/ab_large_n.py): two CTAs (up to 8192 sites) tie with the bulk-copy streaming
    // kernel at N = 8192 (4.8e11 site-updates/s) and win below (N = 5000: 4.3e11 against 2.6e11);
    // larger clusters leave SMs idle (3: 148 = 49 x 3 + 1; 8: at most two per GPC) and lose to
    // streaming with 8 steps per pass (N = 16384: 4.3e11 against 5.3e11) -- unless the ring is not
    // a multiple of 128 sites, where streaming falls back to its cp.async kernel.
    if (prm.n > 8192 && prm.n % 128 == 0 && !(prm.flags & QW_FLAG_FORCE_CLUSTER)) return 0;
    int64_t c = (prm.n + 4095) / 4096;
    // Above 8192 sites: small CTAs (at most 2050 sites, 160 threads, two per SM) where the cluster
    // stays portable -- N = 12289: 4.43e11 with six small CTAs against 3.65e11 with four big ones;
    // up to 8192 sites the big CTAs win (N = 8192: 4.98e11 against 4.65e11, N = 5000: 4.48e11
    // against 3.69e11).  QW_FLAG_OCC4: always one CTA of up to 4096 sites per SM (A/B).
    const int64_t c_small = (prm.n + 2049) / 2050;
    if (prm.n > 8192 && c_small <= 8 && !(prm.flags & QW_FLAG_OCC4)) c = c_small;
    if (prm.n_points * c > 0x7fffffffLL) return 0;
    return (int)c;
}

cudaError_t qw_launch_cluster(QwParams prm, cudaStream_t st, qw_plan *plan)
{
    const int c = qw_cluster_size(prm);
    if (c == 0) return cudaErrorInvalidConfiguration;
    constexpr int R = 16;
    const int64_t nb = (prm.n + c - 1) / c;                        // the largest CTA
    const int nact = (int)((nb - 2 + R - 1) / R);                  // two sites live in the comm warp
    const int threads = ((nact + 31) / 32) * 32 + 32;              // + the comm warp
    const size_t smem = (size_t)threads * 4 * sizeof(double2);
    // n_r - 2 compute sites: some threads hold R - 1
    auto kernel = threads <= 160 ? rk4_cycle_cluster<R, true, 160, 2> : rk4_cycle_cluster<R, true, 288, 1>;
    if (plan) {
        cudaFuncAttributes fa;
        cudaError_t e = cudaFuncGetAttributes(&fa, kernel);
        if (e != cudaSuccess) return e;
        plan->path = 12;
        plan->sites_per_thread = R;
        plan->threads_per_block = threads;
        plan->blocks_per_sm = threads <= 160 ? 2 : 1;
        plan->regs_per_thread = fa.numRegs;
        plan->smem_bytes = (int)(fa.sharedSizeBytes + smem);
        plan->steps_per_pass = c;                                  // CTAs per cluster
        return cudaSuccess;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(prm.n_points * c), 1, 1);
    cfg.blockDim = dim3((unsigned)threads, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)c;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, prm);
}

int qw_horner_sites_per_thread(const QwParams &prm)
{
    if (qw_horner_split(prm) > 1) return 16;
    if (horner_mirror(prm)) {
        // the half ring has n/2 + 1 sites: 17 sites per thread keep the "+ 1" from costing a
        // whole extra warp (n = 2048: 1025 = 61 threads x 17 instead of 65 x 16).  Where 17-site
        // threads leave much of the last warp empty, or a block of five or six warps alone on its
        // SM, 13 or 9 sites per thread do better (n = 1100: 551 sites = 62 threads of 9, six
        // blocks per SM, instead of 33 of 17 in two warps: 5.5e11 -> 9.2e11 full-ring
        // site-updates/s).  The choice maximises filled slots x a weight per R x a weight for the
        // warps resident per SM, all three measured (tools/ab_large_n.py, profiles/
        // r2_ab_mirror_block_r.txt: per filled slot 9 sites run at 0.95, 13 at 1.05, 17 at 1.06;
        // five warps per SM at 0.6, six at 0.8, seven at 0.92 of eight and more; blocks of three
        // warps at 0.9); ties to the larger R.  QW_FLAG_NO_PACK: 17 only (A/B).
        const int64_t m = prm.n / 2 + 1;
        if ((m + 16) / 17 > 256) return 0;
        if (prm.flags & QW_FLAG_NO_PACK) return 17;
        const int cand[3][2] = {{17, 106}, {13, 105}, {9, 95}};
        int best = 17;
        int64_t best_score = -1;
        for (int i = 0; i < 3; ++i) {
            const int R = cand[i][0];
            const int64_t nact = (m + R - 1) / R, warps = (nact + 31) / 32;
            if (nact > 256) continue;
            // blocks per SM of the instantiations in qw_launch_horner
            const int64_t blocks = warps <= 2 ? (R == 17 ? 4 : 6) : warps <= 4 ? (R == 17 ? 2 : 3) : 1;
            const int64_t resident = blocks * warps;
            const int64_t occ = resident >= 8 ? 100 : resident == 7 ? 92 : resident == 6 ? 80 : 60;
            const int64_t shape = warps == 3 ? 90 : 100;    // blocks of three warps run slower
            const int64_t score = m * cand[i][1] * occ * shape / (R * warps * 32);
            if (score > best_score) { best_score = score; best = R; }
        }
        return best;
    }
    const int64_t n = prm.n;
    if (prm.topology != QW_TOPO_CYCLE && (prm.flags & QW_FLAG_EXACT_SUM)) return 0;
    // n = nact (R - 1) + nfull with 1 <= nfull <= nact needs n >= R^2; at least one full warp
    if (!(prm.flags & QW_FLAG_R8) && n >= 512 && (n + 15) / 16 <= 256) return 16;
    if (n >= 256 && (n + 7) / 8 <= 512) return 8;
    return 0;
}

cudaError_t qw_launch_horner(QwParams prm, cudaStream_t st, qw_plan *plan)
{
    const int R = qw_horner_sites_per_thread(prm);
    if (R == 0) return cudaErrorInvalidConfiguration;
    const bool mirror = horner_mirror(prm);
    const int64_t sites = mirror ? prm.n / 2 + 1 : prm.n;
    prm.nact = (int)((sites + R - 1) / R);
    prm.nfull = (int)(sites - (int64_t)prm.nact * (R - 1));       // threads with R sites
    const int threads = ((prm.nact + 31) / 32) * 32;
    const bool cyc = prm.topology == QW_TOPO_CYCLE;
    // cycle kernels: mailbox [2][2][MAXNT] of the instantiation chosen below (set by QW_MB)
    size_t smem = 0;
#define QW_MB(MAXNT) smem = (size_t)(MAXNT) * 4 * sizeof(double2)
    if (plan) {
        plan->path = cyc ? 5 : 6;
        plan->sites_per_thread = R;
    }
    if (const int split = qw_horner_split(prm)) {
        // blocks of a split point: geometry per block inside the kernel (complete_part)
        prm.split = split;
        const int64_t nb = (prm.n + split - 1) / split;            // the largest block
        const int bnact = (int)((nb + R - 1) / R);
        const int bthreads = ((bnact + 31) / 32) * 32;
        if (plan) {
            plan->path = 6;
            plan->sites_per_thread = R;
            plan->steps_per_pass = split;                          // blocks per point
        }
        if (bthreads <= 192)       // up to 3072 sites per block: duties folded, two blocks per SM
            return hlaunch(rk4_complete_folded<16, 192, 2, 2, true, true>, prm, bthreads, 0, st, plan);
        return hlaunch(rk4_complete_horner<16, 256, 1, 2, true, true>, prm, bthreads + 64, 0, st, plan);
    }
    const bool ragged = prm.nfull != prm.nact;
#define QW_HL(KERNEL, ...)                                                                  \
    return ragged ? hlaunch(KERNEL<__VA_ARGS__, true>, prm, bthreads, smem, st, plan)      \
                  : hlaunch(KERNEL<__VA_ARGS__, false>, prm, bthreads, smem, st, plan)
    if (!cyc) {
        if (R == 16 && threads > 192 && !(prm.flags & QW_FLAG_OCC4)) {
            const int bthreads = threads + 64;    // + the scalar warp and the table warp
            QW_HL(rk4_complete_horner, 16, 256, 1, 2);
        }
        if (R == 16 && threads <= 192 && !(prm.flags & QW_FLAG_OCC4)) {
            const int bthreads = threads;         // duties folded into the worker warps
            if (threads <= 32) { QW_HL(rk4_complete_folded, 16, 32, 12, 1); }
            if (threads <= 64) { QW_HL(rk4_complete_folded, 16, 64, 6, 2); }
            if (threads <= 128) { QW_HL(rk4_complete_folded, 16, 128, 3, 2); }
            QW_HL(rk4_complete_folded, 16, 192, 2, 2);        // 5 or 6 warps: two blocks per SM
        }
        if (R == 8 && threads <= 64 && !(prm.flags & QW_FLAG_OCC4)) {      // 256 <= n < 512
            const int bthreads = threads;
            if (threads <= 32) { QW_HL(rk4_complete_folded, 8, 32, 12, 1); }
            QW_HL(rk4_complete_folded, 8, 64, 7, 2);         // 7 blocks: shared memory
        }
        const int bthreads = threads + 32;        // + the scalar warp (both duties)
        if (R == 16) {
            if (threads <= 64) { QW_HL(rk4_complete_horner, 16, 64, 4, 1); }
            if (threads <= 128) { QW_HL(rk4_complete_horner, 16, 128, 2, 1); }
            QW_HL(rk4_complete_horner, 16, 256, 1, 1);        // QW_FLAG_OCC4: A/B only
        }
        if (threads <= 128) { QW_HL(rk4_complete_horner, 8, 128, 3, 1); }
        if (threads <= 256) { QW_HL(rk4_complete_horner, 8, 256, 1, 1); }
        QW_HL(rk4_complete_horner, 8, 512, 1, 1);
    }
    const int bthreads = threads;
    if (mirror) {                                  // R = 17: up to 255 registers
#define QW_HM(...) return hlaunch(rk4_cycle_horner<__VA_ARGS__, true, true>, prm, bthreads, smem, st, plan)
        if (R == 9) {
            if (threads <= 64) { QW_MB(64); QW_HM(9, 64, 6); }
            if (threads <= 128) { QW_MB(128); QW_HM(9, 128, 3); }
            QW_MB(256);
            QW_HM(9, 256, 1);
        }
        if (R == 13) {
            if (threads <= 64) { QW_MB(64); QW_HM(13, 64, 6); }
            if (threads <= 128) { QW_MB(128); QW_HM(13, 128, 3); }
            QW_MB(256);
            QW_HM(13, 256, 1);
        }
        if (threads <= 64) { QW_MB(64); QW_HM(17, 64, 4); }
        if (threads <= 128) { QW_MB(128); QW_HM(17, 128, 2); }
        QW_MB(256);
        QW_HM(17, 256, 1);
#undef QW_HM
    }
    if (R == 16) {
        // register file = 16 K per SM sub-partition: 2 / 3 warps there allow 255 / 168 registers
        if (threads <= 64 && (prm.flags & QW_FLAG_OCC4)) { QW_MB(64); QW_HL(rk4_cycle_horner, 16, 64, 4); }
        if (threads <= 64) { QW_MB(64); QW_HL(rk4_cycle_horner, 16, 64, QW_OCC_R16_64); }
        // 5 or 6 warps: two blocks per SM at 168 registers (one block at 200: QW_FLAG_OCC4, A/B)
        if (threads > 128 && threads <= 192 && !(prm.flags & QW_FLAG_OCC4)) {
            QW_MB(192);
            QW_HL(rk4_cycle_horner, 16, 192, 2);
        }
        if (threads <= 128) { QW_MB(128); QW_HL(rk4_cycle_horner, 16, 128, 3); }
        QW_MB(256);
        QW_HL(rk4_cycle_horner, 16, 256, 1);
    }
    if (threads <= 64) { QW_MB(64); QW_HL(rk4_cycle_horner, 8, 64, 8); }
    if (threads <= 128) { QW_MB(128); QW_HL(rk4_cycle_horner, 8, 128, 4); }
    if (threads <= 256) { QW_MB(256); QW_HL(rk4_cycle_horner, 8, 256, 2); }
    QW_MB(512);
    QW_HL(rk4_cycle_horner, 8, 512, 1);
#undef QW_MB
#undef QW_HL
}
