// Shared pieces of the nested (Horner) form of the RK4 step: step coefficients with the
// tabulated oracle corrections, register state, one level on a thread's sites.  Used by the
// resident kernels (qw_horner.cu) and by the streaming kernel (qw_stream.cu).  See the header
// of qw_horner.cu and DESIGN.md section 2.1 for the algebra.
#pragma once

#include "qw_common.cuh"

namespace {

constexpr int HW0 = 3;               // slot of the marked vertex in thread 0

struct HornerCoef {                  // coefficients of one RK4 step
    double rho[4];                   // r1 (outermost) .. r4 (innermost)
    double2 c[10];                   // sigma_4: c[0]; sigma_3: c[1..2]; sigma_2: c[3..5];
                                     // sigma_1: c[6..9]  (basis y_w, Ly_w, L^2y_w, L^3y_w)
};

__device__ __forceinline__ double2 mul_mi(const double s, const double2 v)   // (-i s) v
{
    return make_double2(s * v.y, -s * v.x);
}

__device__ __forceinline__ double2 axpy2(const double s, const double2 v, const double2 acc)
{
    return make_double2(fma(s, v.x, acc.x), fma(s, v.y, acc.y));
}

__device__ __forceinline__ double safe_inv(const double x) { return x != 0.0 ? 1.0 / x : 0.0; }

// Coefficients of one step.  ad1, ad2, ad4 = (a, d) at t_n, t_n + h/2, t_n + h;
// lww = L_ww, l2ww = (L^2)_ww.
__device__ void horner_coefficients(HornerCoef &o, const double h, const double2 ad1,
                                    const double2 ad2, const double2 ad4, const double lww,
                                    const double l2ww)
{
    const double A1 = 0.5 * h * ad1.x, A2 = 0.5 * h * ad2.x, A3 = h * ad2.x;
    const double B1 = h / 6.0 * ad1.x, B2 = h / 3.0 * ad2.x, B4 = h / 6.0 * ad4.x;
    const double DA1 = 0.5 * h * ad1.y, DA2 = 0.5 * h * ad2.y, DA3 = h * ad2.y;
    const double DB1 = h / 6.0 * ad1.y, DB2 = h / 3.0 * ad2.y, DB4 = h / 6.0 * ad4.y;
    const double b1 = B1 + 2.0 * B2 + B4;
    const double b2 = B2 * A1 + B2 * A2 + B4 * A3;
    const double b3 = B2 * A2 * A1 + B4 * A3 * A2;
    const double b4 = B4 * A3 * A2 * A1;
    const double i1 = safe_inv(b1), i2 = safe_inv(b2), i3 = safe_inv(b3);
    o.rho[0] = b1; o.rho[1] = b2 * i1; o.rho[2] = b3 * i2; o.rho[3] = b4 * i3;

    const double2 zero = make_double2(0.0, 0.0), one = make_double2(1.0, 0.0);
    // q_s = true stage value at the marked vertex, al_s = -i c_s d_s q_s, as coefficient
    // vectors over (y_w, Ly_w, L^2y_w, L^3y_w)
    double2 q2[4], q3[4], q4[4], al1[4], al2[4], al3[4];
    al1[0] = make_double2(0.0, -DA1); al1[1] = zero; al1[2] = zero; al1[3] = zero;
    q2[0] = make_double2(1.0, -DA1); q2[1] = make_double2(0.0, -A1); q2[2] = zero; q2[3] = zero;
    const double2 u3[4] = {one, make_double2(0.0, -A2), make_double2(-A2 * A1, 0.0), zero};
    const double2 u4[4] = {one, make_double2(0.0, -A3), make_double2(-A3 * A2, 0.0),
                           make_double2(0.0, A3 * A2 * A1)};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        al2[j] = mul_mi(DA2, q2[j]);
        const double2 t3 = mul_mi(A2 * lww, al1[j]);
        q3[j] = make_double2(u3[j].x + t3.x + al2[j].x, u3[j].y + t3.y + al2[j].y);
        al3[j] = mul_mi(DA3, q3[j]);
        const double2 t4 = mul_mi(A3 * lww, al2[j]);
        q4[j] = make_double2(u4[j].x - A3 * A2 * l2ww * al1[j].x + t4.x + al3[j].x,
                             u4[j].y - A3 * A2 * l2ww * al1[j].y + t4.y + al3[j].y);
    }
    double2 k0[4], k1[4], k2[4], k3[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        double2 s = (j == 0) ? make_double2(DB1, 0.0) : zero;          // DB1 q1
        s = axpy2(DB2, q2[j], s);
        s = axpy2(DB2, q3[j], s);
        s = axpy2(DB4, q4[j], s);
        k0[j] = make_double2(s.y, -s.x);                               // -i s
        k1[j] = axpy2(B4, al3[j], axpy2(B2, al2[j], make_double2(B2 * al1[j].x, B2 * al1[j].y)));
        k2[j] = axpy2(B4 * A3, al2[j], make_double2(B2 * A2 * al1[j].x, B2 * A2 * al1[j].y));
        k3[j] = make_double2(B4 * A3 * A2 * al1[j].x, B4 * A3 * A2 * al1[j].y);
    }
    o.c[0] = make_double2(k3[0].x * i3, k3[0].y * i3);
    o.c[1] = make_double2(k2[0].x * i2, k2[0].y * i2);
    o.c[2] = make_double2(k2[1].x * i2, k2[1].y * i2);
    o.c[3] = make_double2(k1[0].x * i1, k1[0].y * i1);
    o.c[4] = make_double2(k1[1].x * i1, k1[1].y * i1);
    o.c[5] = make_double2(k1[2].x * i1, k1[2].y * i1);
    o.c[6] = k0[0]; o.c[7] = k0[1]; o.c[8] = k0[2]; o.c[9] = k0[3];
}

// Cycle graph: change the basis of the correction coefficients from (y_w, Ly_w, L^2y_w, L^3y_w)
// to (y_w, s1, s2, s3), s_k = y_{w-k} + y_{w+k}:
//   Ly_w = 2y - s1,  L^2y_w = 6y - 4s1 + s2,  L^3y_w = 20y - 15s1 + 6s2 - s3
__device__ __forceinline__ void horner_cycle_sum_basis(double2 *c)
{
    c[1] = axpy2(2.0, c[2], c[1]);
    c[2] = make_double2(-c[2].x, -c[2].y);
    c[3] = axpy2(6.0, c[5], axpy2(2.0, c[4], c[3]));
    c[4] = make_double2(-c[4].x - 4.0 * c[5].x, -c[4].y - 4.0 * c[5].y);
    c[6] = axpy2(20.0, c[9], axpy2(6.0, c[8], axpy2(2.0, c[7], c[6])));
    c[7] = make_double2(-c[7].x - 4.0 * c[8].x - 15.0 * c[9].x,
                        -c[7].y - 4.0 * c[8].y - 15.0 * c[9].y);
    c[8] = axpy2(6.0, c[9], c[8]);
    c[9] = make_double2(-c[9].x, -c[9].y);
}

// Shared-memory form for the cycle kernel.  The four corrections are evaluated by ALL 32 lanes of
// the owner (or helper) warp at once: lane 8q + l forms the two terms that s_q = y_{w-q} + y_{w+q}
// contributes to output l, and two butterfly additions over q finish the dot products; output
// l = 2m + c is component c (0 re, 1 im) of  y_w + sigma_{4-m},  i.e. the identity is folded into
// the coefficients of y_w, so that the owner simply uses the value as the ADDEND of the level's
// axpy at the marked vertex.
// MIRROR SYMMETRY: the state is symmetric about the marked vertex bit for bit -- every caller
// starts from the uniform state (BCB.py:179-180; the C ABI has no other start), H(t) is invariant
// under the reflection and the neighbour sums of the level are commutative -- so s_q = 2 y_{w-q}
// exactly and only the sites w-3 .. w are needed: the factor 2 is folded into the coefficients
// (exact), a lane multiplies ONE window site: 4 FP64 warp-instructions per step (DMUL, DFMA, 2
// DADD) instead of the 22 of the first version (6 sums + a dot product of length 8 in eight lanes
// + 8 masked corrections), which cost the hot kernel 9 % (tools/hot_probe.cu: the owner warp was
// the longer of the two warps of a block at every barrier).  tests/test_gpu_horner.py checks the
// symmetry of psi itself bitwise.
// coef2[lane] = the lane's two coefficients (of Re y_{w-q}, Im y_{w-q}): one conflict-free LDS.128.
struct HornerStepBase {                  // the one-warp forms (every addend from the dot product)
    double rho[4];
    double2 coef2[32];
};
struct HornerStep : HornerStepBase {
    double2 c4;                          // sigma_4 = c4 y_w (the owner forms it itself)
};


// One step's table entry from the coefficients in the sum basis (horner_cycle_sum_basis).
// fold4: the identity is folded into sigma_4's coefficient as well (the one-warp forms take all
// four addends from the dot product; the owner / helper form computes y_w + c4 y_w itself).
__device__ __forceinline__ void horner_step_store(HornerStepBase &o, const HornerCoef &hc,
                                                  const bool fold4)
{
    const double2 *c = hc.c;
#pragma unroll
    for (int k = 0; k < 4; ++k) o.rho[k] = hc.rho[k];
    const int base[4] = {0, 1, 3, 6};              // first coefficient of sigma_4 .. sigma_1
#pragma unroll
    for (int m = 0; m < 4; ++m) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {              // term k multiplies (Re s_k, Im s_k)
            double2 ck = (k <= m) ? c[base[m] + k] : make_double2(0.0, 0.0);
            if (k == 0) {                          // y_w + sigma
                if (m > 0 || fold4) ck.x += 1.0;
            } else {                               // s_k = 2 y_{w-k}
                ck.x *= 2.0;
                ck.y *= 2.0;
            }
            o.coef2[8 * k + 2 * m] = make_double2(ck.x, -ck.y);        // re
            o.coef2[8 * k + 2 * m + 1] = make_double2(ck.y, ck.x);     // im
        }
    }
}

__device__ __forceinline__ double2 cmul2(const double2 c, const double2 v)
{
    return make_double2(fma(c.x, v.x, -c.y * v.y), fma(c.x, v.y, c.y * v.x));
}

__device__ __forceinline__ double2 cmul2_acc(const double2 c, const double2 v, const double2 acc)
{
    return make_double2(fma(c.x, v.x, fma(-c.y, v.y, acc.x)), fma(c.x, v.y, fma(c.y, v.x, acc.y)));
}

// out = y - i rho (2 c - l - r)
__device__ __forceinline__ void hsite(const double lr, const double li, const double cr,
                                      const double ci, const double rr, const double ri,
                                      const double rho, const double y_r, const double y_i,
                                      double &o_r, double &o_i)
{
    const double vr = fma(2.0, cr, -(lr + rr));
    const double vi = fma(2.0, ci, -(li + ri));
    o_r = fma(rho, vi, y_r);
    o_i = fma(-rho, vr, y_i);
}

template <int R>
struct HState {
    double yr[R], yi[R];   // psi_n
    double zr[R], zi[R];   // nested intermediate
};

// LEVEL 4: x = y, out = z.  LEVEL 3, 2: x = z, out = z (in place).  LEVEL 1: x = z, out = y.
#define HX_R(r) ((LEVEL == 4) ? s.yr[r] : s.zr[r])
#define HX_I(r) ((LEVEL == 4) ? s.yi[r] : s.zi[r])
#define HO_R(r) ((LEVEL == 1) ? s.yr[r] : s.zr[r])
#define HO_I(r) ((LEVEL == 1) ? s.yi[r] : s.zi[r])

// 16-byte shared-memory accesses by 32-bit address (the owner / helper code of the resident kernels
// selects its addresses per lane)
__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ double2 lds2(const uint32_t a)
{
    double2 v;
    // "memory": the compiler must not move ordinary shared-memory stores (e.g. the dot products
    // written just before a __syncwarp) across this load
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a) : "memory");
    return v;
}

// predicated store: a branch around a handful of stores cost the owner warp 3 % of its time in
// branch resolution (ncu source page)
__device__ __forceinline__ void sts2_if(const uint32_t a, const double x, const double y,
                                        const int on)
{
    asm volatile("{\n.reg .pred p;\nsetp.ne.s32 p, %3, 0;\n@p st.shared.v2.f64 [%0], {%1, %2};\n}"
                 ::"r"(a), "d"(x), "d"(y), "r"(on) : "memory");
}

__device__ __forceinline__ void sts2(const uint32_t a, const double x, const double y)
{
    asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(a), "d"(x), "d"(y) : "memory");
}

// WS >= 0: the addend of slot WS is (wr, wi) instead of y[WS] -- the owner of the marked vertex
// passes y_w + sigma_LEVEL there (the oracle correction folded into the level's axpy).
// PUBW > 0 (level 1 only): the new values of slots 0 .. PUBW-1 except WS go to pub[slot] as soon
// as they exist, stored by the lanes with `on` set (the owner publishes the window around the
// marked vertex for the next step's corrections).
template <int R, int LEVEL, int WS = -1, int PUBW = 0>
__device__ __forceinline__ void hedges(HState<R> &s, const double rho, const double2 hl,
                                       const double2 hr, double2 &o0, double2 &oL,
                                       const double wr = 0.0, const double wi = 0.0,
                                       const uint32_t pub = 0, const int on = 0)
{
    o0 = make_double2(HX_R(0), HX_I(0));
    oL = make_double2(HX_R(R - 1), HX_I(R - 1));
    hsite(hl.x, hl.y, o0.x, o0.y, HX_R(1), HX_I(1), rho, WS == 0 ? wr : s.yr[0],
          WS == 0 ? wi : s.yi[0], HO_R(0), HO_I(0));
    if (LEVEL == 1 && PUBW > 0 && WS != 0) sts2_if(pub, HO_R(0), HO_I(0), on);
    hsite(HX_R(R - 2), HX_I(R - 2), oL.x, oL.y, hr.x, hr.y, rho, s.yr[R - 1], s.yi[R - 1],
          HO_R(R - 1), HO_I(R - 1));
}

template <int R, int LEVEL, int WS = -1, int PUBW = 0>
__device__ __forceinline__ void hinterior(HState<R> &s, const double rho, const double2 o0,
                                          const double2 oL, const double wr = 0.0,
                                          const double wi = 0.0, const uint32_t pub = 0,
                                          const int on = 0)
{
    double pr = o0.x, pi = o0.y;
#pragma unroll
    for (int r = 1; r < R - 1; ++r) {
        const double cr = HX_R(r), ci = HX_I(r);
        const double nr = (r + 1 < R - 1) ? HX_R(r + 1 < R ? r + 1 : r) : oL.x;
        const double ni = (r + 1 < R - 1) ? HX_I(r + 1 < R ? r + 1 : r) : oL.y;
        hsite(pr, pi, cr, ci, nr, ni, rho, WS == r ? wr : s.yr[r], WS == r ? wi : s.yi[r],
              HO_R(r), HO_I(r));
        if (LEVEL == 1 && r < PUBW && r != WS) sts2_if(pub + 16 * r, HO_R(r), HO_I(r), on);
        pr = cr;
        pi = ci;
    }
}

__device__ __forceinline__ void horner_step_store(HornerStep &o, const HornerCoef &hc,
                                                  const bool fold4)
{
    horner_step_store(static_cast<HornerStepBase &>(o), hc, fold4);
    o.c4 = hc.c[0];
}

// lane 8q + l: the term of y_{w-q} in output l, then the sum over q; every lane ends up with
// output l = lane & 7.  wa: where the lane finds y_{w-q} (horner_window_addr)
__device__ __forceinline__ double horner_window_dot(const uint32_t wa, const HornerStepBase &hs,
                                                    const int lane)
{
    const double2 ya = lds2(wa);
    const double2 cf = hs.coef2[lane];
    double v = fma(cf.y, ya.y, cf.x * ya.x);
    v += __shfl_xor_sync(0xffffffffu, v, 8);
    v += __shfl_xor_sync(0xffffffffu, v, 16);
    return v;
}

// slot of the owner thread -> its shared-memory copy: the marked slot lives in priv[owner lane]
template <bool MIRROR>
__device__ __forceinline__ uint32_t horner_window_addr(const int slot, const double2 *win,
                                                       const double2 *priv)
{
    return slot == (MIRROR ? 0 : HW0) ? smem_u32(priv) : smem_u32(win + slot);
}

}  // namespace
