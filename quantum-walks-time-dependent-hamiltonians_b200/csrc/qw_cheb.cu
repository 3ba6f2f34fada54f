// Exact time-independent propagator psi(t) = exp(-i H t) psi0 for H = a0 L + d0 |w><w| on the
// cycle graph (sm_100a) -- the comparator the reference evaluates with a dense Pade
// scipy.linalg.expm (BCB.py:137-159, O(N^3) per point).
//
// Chebyshev expansion of the propagator over the implicit stencil:
//     exp(-i H t) = e^{-i b t} sum_k (2 - delta_k0) (-i)^k J_k(a t) T_k(Ht),  Ht = (H - b)/a,
// with [b - a, b + a] a Gershgorin/Weyl enclosure of the spectrum.  T_k(Ht) psi0 is REAL for
// the uniform initial state, so a term costs 1 DADD + 3 DFMA per site (two for the
// three-term recurrence, one for the accumulation into Re or Im according to k mod 4), and
// K ~ a t + 12 (a t)^(1/3) terms reach 1e-16.  Same register-resident layout as the RK4
// kernels: one block per (t, gamma) point, R consecutive sites per thread, marked vertex
// relabelled to site 0 of thread 0, edge values through a double-buffered shared mailbox,
// one __syncthreads() per term.  Thread 0 builds J_0..J_K by Miller's backward recurrence.
#include "qw_common.cuh"

namespace {

constexpr int CHEB_KMAX = 4096;        // Bessel table in shared memory (32 KB): a*|t| < ~3800

struct ChebShared {
    double jtab[CHEB_KMAX + 8];
    double red[128];
    double cself, cnb, cdiag, bshift;
    int kterms;
    int status;                        // 0 ok, 1 too many terms
};

// J_0..J_K of z by Miller's algorithm (downward recurrence from M > K, normalised with
// J_0 + 2 sum J_2m = 1).  Sequential; run by one thread.
__device__ void bessel_table(double z, int K, double *jt)
{
    if (!(z > 1e-300)) {
        jt[0] = 1.0;
        for (int k = 1; k <= K; ++k) jt[k] = 0.0;
        return;
    }
    const int M = K + 40 + (int)(0.1 * K);
    const double tz = 2.0 / z;
    double jp = 0.0, jc = 1e-280, sum = 0.0;      // j_{k+1}, j_k at k = M
    for (int k = M; k >= 1; --k) {
        const double jm = fma((double)k * tz, jc, -jp);     // j_{k-1}
        jp = jc;
        jc = jm;
        if (k - 1 <= K) jt[k - 1] = jc;
        if (((k - 1) & 1) == 0 && k - 1 > 0) sum += jc;
        if (fabs(jc) > 1e250) {                   // rescale everything kept so far
            const double sc = 1e-250;
            jc *= sc; jp *= sc; sum *= sc;
            for (int i = k - 1; i <= K; ++i) jt[i] *= sc;
        }
    }
    const double norm = 1.0 / fma(2.0, sum, jc);  // jc = j_0 here
    for (int k = 0; k <= K; ++k) jt[k] *= norm;
}

__device__ __forceinline__ double2 cheb_block_sum(double a, double b, double *red)
{
    return qw_block_sum2(a, b, red);
}

template <int R, int MAXNT, int MINB>
__global__ void __launch_bounds__(MAXNT, MINB) cheb_cycle_resident(const QwParams prm)
{
    extern __shared__ double mbox[];          // [2][2][blockDim.x] doubles: parity, first/last
    __shared__ ChebShared sh;
    const int t = threadIdx.x, nt = blockDim.x, nact = prm.nact;
    const int64_t q = prm.order ? prm.order[blockIdx.x] : (int64_t)blockIdx.x;
    const double time = prm.time_array ? prm.time_array[q % prm.n_time] : prm.T[q];
    const double gamma = prm.time_array ? prm.beta_array[prm.beta_begin + q / prm.n_time]
                                        : prm.gamma[q];
    const bool active = t < nact;
    const int tl = active ? (t == 0 ? nact - 1 : t - 1) : t;
    const int tr = active ? (t == nact - 1 ? 0 : t + 1) : t;

    if (t == 0) {
        // H = a0 L + d0 |w><w|; spectrum of L_cycle in [0, 4]
        const double a0 = prm.gamma_on == QW_GAMMA_ON_ORACLE ? 1.0 : gamma;
        const double d0 = prm.gamma_on == QW_GAMMA_ON_ORACLE ? -gamma : -1.0;
        const double lo = fmin(0.0, 4.0 * a0) + fmin(0.0, d0);
        const double hi = fmax(0.0, 4.0 * a0) + fmax(0.0, d0);
        double half = 0.5 * (hi - lo) * (1.0 + 1e-9);
        if (!(half > 1e-300)) half = 1.0;
        const double b = 0.5 * (hi + lo);
        const double z = half * fabs(time);
        const double kk = z + 12.0 * cbrt(z) + 12.0;
        sh.status = (kk > (double)(CHEB_KMAX - 64)) ? 1 : 0;
        const int K = sh.status ? 0 : (((int)kk + 3) / 4) * 4;     // multiple of 4
        sh.kterms = K;
        sh.cself = (2.0 * a0 - b) / half;
        sh.cnb = -a0 / half;
        sh.cdiag = d0 / half;
        sh.bshift = b;
        if (!sh.status) bessel_table(z, K, sh.jtab);
    }
    __syncthreads();
    const int K = sh.kterms;
    const double cs2 = 2.0 * sh.cself, cn2 = 2.0 * sh.cnb, cd2 = 2.0 * sh.cdiag;
    const double sgn = time < 0.0 ? -1.0 : 1.0;          // exp(+i H |t|): conjugate odd terms

    double tkm[R], tk[R], ar[R], ai[R];
    const double amp = active ? 1.0 / sqrt((double)prm.n) : 0.0;
    // T_0 = psi0, T_1 = Ht psi0 = (cself + 2 cnb) amp everywhere, + cdiag amp at w
#pragma unroll
    for (int r = 0; r < R; ++r) {
        tkm[r] = amp;
        tk[r] = (sh.cself + 2.0 * sh.cnb) * amp;
        ar[r] = sh.jtab[0] * amp;
        ai[r] = 0.0;
    }
    if (t == 0) tk[0] = fma(sh.cdiag, amp, tk[0]);

    double *first = mbox, *last = mbox + 2 * nt;
    int par = 0;
    first[t] = tk[0];
    last[t] = tk[R - 1];

    // One term: accumulate c_k T_k (k mod 4 = PH selects Re/Im and the sign), then
    // T_{k+1} = 2 Ht T_k - T_{k-1} in place of T_{k-1}.  Arrays swap roles every term.
#define QW_CHEB_TERM(PH, X, Y, KIDX)                                                    \
    {                                                                                   \
        __syncthreads();                                                                \
        const double hl = last[par * nt + tl];                                          \
        const double hr = first[par * nt + tr];                                         \
        const double cj = 2.0 * sh.jtab[KIDX];                                          \
        const double w = (PH == 0) ? cj : (PH == 1) ? -sgn * cj : (PH == 2) ? -cj : sgn * cj; \
        const double x0 = X[0];                                                         \
        double pl = hl;                                                                 \
        _Pragma("unroll")                                                               \
        for (int r = 0; r < R; ++r) {                                                   \
            const double c = X[r];                                                      \
            const double nb = pl + ((r + 1 < R) ? X[(r + 1 < R) ? r + 1 : r] : hr);     \
            if (PH == 0 || PH == 2) ar[r] = fma(w, c, ar[r]);                           \
            else ai[r] = fma(w, c, ai[r]);                                              \
            Y[r] = fma(cs2, c, fma(cn2, nb, -Y[r]));                                    \
            pl = c;                                                                     \
        }                                                                               \
        if (t == 0) Y[0] = fma(cd2, x0, Y[0]);                                          \
        par ^= 1;                                                                       \
        first[par * nt + t] = Y[0];                                                     \
        last[par * nt + t] = Y[R - 1];                                                  \
    }
    for (int k = 1; k <= K; k += 4) {
        QW_CHEB_TERM(1, tk, tkm, k)
        QW_CHEB_TERM(2, tkm, tk, k + 1)
        QW_CHEB_TERM(3, tk, tkm, k + 2)
        QW_CHEB_TERM(0, tkm, tk, k + 3)
    }
#undef QW_CHEB_TERM

    // epilogue: p = |psi_w|^2 (not renormalised: the propagator is unitary), ||psi||^2
    double nrm = 0.0;
#pragma unroll
    for (int r = 0; r < R; ++r) nrm = fma(ar[r], ar[r], fma(ai[r], ai[r], nrm));
    const double pw = (t == 0) ? fma(ar[0], ar[0], ai[0] * ai[0]) : 0.0;
    const double2 tot = cheb_block_sum(nrm, pw, sh.red);
    if (t == 0) {
        const bool bad = sh.status != 0;
        prm.p[q] = bad ? NAN : tot.y;
        if (prm.norm) prm.norm[q] = bad ? NAN : tot.x;
    }
    if (prm.psi && active) {
        double sp, cp;
        sincos(-sh.bshift * time, &sp, &cp);                 // global phase e^{-i b t}
        double2 *out = prm.psi + (size_t)q * (size_t)prm.n;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            int64_t j = (int64_t)t * R + r + prm.target;
            if (j >= prm.n) j -= prm.n;
            out[j] = make_double2(ar[r] * cp - ai[r] * sp, ar[r] * sp + ai[r] * cp);
        }
    }
}

template <int R, int MAXNT, int MINB>
cudaError_t launch_cheb(const QwParams &prm, int threads, cudaStream_t st)
{
    const size_t smem = (size_t)threads * 4 * sizeof(double);
    static QwOncePerDevice configured;       // static + dynamic shared memory can pass 48 KB
    if (configured.needed()) {
        cudaError_t e = cudaFuncSetAttribute(cheb_cycle_resident<R, MAXNT, MINB>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             MAXNT * 4 * (int)sizeof(double));
        if (e != cudaSuccess) return e;
        configured.mark();
    }
    cheb_cycle_resident<R, MAXNT, MINB><<<(unsigned)prm.n_points, threads, smem, st>>>(prm);
    return cudaGetLastError();
}

template <int R>
cudaError_t dispatch_cheb(const QwParams &prm, int threads, cudaStream_t st)
{
    if (threads <= 128) return launch_cheb<R, 128, 3>(prm, threads, st);
    if (threads <= 256) return launch_cheb<R, 256, 1>(prm, threads, st);
    if constexpr (R <= 8) {
        if (threads <= 512) return launch_cheb<R, 512, 1>(prm, threads, st);
    }
    if constexpr (R <= 2) {
        if (threads <= 1024) return launch_cheb<R, 1024, 1>(prm, threads, st);
    }
    return cudaErrorInvalidConfiguration;
}

}  // namespace

int qw_resident_sites_per_thread(int64_t n);

// cycle graph, n resident (see qw_resident_sites_per_thread)
bool qw_cheb_supported(const QwParams &prm)
{
    return prm.topology == QW_TOPO_CYCLE && qw_resident_sites_per_thread(prm.n) > 0;
}

cudaError_t qw_launch_cheb(QwParams prm, cudaStream_t st)
{
    int R = qw_resident_sites_per_thread(prm.n);
    if (prm.n % 16 == 0 && prm.n / 16 >= 32 && prm.n / 16 <= 256) R = 16;
    prm.nact = (int)(prm.n / R);
    const int threads = ((prm.nact + 31) / 32) * 32;
    switch (R) {
    case 16: return dispatch_cheb<16>(prm, threads, st);
    case 8: return dispatch_cheb<8>(prm, threads, st);
    case 4: return dispatch_cheb<4>(prm, threads, st);
    case 2: return dispatch_cheb<2>(prm, threads, st);
    default: return dispatch_cheb<1>(prm, threads, st);
    }
}
