// Tile-fused HBM-streaming RK4 for rings too large for one SM's registers (cycle graph,
// N = 65 536 in BASELINE config 5), sm_100a.
//
// psi of every point in flight lives in HBM (two buffers, ping-pong).  One launch advances
// every point by KB RK4 steps: a thread block loads a window of W = 128 threads x 8 sites
// (coalesced, transposed through padded shared memory), runs the same register-resident
// stage code as qw_resident.cu on it with open ends, and writes back the W - 2*HALO
// central sites (HALO = 4*KB: every stage invalidates one more site at each open end, so
// all four stages of KB steps are fused per tile).  Algorithmic traffic: one read and one
// write of psi per pass = 32 B per site per pass, i.e. 32/KB bytes per site-update; KB = 1
// is the plain "read psi_n, write psi_n+1" streaming roofline of SURVEY.md section 8d,
// KB = 4 (default) is temporal blocking that turns the path FP64-bound again.
//
// Replaces the same reference path as the resident kernels (BCB.py:163-203 with the step
// rule of QW_Search.py:149-164); vertices keep their physical labels, so the marked vertex
// can sit in any slot of any window (including the redundant halo copies).
#include "qw_stage.cuh"

namespace {

constexpr int SR = 8;                  // sites per thread
constexpr int SNT = 128;               // threads per block
constexpr int SW = SR * SNT;           // window
constexpr int SPAD = SW + SW / SR;     // one pad element per thread segment: no bank conflicts

__device__ __forceinline__ int pad_index(int e) { return e + e / SR; }

struct StreamTables {
    double2 ab[4 * 4];
    double2 d[4 * 4];
    double2 ad[3 * 4];
    double red[128];
};

template <int STAGE>
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
template <int STAGE>
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

template <int KB>
__global__ void __launch_bounds__(SNT, 4)
rk4_cycle_stream(const QwParams prm, const double2 *__restrict__ in, double2 *__restrict__ out,
                 int64_t step0, int tiles)
{
    constexpr int HALO = 4 * KB;
    constexpr int TILE = SW - 2 * HALO;
    __shared__ double2 xbuf[SPAD];                 // transpose buffer
    __shared__ double2 mbox[4 * SNT];              // [parity][first/last][thread]
    __shared__ StreamTables tb;

    const int t = threadIdx.x;
    const int64_t local = blockIdx.x / tiles;      // point within the chunk in flight
    const int tile = blockIdx.x - (int)(local * tiles);
    const QwPoint pt = qw_point(prm, prm.point_offset + local);
    const int64_t n = prm.n;
    const double2 *src = in + local * n;
    double2 *dst = out + local * n;
    const int64_t first = (int64_t)tile * TILE - HALO;         // window start (may be < 0)
    int64_t base = first % n;
    if (base < 0) base += n;

    int64_t todo = pt.steps - step0;
    const int kb = todo < 0 ? 0 : (todo > KB ? KB : (int)todo);

    // coalesced load -> padded shared memory -> 8 consecutive sites per thread
#pragma unroll
    for (int i = 0; i < SR; ++i) {
        const int e = i * SNT + t;
        int64_t j = base + e;
        if (j >= n) { j -= n; if (j >= n) j %= n; }
        xbuf[pad_index(e)] = src[j];
    }
    __syncthreads();
    State<SR> s;
#pragma unroll
    for (int r = 0; r < SR; ++r) {
        const double2 v = xbuf[pad_index(t * SR + r)];
        s.yr[r] = v.x; s.yi[r] = v.y;
        s.ur[r] = 0.0; s.ui[r] = 0.0; s.ar[r] = 0.0; s.ai[r] = 0.0;
    }

    if (kb > 0) {
        // slot of the marked vertex in this thread's segment, or -1 (windows repeat mod n)
        int64_t wrel = (prm.target - first - (int64_t)t * SR) % n;
        if (wrel < 0) wrel += n;
        const int own_slot = wrel < SR ? (int)wrel : -1;
        const int tl = t == 0 ? 0 : t - 1, tr = t == SNT - 1 ? SNT - 1 : t + 1;   // open ends
        const double2 zero2 = make_double2(0.0, 0.0);
        double2 *fst = mbox, *lst = mbox + 2 * SNT;
        int par = 0;
        fst[t] = make_double2(s.yr[0], s.yi[0]);
        lst[t] = make_double2(s.yr[SR - 1], s.yi[SR - 1]);
        qw_fill_table(tb.ab, tb.d, tb.ad, pt, step0, prm.schedule, prm.gamma_on, kb);
        for (int sidx = 0; sidx < kb; ++sidx) {
#define QW_STREAM_STAGE(K)                                                              \
            {                                                                           \
                __syncthreads();                                                        \
                const double2 hl = lst[par * SNT + tl];                                 \
                const double2 hr = fst[par * SNT + tr];                                 \
                const double2 ab = tb.ab[4 * sidx + K];                                 \
                double xr = 0.0, xi = 0.0;                                              \
                if (own_slot >= 0) read_slot<K>(s, own_slot, xr, xi);                   \
                rk_stage<SR, K, true>(s, ab.x, ab.y, hl, hr, zero2, 0.0, false, zero2); \
                if (own_slot >= 0) fix_slot<K>(s, own_slot, tb.d[4 * sidx + K], xr, xi); \
                par ^= 1;                                                               \
                if (K < 3) {                                                            \
                    fst[par * SNT + t] = make_double2(s.ur[0], s.ui[0]);                \
                    lst[par * SNT + t] = make_double2(s.ur[SR - 1], s.ui[SR - 1]);      \
                } else {                                                                \
                    fst[par * SNT + t] = make_double2(s.yr[0], s.yi[0]);                \
                    lst[par * SNT + t] = make_double2(s.yr[SR - 1], s.yi[SR - 1]);      \
                }                                                                       \
            }
            QW_STREAM_STAGE(0)
            QW_STREAM_STAGE(1)
            QW_STREAM_STAGE(2)
            QW_STREAM_STAGE(3)
#undef QW_STREAM_STAGE
        }
    }

    // registers -> padded shared memory -> coalesced store of the central TILE sites
    __syncthreads();
#pragma unroll
    for (int r = 0; r < SR; ++r) xbuf[pad_index(t * SR + r)] = make_double2(s.yr[r], s.yi[r]);
    __syncthreads();
#pragma unroll
    for (int i = 0; i < SR; ++i) {
        const int e = i * SNT + t;
        const int64_t j = first + e;
        if (e >= HALO && e < HALO + TILE && j < n) dst[j] = xbuf[pad_index(e)];
    }
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
    const double2 *src = buf + local * n;
    double nrm = 0.0;
    for (int64_t j = threadIdx.x; j < n; j += blockDim.x) {
        const double2 v = src[j];
        nrm = fma(v.x, v.x, fma(v.y, v.y, nrm));
        if (prm.psi) prm.psi[(size_t)pt.q * (size_t)n + j] = v;
    }
    const double2 tot = qw_block_sum2(nrm, 0.0, red);
    if (threadIdx.x == 0) {
        const double2 v = src[prm.target];
        prm.p[pt.q] = fma(v.x, v.x, v.y * v.y) / tot.x;
        if (prm.norm) prm.norm[pt.q] = tot.x;
    }
}

int stream_kb(const QwParams &prm)
{
    if (prm.flags & QW_FLAG_STREAM_KB1) return 1;
    if (prm.flags & QW_FLAG_STREAM_KB2) return 2;
    return 4;
}

template <int KB>
cudaError_t launch_pass(const QwParams &prm, const double2 *in, double2 *out, int64_t step0,
                        int64_t points, cudaStream_t st)
{
    constexpr int TILE = SW - 8 * KB;
    const int tiles = (int)((prm.n + TILE - 1) / TILE);
    rk4_cycle_stream<KB><<<(unsigned)(points * tiles), SNT, 0, st>>>(prm, in, out, step0, tiles);
    return cudaGetLastError();
}

}  // namespace

bool qw_stream_supported(const QwParams &prm)
{
    return prm.topology == QW_TOPO_CYCLE && prm.n >= 8;
}

cudaError_t qw_stream_plan(const QwParams &prm, qw_plan *plan)
{
    cudaFuncAttributes fa;
    int occ = 0;
    cudaError_t e;
    const int kb = stream_kb(prm);
    if (kb == 1) {
        e = cudaFuncGetAttributes(&fa, rk4_cycle_stream<1>);
        if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, rk4_cycle_stream<1>, SNT, 0);
    } else if (kb == 2) {
        e = cudaFuncGetAttributes(&fa, rk4_cycle_stream<2>);
        if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, rk4_cycle_stream<2>, SNT, 0);
    } else {
        e = cudaFuncGetAttributes(&fa, rk4_cycle_stream<4>);
        if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, rk4_cycle_stream<4>, SNT, 0);
    }
    if (e != cudaSuccess) return e;
    plan->path = 3;
    plan->sites_per_thread = SR;
    plan->threads_per_block = SNT;
    plan->blocks_per_sm = occ;
    plan->regs_per_thread = fa.numRegs;
    plan->smem_bytes = (int)fa.sharedSizeBytes;
    plan->steps_per_pass = kb;
    return cudaSuccess;
}

cudaError_t qw_launch_stream(const QwParams &prm_in, cudaStream_t st, int64_t *launches)
{
    QwParams prm = prm_in;
    const int kb = stream_kb(prm);
    const int64_t n = prm.n;
    // points in flight: bounded by ~8 GiB of ping-pong scratch
    int64_t chunk = (int64_t)(8ull << 30) / (2 * n * (int64_t)sizeof(double2));
    if (chunk < 1) chunk = 1;
    if (chunk > prm.n_points) chunk = prm.n_points;
    double2 *bufA = nullptr;
    unsigned long long *dmax = nullptr;
    cudaError_t e = cudaMallocAsync((void **)&bufA, (size_t)chunk * 2 * n * sizeof(double2), st);
    if (e != cudaSuccess) return e;
    double2 *bufB = bufA + chunk * n;
    const bool varying = prm.step_size > 0.0 || prm.n_steps != nullptr;
    if (varying) {
        e = cudaMallocAsync((void **)&dmax, sizeof(unsigned long long), st);
        if (e != cudaSuccess) { cudaFreeAsync(bufA, st); return e; }
    }
    const double amp = 1.0 / sqrt((double)n);

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
        stream_init<<<1184, 256, 0, st>>>(bufA, cnt * n, amp);
        ++*launches;
        double2 *cur = bufA, *nxt = bufB;
        for (int64_t s0 = 0; s0 < max_steps && e == cudaSuccess; s0 += kb) {
            if (kb == 1) e = launch_pass<1>(prm, cur, nxt, s0, cnt, st);
            else if (kb == 2) e = launch_pass<2>(prm, cur, nxt, s0, cnt, st);
            else e = launch_pass<4>(prm, cur, nxt, s0, cnt, st);
            ++*launches;
            double2 *tmp = cur; cur = nxt; nxt = tmp;
        }
        if (e != cudaSuccess) break;
        stream_finish<<<(unsigned)cnt, 512, 0, st>>>(prm, cur);
        ++*launches;
        e = cudaGetLastError();
    }
    cudaFreeAsync(bufA, st);
    if (dmax) cudaFreeAsync(dmax, st);
    return e;
}
