// Development probe: what limits FP64 issue on sm_100a for the instruction patterns of the
// resident kernels?  Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o
// tools/fp64_mix_probe tools/fp64_mix_probe.cu ;  run on the GPU box.
// Prints FP64 warp-instructions per cycle per SM sub-partition (the pipe peak is 0.5).
#include <cstdio>
#include <cuda_runtime.h>

constexpr int C = 16;

// A: one varying operand (the library's peak probe)
__global__ void __launch_bounds__(256) kA(int iters, double seed, double *sink)
{
    double acc[C];
#pragma unroll
    for (int c = 0; c < C; ++c) acc[c] = seed + c + threadIdx.x;
    const double m = 1.0 + 1e-9 * seed, b = 1e-9;
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int k = 0; k < 16; ++k)
#pragma unroll
            for (int c = 0; c < C; ++c) acc[c] = fma(acc[c], m, b);
    double s = 0;
#pragma unroll
    for (int c = 0; c < C; ++c) s += acc[c];
    if (s == 123.456) sink[0] = s;
}

// B: three distinct varying register operands
__global__ void __launch_bounds__(256) kB(int iters, double seed, double *sink)
{
    double acc[C], x[C], y[C];
#pragma unroll
    for (int c = 0; c < C; ++c) { acc[c] = seed + c + threadIdx.x; x[c] = 1.0 + 1e-9 * (c + seed); y[c] = 1e-9 * (c + 1) * seed; }
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int k = 0; k < 16; ++k)
#pragma unroll
            for (int c = 0; c < C; ++c) acc[c] = fma(x[c], acc[c], y[(c + k) % C]);
    double s = 0;
#pragma unroll
    for (int c = 0; c < C; ++c) s += acc[c];
    if (s == 123.456) sink[0] = s;
}

// Cc: DADD with two varying operands
__global__ void __launch_bounds__(256) kC(int iters, double seed, double *sink)
{
    double acc[C], x[C];
#pragma unroll
    for (int c = 0; c < C; ++c) { acc[c] = seed + c + threadIdx.x; x[c] = 1e-9 * (c + seed); }
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int k = 0; k < 16; ++k)
#pragma unroll
            for (int c = 0; c < C; ++c) {           // bounded: (a, b) -> (a - b, a), period 6
                if (k & 1) x[c] = x[c] + acc[c]; else acc[c] = acc[c] - x[c];
            }
    double s = 0;
#pragma unroll
    for (int c = 0; c < C; ++c) s += acc[c];
    if (s == 123.456) sink[0] = s;
}

// F: two fresh operands and one reused: acc = fma(rho, acc, y[c])
__global__ void __launch_bounds__(256) kF(int iters, double seed, double *sink)
{
    double acc[C], y[C];
#pragma unroll
    for (int c = 0; c < C; ++c) { acc[c] = seed + c + threadIdx.x; y[c] = 1e-9 * (c + 1) * seed; }
    const double rho = 1.0 - 1e-9 * seed;
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int k = 0; k < 16; ++k)
#pragma unroll
            for (int c = 0; c < C; ++c) acc[c] = fma(rho, acc[c], y[c]);
    double s = 0;
#pragma unroll
    for (int c = 0; c < C; ++c) s += acc[c];
    if (s == 123.456) sink[0] = s;
}

// H: the complete-graph level: t = fma(nd, z, -S) (1 fresh), z = fma(rho, t, y) (2 fresh)
__global__ void __launch_bounds__(256) kH(int iters, double seed, double *sink)
{
    double z[C], y[C];
#pragma unroll
    for (int c = 0; c < C; ++c) { z[c] = seed + c + threadIdx.x; y[c] = 1e-9 * (c + 1) * seed; }
    const double rho = 1e-3 * seed, nd = 1.0 + 1e-9 * seed, mS = 1e-9 * seed;
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int k = 0; k < 8; ++k)
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const double t = fma(nd, z[c], mS);
                z[c] = fma(rho, t, y[c]);
            }
    double s = 0;
#pragma unroll
    for (int c = 0; c < C; ++c) s += z[c];
    if (s == 123.456) sink[0] = s;
}

// H2: the same with the two instruction types grouped (all t first, then all z): does the
// operand reuse cache (same register in the same operand slot of consecutive DFMAs) matter?
template <int G>
__global__ void __launch_bounds__(256) kH2(int iters, double seed, double *sink)
{
    double z[C], y[C];
#pragma unroll
    for (int c = 0; c < C; ++c) { z[c] = seed + c + threadIdx.x; y[c] = 1e-9 * (c + 1) * seed; }
    const double rho = 1e-3 * seed, nd = 1.0 + 1e-9 * seed, mS = 1e-9 * seed;
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int k = 0; k < 8; ++k)
#pragma unroll
            for (int g = 0; g < C; g += G) {
                double t[G];
#pragma unroll
                for (int c = 0; c < G; ++c) t[c] = fma(nd, z[g + c], mS);
                asm volatile("" ::: "memory");
#pragma unroll
                for (int c = 0; c < G; ++c) z[g + c] = fma(rho, t[c], y[g + c]);
                asm volatile("" ::: "memory");
            }
    double s = 0;
#pragma unroll
    for (int c = 0; c < C; ++c) s += z[c];
    if (s == 123.456) sink[0] = s;
}

// H3: d = z - zbar (DADD), z = fma(rhoN, d, y)
__global__ void __launch_bounds__(256) kH3(int iters, double seed, double *sink)
{
    double z[C], y[C];
#pragma unroll
    for (int c = 0; c < C; ++c) { z[c] = seed + c + threadIdx.x; y[c] = 1e-9 * (c + 1) * seed; }
    const double rho = 1e-3 * seed, zbar = 1e-9 * seed;
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int k = 0; k < 8; ++k)
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const double d = z[c] - zbar;
                z[c] = fma(rho, d, y[c]);
            }
    double s = 0;
#pragma unroll
    for (int c = 0; c < C; ++c) s += z[c];
    if (s == 123.456) sink[0] = s;
}

// D / E: the nested-form level on R = 16 sites per thread (ring closed inside the thread),
// 24 FP64 instructions per site and step; SYNC adds one block barrier per level
// RHO: 0 = kernel-parameter derived (uniform register), 1 = loaded from shared memory per level
// (vector register, as in the library), 2 = shared memory value moved to a uniform register
// through two warp-wide REDUX
template <bool SYNC, int MINB, int RHO = 0>
__global__ void __launch_bounds__(64, MINB) kD(int steps, double seed, double *sink)
{
    constexpr int R = 16;
    __shared__ double tab[64];
    if (threadIdx.x < 64) tab[threadIdx.x] = 1e-3 * seed * (1.0 + 1e-3 * threadIdx.x);
    __syncthreads();
    double yr[R], yi[R], zr[R], zi[R];
#pragma unroll
    for (int r = 0; r < R; ++r) { yr[r] = seed + r + threadIdx.x; yi[r] = 0.5 * r; zr[r] = 0; zi[r] = 0; }
    double rho = 1e-3 * seed;
    for (int st = 0; st < steps; ++st) {
#pragma unroll
        for (int lev = 0; lev < 4; ++lev) {
            if (SYNC) __syncthreads();
            if (RHO >= 1) rho = ((volatile double *)tab)[(4 * st + lev) & 63];
            if (RHO == 2) {
                const unsigned lo = __reduce_or_sync(0xffffffffu, (unsigned)__double2loint(rho));
                const unsigned hi = __reduce_or_sync(0xffffffffu, (unsigned)__double2hiint(rho));
                rho = __hiloint2double((int)hi, (int)lo);
            }
            // in place, ascending, with the old left neighbour carried along (as the kernel does)
            const double f0r = lev == 0 ? yr[0] : zr[0], f0i = lev == 0 ? yi[0] : zi[0];
            double pr = lev == 0 ? yr[R - 1] : zr[R - 1], pi = lev == 0 ? yi[R - 1] : zi[R - 1];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const double cr = lev == 0 ? yr[r] : zr[r], ci = lev == 0 ? yi[r] : zi[r];
                const double qr = r + 1 < R ? (lev == 0 ? yr[(r + 1) % R] : zr[(r + 1) % R]) : f0r;
                const double qi = r + 1 < R ? (lev == 0 ? yi[(r + 1) % R] : zi[(r + 1) % R]) : f0i;
                const double vr = fma(2.0, cr, -(pr + qr));
                const double vi = fma(2.0, ci, -(pi + qi));
                if (lev == 3) { yr[r] = fma(rho, vi, yr[r]); yi[r] = fma(-rho, vr, yi[r]); }
                else { zr[r] = fma(rho, vi, yr[r]); zi[r] = fma(-rho, vr, yi[r]); }
                pr = cr; pi = ci;
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int r = 0; r < R; ++r) s += yr[r] + yi[r];
    if (s == 123.456) sink[0] = s;
}

// G: the nested level in DIFFERENCE form: d_r = x_r - x_{r-1} (DADD), v_r = d_r - d_{r+1} (DADD),
// out = fma(rho, v, y) -- the same three instructions per component, but two two-operand DADDs
// instead of DADD + DFMA-with-immediate (round 2: does the cheaper operand pattern pay?)
template <bool SYNC, int MINB>
__global__ void __launch_bounds__(64, MINB) kG(int steps, double seed, double *sink)
{
    constexpr int R = 16;
    __shared__ double tab[64];
    if (threadIdx.x < 64) tab[threadIdx.x] = 1e-3 * seed * (1.0 + 1e-3 * threadIdx.x);
    __syncthreads();
    double yr[R], yi[R], zr[R], zi[R];
#pragma unroll
    for (int r = 0; r < R; ++r) { yr[r] = seed + r + threadIdx.x; yi[r] = 0.5 * r; zr[r] = 0; zi[r] = 0; }
    double rho = 1e-3 * seed;
    for (int st = 0; st < steps; ++st) {
#pragma unroll
        for (int lev = 0; lev < 4; ++lev) {
            if (SYNC) __syncthreads();
            rho = ((volatile double *)tab)[(4 * st + lev) & 63];
            double dr[R + 1], di[R + 1];
#pragma unroll
            for (int r = 0; r <= R; ++r) {
                const int a = r % R, b = (r + R - 1) % R;
                dr[r] = (lev == 0 ? yr[a] : zr[a]) - (lev == 0 ? yr[b] : zr[b]);
                di[r] = (lev == 0 ? yi[a] : zi[a]) - (lev == 0 ? yi[b] : zi[b]);
            }
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const double vr = dr[r] - dr[r + 1], vi = di[r] - di[r + 1];
                if (lev == 3) { yr[r] = fma(rho, vi, yr[r]); yi[r] = fma(-rho, vr, yi[r]); }
                else { zr[r] = fma(rho, vi, yr[r]); zi[r] = fma(-rho, vr, yi[r]); }
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int r = 0; r < R; ++r) s += yr[r] + yi[r];
    if (s == 123.456) sink[0] = s;
}

template <typename K>
void run(const char *name, K k, int blocks_per_sm, int threads, int iters, double inst_per_thread_iter)
{
    int dev = 0, sms = 0, khz = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
    double *sink;
    cudaMalloc(&sink, 8);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int blocks = sms * blocks_per_sm;
    k<<<blocks, threads>>>(iters / 8 + 1, 1.0, sink);
    cudaEventRecord(e0);
    k<<<blocks, threads>>>(iters, 1.0, sink);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double warp_inst = (double)blocks * (threads / 32) * iters * inst_per_thread_iter;
    const double cycles = ms * 1e-3 * khz * 1e3;
    printf("%-34s %d blk/SM x %3d thr  %8.3f ms  %.4f FP64 warp-inst/clk/SMSP (peak 0.5)  err=%s\n", name,
           blocks_per_sm, threads, ms, warp_inst / cycles / (sms * 4.0), cudaGetErrorString(cudaGetLastError()));
    cudaFree(sink);
}

int main(int argc, char **argv)
{
    if (argc > 1) {          // round 2: only the level forms
        run("E nested level, barrier/level", kD<true, 4, 1>, 4, 64, 20000, 16 * 24.0);
        run("E nested level, barrier/level", kD<true, 6, 1>, 6, 64, 20000, 16 * 24.0);
        run("G difference form, barrier/level", kG<true, 4>, 4, 64, 20000, 16 * 24.0 + 4 * 2.0);
        run("G difference form, barrier/level", kG<true, 6>, 6, 64, 20000, 16 * 24.0 + 4 * 2.0);
        run("G difference form, no barrier", kG<false, 6>, 6, 64, 20000, 16 * 24.0 + 4 * 2.0);
        return 0;
    }
    run("A dfma 1 varying operand", kA, 2, 256, 4000, 256.0);
    run("A dfma 1 varying operand", kA, 1, 256, 4000, 256.0);
    run("B dfma 3 varying operands", kB, 2, 256, 4000, 256.0);
    run("B dfma 3 varying operands", kB, 1, 256, 4000, 256.0);
    run("C dadd 2 varying operands", kC, 2, 256, 4000, 256.0);
    run("C dadd 2 varying operands", kC, 1, 256, 4000, 256.0);
    run("F dfma 2 fresh + 1 reused", kF, 2, 256, 4000, 256.0);
    run("F dfma 2 fresh + 1 reused", kF, 1, 256, 4000, 256.0);
    run("H complete-graph level pattern", kH, 2, 256, 4000, 256.0);
    run("H complete-graph level pattern", kH, 1, 256, 4000, 256.0);
    run("H2 grouped by type, groups of 16", kH2<16>, 2, 256, 4000, 256.0);
    run("H2 grouped by type, groups of 16", kH2<16>, 1, 256, 4000, 256.0);
    run("H2 grouped by type, groups of 8", kH2<8>, 2, 256, 4000, 256.0);
    run("H2 grouped by type, groups of 4", kH2<4>, 2, 256, 4000, 256.0);
    run("H3 dadd (z - zbar) + dfma", kH3, 2, 256, 4000, 256.0);
    run("H3 dadd (z - zbar) + dfma", kH3, 1, 256, 4000, 256.0);
    run("D nested level, no barrier", kD<false, 4>, 4, 64, 20000, 16 * 24.0);
    run("D nested level, no barrier", kD<false, 6>, 6, 64, 20000, 16 * 24.0);
    run("D nested level, no barrier", kD<false, 4>, 2, 64, 20000, 16 * 24.0);
    run("E nested level, barrier/level", kD<true, 4>, 4, 64, 20000, 16 * 24.0);
    run("E nested level, barrier/level", kD<true, 6>, 6, 64, 20000, 16 * 24.0);
    run("E rho from shared (vector reg)", kD<true, 4, 1>, 4, 64, 20000, 16 * 24.0);
    run("E rho from shared (vector reg)", kD<true, 6, 1>, 6, 64, 20000, 16 * 24.0);
    run("E rho shared -> REDUX -> uniform", kD<true, 4, 2>, 4, 64, 20000, 16 * 24.0);
    run("E rho shared -> REDUX -> uniform", kD<true, 6, 2>, 6, 64, 20000, 16 * 24.0);
    return 0;
}
