// Development probe: do non-FP64 instructions cost FP64 issue time on sm_100a?
// Each iteration issues 64 independent-chain DFMAs plus K instructions of another kind.
// Prints clocks per iteration per warp-scheduler share (3 warps per SM sub-partition, as in the
// hot kernel).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/hp_bin/dispatch_probe tools/dispatch_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int K, int KIND>
__global__ void __launch_bounds__(64, 6) kmix(int iters, double seed, int iseed, double *sink)
{
    __shared__ double2 buf[64 * 4];
    buf[threadIdx.x] = make_double2(seed, seed); buf[threadIdx.x + 64] = buf[threadIdx.x];
    buf[threadIdx.x + 128] = buf[threadIdx.x]; buf[threadIdx.x + 192] = buf[threadIdx.x];
    __syncthreads();
    const unsigned sa = (unsigned)__cvta_generic_to_shared(buf + threadIdx.x);
    double lx = 0, ly = 0;
    double acc[16];
    int x[16];
    float f[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) { acc[c] = seed + c + threadIdx.x; x[c] = iseed + c * 77 + threadIdx.x; f[c] = (float)(seed + c); }
    const double m = 1.0 + 1e-9 * seed, b = 1e-9;
    const int y = iseed * 3 + 1;
    const float fm = 1.0f + 1e-7f * (float)seed;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
#pragma unroll
            for (int c = 0; c < 16; ++c) {
                acc[c] = fma(acc[c], m, b);
                if ((k * 16 + c) * K / 64 != (k * 16 + c + 1) * K / 64) {   // K extra instructions per 64 DFMA
                    const int j = (k * 16 + c) & 15;
                    if (KIND == 0) asm volatile("add.s32 %0, %0, %1;" : "+r"(x[j]) : "r"(y));
                    if (KIND == 1) asm volatile("add.f32 %0, %0, %1;" : "+f"(f[j]) : "f"(fm));
                    if (KIND == 2) asm volatile("slct.s32.s32 %0, %0, %1, %2;" : "+r"(x[j]) : "r"(y), "r"(iseed));
                    if (KIND == 3) asm volatile("shfl.sync.idx.b32 %0, %0, %1, 0x1f, 0xffffffff;" : "+r"(x[j]) : "r"(j));
                    if (KIND == 4) asm volatile("mad.lo.s32 %0, %0, %1, %1;" : "+r"(x[j]) : "r"(y));
                    if (KIND == 5) asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(lx), "=d"(ly) : "r"(sa + 1024 * (j & 3)));
                    if (KIND == 6) asm volatile("st.shared.v2.f64 [%0], {%1, %2};" :: "r"(sa + 1024 * (j & 3)), "d"(m), "d"(b) : "memory");
                    if (KIND == 7) asm volatile("{\n.reg .pred p;\nsetp.eq.s32 p, %3, 12345;\n@p st.shared.v2.f64 [%0], {%1, %2};\n}" :: "r"(sa + 1024 * (j & 3)), "d"(m), "d"(b), "r"(iseed) : "memory");
                    if (KIND == 8) asm volatile("bar.sync 0;" ::: "memory");
                }
            }
        }
    }
    double s = 0; int xs = 0; float fs = 0;
#pragma unroll
    for (int c = 0; c < 16; ++c) { s += acc[c]; xs += x[c]; fs += f[c]; }
    if (s == 123.456 || xs == 123457 || fs == 1.2345f || lx + ly == 77.7) sink[0] = s + xs + fs;
}

template <typename Kf>
void run(const char *name, Kf k, int extra)
{
    int sms = 0, khz = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    double *sink; cudaMalloc(&sink, 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000, blocks = sms * 6;
    k<<<blocks, 64>>>(iters / 8, 1.0, 3, sink);
    cudaEventRecord(e0);
    k<<<blocks, 64>>>(iters, 1.0, 3, sink);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double clk = ms * 1e-3 * khz * 1e3;
    // 12 warps per SM = 3 per sub-partition: clocks per iteration of one sub-partition's 3 warps
    printf("%-10s +%2d per 64 DFMA: %7.1f clk per (3 warps x 1 iteration)  = %.3f clk per DFMA-equivalent slot; err=%s\n",
           name, extra, clk / iters, clk / iters / (3.0 * 64), cudaGetErrorString(cudaGetLastError()));
}

int main()
{
    run("none", kmix<0, 0>, 0);
    run("IADD", kmix<16, 0>, 16); run("IADD", kmix<32, 0>, 32); run("IADD", kmix<64, 0>, 64);
    run("FADD", kmix<16, 1>, 16); run("FADD", kmix<32, 1>, 32); run("FADD", kmix<64, 1>, 64);
    run("SEL", kmix<16, 2>, 16); run("SEL", kmix<32, 2>, 32); run("SEL", kmix<64, 2>, 64);
    run("SHFL", kmix<16, 3>, 16); run("SHFL", kmix<32, 3>, 32);
    run("IMAD", kmix<16, 4>, 16); run("IMAD", kmix<32, 4>, 32); run("IMAD", kmix<64, 4>, 64);
    run("LDS.128", kmix<4, 5>, 4); run("LDS.128", kmix<8, 5>, 8); run("LDS.128", kmix<16, 5>, 16); run("LDS.128", kmix<32, 5>, 32);
    run("STS.128", kmix<4, 6>, 4); run("STS.128", kmix<8, 6>, 8); run("STS.128", kmix<16, 6>, 16); run("STS.128", kmix<32, 6>, 32);
    run("@!p STS", kmix<8, 7>, 8); run("@!p STS", kmix<16, 7>, 16); run("@!p STS", kmix<32, 7>, 32);
    run("BAR", kmix<1, 8>, 1); run("BAR", kmix<2, 8>, 2); run("BAR", kmix<4, 8>, 4);
    return 0;
}
