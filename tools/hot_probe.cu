// Development probe for the hot cycle kernel (C5): compiles csrc/qw_horner.cu into this file and
// launches rk4_cycle_horner<16, 64, 6> directly, so that ablations (-DQW_ABL_*) can be timed
// without touching the library.  Build:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo [-DQW_ABL_NOBAR ...]
//        -o tools/hot_probe tools/hot_probe.cu
#include <cstdio>
#include <vector>
#include "../quantum-walks-time-dependent-hamiltonians_b200/csrc/qw_horner.cu"

cudaError_t qw_scratch_alloc(void **p, size_t b, cudaStream_t st) { return cudaMallocAsync(p, b, st); }
cudaError_t qw_scratch_free(void *p, cudaStream_t st) { return cudaFreeAsync(p, st); }
int qw_sm_count() { return 148; }

int main(int argc, char **argv)
{
    const int rows = argc > 1 ? atoi(argv[1]) : 8;
    const int nT = 1024, n = 1024, S = 4096;
    std::vector<double> T(nT), B(rows);
    for (int i = 0; i < nT; ++i) T[i] = 0.1 + (1024.0 - 0.1) * i / (nT - 1);
    for (int i = 0; i < rows; ++i) B[i] = rows > 1 ? 2.5 * i / (rows - 1) : 1.0;
    double *dT, *dB, *dp, *dn;
    cudaMalloc(&dT, nT * 8); cudaMalloc(&dB, rows * 8);
    cudaMalloc(&dp, (size_t)nT * rows * 8); cudaMalloc(&dn, (size_t)nT * rows * 8);
    cudaMemcpy(dT, T.data(), nT * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), rows * 8, cudaMemcpyHostToDevice);
    QwParams prm = {};
    prm.n = n; prm.topology = QW_TOPO_CYCLE; prm.schedule = QW_SCHED_LIN; prm.gamma_on = QW_GAMMA_ON_ORACLE;
    prm.target = (n - 1) / 2; prm.n_steps_uniform = S; prm.n_points = (int64_t)nT * rows;
    prm.time_array = dT; prm.beta_array = dB; prm.n_time = nT; prm.p = dp; prm.norm = dn;
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        cudaError_t e = qw_launch_horner(prm, 0, nullptr);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
        if (e != cudaSuccess || cudaGetLastError() != cudaSuccess) { printf("launch error\n"); return 1; }
    }
    std::vector<double> p((size_t)nT * rows);
    cudaMemcpy(p.data(), dp, p.size() * 8, cudaMemcpyDeviceToHost);
    double cs = 0; for (double v : p) cs += v;
    printf("%-28s rows=%d  %9.3f ms  %.4e site-updates/s  checksum %.15e  p[last]=%.12f\n",
#ifdef QW_ABL_NAME
           QW_ABL_NAME,
#else
           "baseline",
#endif
           rows, best, (double)nT * rows * n * S / (best * 1e-3), cs, p.back());
    return 0;
}
