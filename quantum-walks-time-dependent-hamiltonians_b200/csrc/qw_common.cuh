// Shared device-side pieces of the RK4 heatmap kernels (sm_100a).
//
// H(t) = a(t) L + d(t) |w><w| is never materialised (the reference rebuilds a dense
// matrix on every right-hand-side call: BCB.py:28-99, 163-174).  Per RK4 stage the whole
// operator is two scalars, pre-multiplied by the stage weights and kept in a small shared
// memory table that is rebuilt every QW_CHUNK steps by the block's own threads.
#pragma once

#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include <atomic>

#include "../../include/qw_b200.h"

#define QW_CHUNK 32               // RK4 steps per coefficient-table refill

// Function attributes (cudaFuncSetAttribute) belong to a device: remember on which devices of
// this process a kernel has been configured.  Setting one twice is harmless, so a benign race
// between host threads only repeats the call.
struct QwOncePerDevice {
    std::atomic<uint64_t> mask{0};
    static int device()
    {
        int d = 0;
        return cudaGetDevice(&d) == cudaSuccess ? d : -1;
    }
    bool needed() const
    {
        const int d = device();
        return d < 0 || d >= 64 || !((mask.load(std::memory_order_acquire) >> d) & 1ull);
    }
    void mark()
    {
        const int d = device();
        if (d >= 0 && d < 64) mask.fetch_or(1ull << d, std::memory_order_release);
    }
};

// Scratch of the large-N paths comes from a memory pool OWNED BY THE LIBRARY (one per device,
// release threshold 1 GiB: freed scratch above that goes back to the driver at the next
// synchronisation), not from the device's default pool, whose attributes belong to the
// application.  qw_scratch_trim() (C ABI) returns everything.
cudaError_t qw_scratch_alloc(void **ptr, size_t bytes, cudaStream_t st);
cudaError_t qw_scratch_free(void *ptr, cudaStream_t st);
// multiprocessors of the current device (cached per device)
int qw_sm_count();

struct QwParams {
    int64_t n;
    int32_t topology, schedule, gamma_on, flags;
    int64_t target;
    double step_size;
    int64_t n_steps_uniform;
    // batch mode
    const double *T;
    const double *gamma;
    const int64_t *n_steps;
    const int64_t *order;
    int64_t n_points;
    // grid mode (time_array != nullptr): point q -> row beta_begin + q / n_time, col q % n_time
    const double *time_array;
    const double *beta_array;
    int64_t n_time;
    int64_t beta_begin;
    // outputs
    double *p;
    double *norm;
    double2 *psi;
    // geometry of the resident kernels
    int32_t nact;                 // active threads per point (n / sites_per_thread)
    int32_t nfull;                // nested-form kernels: threads 0 .. nfull-1 hold R sites, threads
                                  // nfull .. nact-1 hold R-1 (their last slot is a ghost)
    // scratch of the global-memory kernels (generic: 5 complex vectors per running block;
    // streaming: two psi buffers per point of the chunk in flight)
    double2 *scratch;
    int64_t point_offset;         // streaming: first point of the chunk in flight
    int64_t stream_rot;           // streaming: psi is stored rotated, logical l = (j - rot) mod n
    int64_t mirror_m;             // streaming, QW_FLAG_MIRROR: sites of the half ring (n/2 + 1),
                                  // stored behind 3 mirrored ghost sites; 0 = full ring
    int32_t stream_pad;           // streaming, bulk-copy kernel: psi is stored with one pad element
                                  // after every stream_pad sites (0: dense)
    // complete graph above 4096 vertices: a point is split over `split` thread blocks that share
    // nothing but the tabulated (y_w, S) subsystem; partial norms meet in split_part, the last
    // block of a point to arrive (split_done) writes p and norm
    int64_t steps_bound;          // largest step count of the call if the HOST knows it (host-buffer
                                  // entry points; the uniform n_steps rule needs no bound), else 0
    int32_t split;
    double *split_part;           // [n_points][split][2]
    unsigned int *split_done;     // [n_points], zero before the launch
};

struct QwPoint {
    double T, gamma, h;
    int64_t steps;
    int64_t q;                    // output index
};

// g_T(tau): BCB.py:75-87, QW_Search.py:104-109 (thesis Eq. 2.5-2.6)
__host__ __device__ __forceinline__ double qw_schedule(double tau, int schedule)
{
    double x;
    switch (schedule) {
    case QW_SCHED_LIN:   return tau;
    case QW_SCHED_SQRT:  return sqrt(tau);
    case QW_SCHED_CBRT:  return cbrt(tau);
    case QW_SCHED_CERF3: x = 2.0 * tau - 1.0; return 0.5 * (1.0 + x * x * x);
    case QW_SCHED_CERF5: x = 2.0 * tau - 1.0; return 0.5 * (1.0 + x * x * x * x * x);
    case QW_SCHED_CERF7: x = 2.0 * tau - 1.0; return 0.5 * (1.0 + x * x * x * x * x * x * x);
    default:             return 1.0;
    }
}

// (a, d) of H = a L + d |w><w| at normalised time tau.
//   gamma on the oracle (BCB.py:93-97):      a = 1-g,          d = -g*gamma
//   gamma on the Laplacian (thesis Eq 2.14): a = (1-g)*gamma,  d = -g
//   static (BCB.py:101-135):                 a = 1 (gamma),    d = -gamma (-1)
__host__ __device__ __forceinline__ double2 qw_operator_scalars(double tau, double gamma, int schedule,
                                                       int gamma_on)
{
    if (schedule == QW_SCHED_STATIC)
        return gamma_on == QW_GAMMA_ON_ORACLE ? make_double2(1.0, -gamma)
                                              : make_double2(gamma, -1.0);
    double g = qw_schedule(tau, schedule);
    return gamma_on == QW_GAMMA_ON_ORACLE ? make_double2(1.0 - g, -g * gamma)
                                          : make_double2((1.0 - g) * gamma, -g);
}

// Which point does block `b` integrate, and with which step plan.
//   step_size > 0 : h = step_size, steps = int(T/h)           (QW_Search.py:153-154)
//   otherwise     : steps = n_steps, h = T/steps; T == 0 -> no steps (BCB.py:72-73)
__device__ __forceinline__ QwPoint qw_point(const QwParams &prm, int64_t b)
{
    QwPoint pt;
    int64_t q = prm.order ? prm.order[b] : b;
    if (prm.time_array && prm.step_size > 0.0 && !prm.order) {
        // fixed step size: work per point ~ T.  Walk the grid column by column from the
        // longest T to the shortest so that the last blocks to start are the cheapest
        // (longest-processing-time-first; blocks are dispatched in blockIdx order).
        const int64_t rows = prm.n_points / prm.n_time;
        const int64_t col = b / rows, row = b - col * rows;
        const bool ascending = prm.time_array[0] <= prm.time_array[prm.n_time - 1];
        q = row * prm.n_time + (ascending ? prm.n_time - 1 - col : col);
    }
    pt.q = q;
    if (prm.time_array) {
        pt.T = prm.time_array[q % prm.n_time];
        pt.gamma = prm.beta_array[prm.beta_begin + q / prm.n_time];
    } else {
        pt.T = prm.T[q];
        pt.gamma = prm.gamma[q];
    }
    if (prm.step_size > 0.0) {
        pt.h = prm.step_size;
        double r = pt.T / pt.h;
        pt.steps = (r > 0.0) ? (int64_t)r : 0;
    } else {
        int64_t s = prm.n_steps ? prm.n_steps[q] : prm.n_steps_uniform;
        if (pt.T == 0.0 || s <= 0) { pt.steps = 0; pt.h = 0.0; }
        else { pt.steps = s; pt.h = pt.T / (double)s; }
    }
    return pt;
}

// Stage-weighted operator scalars for QW_CHUNK steps starting at step n0.
//   tabAB[4*s+k] = (A, B): next-stage input is y + A*(-i)(L x), accumulator gets B*(-i)(L x)
//   tabD [4*s+k] = (DA, DB): the same weights applied to d(t), used at the marked vertex
// with k = 0..3 the RK4 stage, stage times t_n, t_n+h/2, t_n+h/2, t_n+h, t_n = n*h and
// tau = t/T by division (Schroedinger_Solver.py:56).  `ad` is staging for 3*QW_CHUNK
// operator evaluations.  Ends with a __syncthreads().
__device__ __forceinline__ void qw_fill_table(double2 *tabAB, double2 *tabD, double2 *ad,
                                              const QwPoint &pt, int64_t n0, int schedule,
                                              int gamma_on, int count = QW_CHUNK)
{
    for (int i = threadIdx.x; i < 3 * count; i += blockDim.x) {
        int s = i / 3, k = i - 3 * s;
        double t = (double)(n0 + s) * pt.h;
        if (k == 1) t += 0.5 * pt.h;
        if (k == 2) t += pt.h;
        ad[i] = qw_operator_scalars(t / pt.T, pt.gamma, schedule, gamma_on);
    }
    __syncthreads();
    const double h = pt.h;
    for (int i = threadIdx.x; i < 4 * count; i += blockDim.x) {
        int s = i >> 2, k = i & 3;
        double2 c = ad[3 * s + (k == 0 ? 0 : (k == 3 ? 2 : 1))];
        double wa = (k < 2) ? 0.5 * h : (k == 2 ? h : 0.0);
        double wb = (k == 0 || k == 3) ? h / 6.0 : h / 3.0;
        tabAB[i] = make_double2(wa * c.x, wb * c.x);
        tabD[i] = make_double2(wa * c.y, wb * c.y);
    }
    __syncthreads();
}

__device__ __forceinline__ double qw_warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Block-wide sum of two doubles; result valid in every thread.  red: >= 64 doubles of
// shared memory.  Fixed summation tree -> bitwise reproducible run to run.
__device__ __forceinline__ double2 qw_block_sum2(double a, double b, double *red)
{
    a = qw_warp_sum(a);
    b = qw_warp_sum(b);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nwarps = (blockDim.x + 31) >> 5;
    __syncthreads();
    if (lane == 0) { red[2 * warp] = a; red[2 * warp + 1] = b; }
    __syncthreads();
    double x = (lane < nwarps) ? red[2 * lane] : 0.0;
    double y = (lane < nwarps) ? red[2 * lane + 1] : 0.0;
    x = qw_warp_sum(x);
    y = qw_warp_sum(y);
    return make_double2(x, y);
}
