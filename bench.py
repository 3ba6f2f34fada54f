#!/usr/bin/env python
"""bench.py -- RK4 site-updates/s and heatmap grid-points/s of the B200 path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
                    [--workload c5|c2|c3|c4] [--rows-per-gpu R] [--scaling weak|strong]

One "step" is one pass of the hot path over the workload's (T, gamma) points.  The default
workload is BASELINE config 5: cycle graph N=1024, n_steps=4096, schedule lin,
T = linspace(0.1, 1024, 1024); every GPU integrates a block of 128 gamma rows of
gamma = linspace(0, 2.5, 128*N_gpus), so that 8 GPUs compute exactly the named
1024 x 1024 heatmap (weak scaling, no data-path collective; the e2e leg adds the final
gather that replaces ray.get + np.concatenate of BCB.py:284-292).  `--scaling strong` keeps
the TOTAL grid fixed instead (c5: the named 1024 x 1024 heatmap at every N).

Prints ONE JSON line (rank 0).  `value` is device-resident throughput (inputs in HBM,
CUDA events, max over ranks); `e2e` goes through the public Python API with pinned host
inputs, H2D/D2H (and the gather) inside the timed region; `roofline` is the FP64-pipe
roofline of the resident kernel against a DFMA peak measured in this same process;
`cpu_baseline` is the C/OpenMP port of the oracle on this host's cores (bounded sample).
`--impl reference` times that CPU port alone (the reference itself is Python + SciPy RK45
and cannot run these sizes: BASELINE.md section 2).
"""

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOPS_PER_SITE_UPDATE = 52          # SURVEY.md section 8d (22 DFMA + 8 DADD): ALGORITHMIC flops
METRIC = "rk4_site_updates_per_sec"


def kernel_counters(plan):
    """Hardware counters of the kernel the plan names, from committed ncu captures
    (profiles/kernel_counters.json); {} when that kernel has none."""
    try:
        with open(os.path.join(ROOT, "profiles", "kernel_counters.json")) as fh:
            table = json.load(fh)
    except Exception:
        return {}
    key = "%s/R%d/T%d" % (plan["path"], plan["sites_per_thread"], plan["threads_per_block"])
    return table.get(key, {})


def source_fp64_per_site_update(plan):
    """FP64 instructions per site-update counted in the kernel SOURCE / SASS (tools/sass_loops.py):
    the fallback when no ncu opcode counters are committed for the kernel."""
    path, r, thr = plan["path"], plan["sites_per_thread"], plan["threads_per_block"]
    n_block = max(1, r * thr)
    if path == "horner-cycle":
        return 24.0 + 30.0 * 32.0 / n_block
    if path == "packed-horner-cycle":
        return 24.0 + 22.0 * 32.0 / 512.0
    if path == "horner-complete":
        return 16.0 + 22.0 * 32.0 / n_block
    if path == "resident-complete":
        return 22.0
    return 30.0


WORKLOADS = {
    # name: n, topology, schedule, gamma_on, n_steps, n_time, T range, gamma range, rows/GPU
    "c5": dict(n=1024, topology="circular", schedule="lin", gamma_on="oracle", n_steps=4096,
               n_time=1024, t=(0.1, 1024.0), g=(0.0, 2.5), rows=128, rows_total=1024,
               label="C5 cycle N=1024 n_steps=4096 lin, 1024 T x 128 gamma rows per GPU "
                     "(the full 1024x1024 heatmap at 8 GPUs)"),
    "c2": dict(n=64, topology="circular", schedule="lin", gamma_on="oracle", n_steps=1024,
               n_time=100, t=(0.1, 64.0), g=(0.0, 2.5), rows=100, rows_total=100,
               label="C2 cycle N=64 n_steps=1024 lin, 100x100 heatmap per GPU"),
    "c3": dict(n=4096, topology="complete", schedule="cerf_3", gamma_on="laplacian",
               n_steps=2048, n_time=256, t=(0.1, 512.0), g=(0.5 / 4096, 2.5 / 4096), rows=32,
               rows_total=256,
               label="C3 complete N=4096 n_steps=2048 cerf_3 gamma-on-Laplacian, "
                     "256 T x 32 gamma rows per GPU (256x256 at 8 GPUs)"),
    "c4": dict(n=256, topology="circular", schedule="lin", gamma_on="oracle", n_steps=512,
               n_time=1000, t=(20.0, 30.0), g=(0.8, 1.2), rows=100, rows_total=800,
               label="C4-shaped cycle N=256 n_steps=512 lin, 1e5 points per GPU"),
}


def _workload(name, n_gpus, rows_per_gpu=None, scaling="weak"):
    w = dict(WORKLOADS[name])
    if rows_per_gpu:
        w["rows"] = int(rows_per_gpu)
    n_rows = w["rows_total"] if scaling == "strong" else w["rows"] * n_gpus
    w["scaling"] = scaling
    w["time_array"] = np.linspace(w["t"][0], w["t"][1], w["n_time"])
    w["beta_array"] = np.linspace(w["g"][0], w["g"][1], n_rows)
    w["site_updates_per_point"] = w["n"] * w["n_steps"]
    return w


def _config(w, n_gpus):
    """The `config` object of the JSON line: the workload only, identical in both arms."""
    return {"workload": w["label"] if w["scaling"] == "weak" else
            w["label"].split(",")[0] + ", %d T x %d gamma rows in total (strong scaling)"
            % (w["n_time"], w["beta_array"].size),
            "n": w["n"], "topology": w["topology"], "schedule": w["schedule"],
            "gamma_on": w["gamma_on"], "n_steps": w["n_steps"], "n_time": w["n_time"],
            "n_gamma": int(w["beta_array"].size), "n_gpus": n_gpus, "scaling": w["scaling"],
            "points_total": int(w["beta_array"].size * w["n_time"]),
            "site_updates_per_point": w["site_updates_per_point"]}


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons while the timed region runs (nvidia-smi semantics)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag = index, threading.Event()
        self.sm, self.reasons, self.sm_max, self.power = [], set(), None, []

    def run(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.sm_max = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            names = {
                pynvml.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                pynvml.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                pynvml.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                pynvml.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
            }
            while not self.stop_flag.is_set():
                self.sm.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                try:
                    self.power.append(pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0)
                except Exception:
                    pass
                mask = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, nm in names.items():
                    if mask & bit:
                        self.reasons.add(nm)
                time.sleep(0.1)
        except Exception as exc:                      # NVML missing: report it, do not die
            self.reasons.add("nvml_unavailable:%s" % type(exc).__name__)

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=2.0)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None,
                "sm_max_mhz": self.sm_max, "reasons": sorted(self.reasons),
                "power_w_max": max(self.power) if self.power else None,
                "samples": len(self.sm)}


def hbm_peak_gbs():
    """Measured copy bandwidth of this pool's B200s (driver-written), else the fallback of
    /opt/skills/guides/B200_PROFILING.md."""
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
    except Exception:
        return 6650.0, "fallback 6.65 TB/s of B200_PROFILING.md (of fallback)"


def stream_measure(qw, torch, dev, reps=3):
    """HBM-streaming variant of config 5: cycle N=65536, n_steps=256, 32x32 (T, gamma) grid
    = 1024 points in flight (2 GiB of psi ping-pong buffers, far above the 126 MB L2).
    KB = RK4 steps fused per pass over HBM: KB=1 is the plain 32 B/site-update stream."""
    n, S, nT, nB = 65536, 256, 32, 32
    t = torch.linspace(0.1, 64.0, nT, dtype=torch.float64, device=dev)
    b = torch.linspace(0.0, 2.5, nB, dtype=torch.float64, device=dev)
    peak, src = hbm_peak_gbs()
    out = {"workload": "cycle N=65536 n_steps=256 lin, 32x32 heatmap (1024 points, 2 GiB "
                       "working set > L2)", "peak_gbs": peak, "peak_source": src, "variants": []}
    for kb, flag in ((1, 8), (2, 16), (4, 24), (8, 0)):
        prob = qw.Problem(n=n, topology="circular", schedule="lin", n_steps=S, flags=flag)
        qw.rk4_heatmap(prob, t, b)
        best = 1e30
        l0 = qw.launch_count()
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            qw.rk4_heatmap(prob, t, b)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        launches = (qw.launch_count() - l0) // reps
        rate = nT * nB * n * S / (best * 1e-3)
        sr = 8 if kb == 1 else 16                   # sites per thread of the variant
        halo = 1024.0 / (1024.0 - 2.0 * (-(-4 * kb // sr) * sr))   # window / written sites
        out["variants"].append({
            "steps_per_pass": kb, "ms": best, "site_updates_per_sec": rate,
            "launches_per_heatmap": launches,
            "algorithmic_gbs_at_32B_per_site_update": 32.0 * rate / 1e9,
            "frac_of_hbm_peak_algorithmic": 32.0 * rate / 1e9 / peak,
            "actual_gbs_incl_halo_rereads": 32.0 / kb * halo * rate / 1e9,
            "frac_of_hbm_peak_actual": 32.0 / kb * halo * rate / 1e9 / peak})
    kb1 = out["variants"][0]
    try:
        with open(os.path.join(ROOT, "profiles", "kernel_counters.json")) as fh:
            ctr = json.load(fh).get("stream/KB1", {})
    except Exception:
        ctr = {}
    out["roofline"] = {            # the plain read-psi_n / write-psi_n+1 stream (KB = 1)
        "bound": "hbm", "achieved": kb1["algorithmic_gbs_at_32B_per_site_update"],
        "peak": peak, "unit": "GB/s", "frac": kb1["frac_of_hbm_peak_algorithmic"],
        "traffic": ctr.get("dram_bytes_per_pass"),
        "algorithmic_bytes_per_launch": nT * nB * n * 32,
        "traffic_note": ctr.get("dram_source", "no ncu capture committed for this kernel"),
        "kernel": qw.plan(qw.Problem(n=n, topology="circular", schedule="lin", n_steps=S,
                                     flags=8))}
    return out


def mirror_measure(qw, torch, w, d_time, d_beta, rows, flush, pts_local, p_dev, reps=2):
    """The exact symmetry reductions the drop-in modules use by default (QW_FLAG_MIRROR: the
    half ring / the (psi_w, psi_other) pair) on the same rows: the unit of `value` stays the
    full-ring site-update, this reports what a user of the drop-in sees in grid points/s."""
    prob = qw.Problem(n=w["n"], topology=w["topology"], schedule=w["schedule"],
                      gamma_on=w["gamma_on"], n_steps=w["n_steps"], flags=8192)
    qw.rk4_heatmap(prob, d_time, d_beta, rows=rows)
    best, pm = 1e30, None
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        pm, _ = qw.rk4_heatmap(prob, d_time, d_beta, rows=rows)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return {"kernel": qw.plan(prob), "ms_per_step_this_gpu": best,
            "effective_grid_points_per_sec_this_gpu": pts_local / (best * 1e-3),
            "effective_full_ring_site_updates_per_sec_this_gpu":
                pts_local * w["site_updates_per_point"] / (best * 1e-3),
            "max_abs_dp_vs_full_ring_kernel": float((pm - p_dev).abs().max())}


def cpu_port_rate(w, seconds, threads=0):
    """C/OpenMP oracle port on a bounded sample of the workload's own points."""
    from oracle import c_oracle
    nthr = c_oracle.max_threads() if threads <= 0 else threads
    t_arr, b_arr = w["time_array"], w["beta_array"]
    rng = np.random.default_rng(5)

    def sample(k):
        return t_arr[rng.integers(0, t_arr.size, k)], b_arr[rng.integers(0, b_arr.size, k)]

    def run(k):
        T, g = sample(k)
        t0 = time.perf_counter()
        c_oracle.rk4_batch(w["n"], w["topology"], w["schedule"], T, g, n_steps=w["n_steps"],
                           gamma_on=w["gamma_on"], n_threads=nthr)
        return time.perf_counter() - t0

    run(nthr)                                         # warm the thread pool / caches
    t1 = run(4 * nthr) / 4.0                          # pilot: seconds per round of nthr points
    k = int(nthr * max(1, min(256, round(seconds / max(t1, 1e-3)))))
    dt = run(k)
    return {"value": k * w["site_updates_per_point"] / dt, "unit": "site-updates/s",
            "cores": nthr, "kind": "port", "grid_points_per_sec": k / dt,
            "sample": "%d random points of the workload grid, oracle/rk4_oracle.c with "
                      "OpenMP over points, %.1f s" % (k, dt)}, k, dt


def cpu_numpy_row(w, seconds=3.0):
    """the numpy restatement of the oracle (stencil / rank-one, one
    core) on a few points of the workload."""
    from oracle import rk4_numpy as rk
    t_arr, b_arr = w["time_array"], w["beta_array"]
    rng = np.random.default_rng(6)
    k, done, dt = 1, 0, 0.0
    while dt < seconds and done < 64:
        T, g = t_arr[rng.integers(0, t_arr.size, k)], b_arr[rng.integers(0, b_arr.size, k)]
        t0 = time.perf_counter()
        rk.rk4_L2(w["n"], w["topology"], w["schedule"], T, g, n_steps=w["n_steps"],
                  gamma_on=w["gamma_on"])
        dt += time.perf_counter() - t0
        done += k
    return {"value": done * w["site_updates_per_point"] / dt, "unit": "site-updates/s",
            "cores": 1, "kind": "port-numpy",
            "sample": "%d random points of the workload grid, oracle/rk4_numpy.py (L2 stencil "
                      "restatement), one core, %.1f s" % (done, dt)}


def literal_reference_row(w):
    """SURVEY.md 8(d) row B1.  The unmodified reference (evaluate_probability, BCB.py:193-203:
    SciPy RK45 over a dense H(t) rebuilt in Python loops on every right-hand side) cannot run
    on the GPU box -- /root/reference does not travel -- so this row is an EXTRAPOLATION from
    the survey's measurement in the build container: one RHS call
    costs 0.66 us x N^2 on one core, and an RK4 step needs four of them."""
    n = w["n"]
    per_rhs = 0.66e-6 * n * n
    return {"value": 1.0 / (4.0 * per_rhs / n), "unit": "site-updates/s", "cores": 1,
            "kind": "literal-reference-extrapolated",
            "sample": "not run here: 0.66 us x N^2 per schrodinger_equation call (measured on "
                      "the unmodified reference in the build container, BASELINE.md section 2) "
                      "x 4 calls per RK4 step, N = %d; the literal reference at this size is "
                      "infeasible (%.2f s per right-hand side)" % (n, per_rhs)}


def run_reference(args):
    """CPU arm: rank 0 only.  A throughput measurement: every step integrates a bounded random
    sample of the workload's own (T, gamma) points, not the whole grid."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    w = _workload(args.workload, args.gpus, args.rows_per_gpu, args.scaling)
    per_step = max(1.0, min(20.0, 120.0 / max(1, args.steps + args.warmup)))
    for _ in range(args.warmup):
        cpu_port_rate(w, per_step * 0.25)
    tot_pts, tot_t, info = 0, 0.0, None
    for _ in range(args.steps):
        info, k, dt = cpu_port_rate(w, per_step)
        tot_pts += k
        tot_t += dt
    value = tot_pts * w["site_updates_per_point"] / tot_t
    info["value"] = value
    info["sample"] = ("%d random points of the workload grid per step (%d in %d steps, %.1f s), "
                      "oracle/rk4_oracle.c with OpenMP over points: a bounded sample, not the "
                      "whole grid" % (tot_pts // max(1, args.steps), tot_pts, args.steps, tot_t))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "site-updates/s",
            "grid_points_per_sec": tot_pts / tot_t, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / max(1, args.steps),
            "higher_is_better": True, "scaling": w["scaling"], "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": _config(w, args.gpus),
            "cpu_baseline": info,
            "e2e": {"value": value, "unit": "site-updates/s", "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def run_b200(args):
    import torch
    import torch.distributed as dist
    import qwtdh_b200 as qw

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d (launch with torch.distributed.run)"
                         % (args.gpus, world))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # NCCL prints its banner / debug lines on stdout; stdout carries the ONE JSON line
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)

    w = _workload(args.workload, world, args.rows_per_gpu, args.scaling)
    prob = qw.Problem(n=w["n"], topology=w["topology"], schedule=w["schedule"],
                      gamma_on=w["gamma_on"], n_steps=w["n_steps"])
    rows = qw.shard_rows(w["beta_array"].size, world, rank)
    pts_local = (rows[1] - rows[0]) * w["n_time"]
    pts_total = w["beta_array"].size * w["n_time"]
    d_time = torch.from_numpy(w["time_array"]).to(dev)
    d_beta = torch.from_numpy(w["beta_array"]).to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # FP64 roofline denominator, measured here (MEASURED_PEAKS.json has no FP64 entry)
    probe_clock = ClockSampler(local)
    probe_clock.start()
    fp64_peak = qw.fp64_peak_tflops(600.0)
    probe_clock = probe_clock.summary()
    plan = qw.plan(prob)

    # ---- device-resident: K steps, CUDA events per step on the launching stream ---------
    for _ in range(args.warmup):
        flush.zero_()
        qw.rk4_heatmap(prob, d_time, d_beta, rows=rows)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = qw.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
          for _ in range(args.steps)]
    for a, b in ev:
        flush.zero_()                                 # L2 flush between timed iterations
        a.record()
        p_dev, n_dev = qw.rk4_heatmap(prob, d_time, d_beta, rows=rows)
        b.record()
    barrier()
    launches = qw.launch_count() - launches0
    clocks = sampler.summary()
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    dev_ms = max_over_ranks(dev_ms)
    ms_per_step = dev_ms / args.steps
    value = pts_total * w["site_updates_per_point"] / (ms_per_step * 1e-3)

    # ---- end to end: pinned host inputs -> public API -> host heatmap --------------------
    h_time = torch.from_numpy(w["time_array"]).pin_memory()
    h_beta = torch.from_numpy(w["beta_array"]).pin_memory()

    def e2e_step():
        if world == 1:
            p, nrm = qw.rk4_heatmap(prob, h_time, h_beta, rows=rows)
        else:                       # one all_gather_into_tensor, one D2H on rank 0
            p, nrm = qw.heatmap_sharded(prob, h_time.numpy(), h_beta.numpy(), dst=0)
        return p, nrm

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        p_host, n_host = e2e_step()
    barrier()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3) / args.steps
    e2e_value = pts_total * w["site_updates_per_point"] / (e2e_ms * 1e-3)
    # bytes over PCIe per step, summed over the ranks: every rank uploads both axes; the
    # heatmap (p and norm, padded to equal row blocks) comes down once, on rank 0
    h2d = world * (h_time.numel() + h_beta.numel()) * 8
    d2h = sum(qw.sharded_d2h_bytes(w["beta_array"].size, w["n_time"], world, r, dst=0)
              for r in range(world))

    ok = True
    if rank == 0:
        ok = bool(np.isfinite(p_host).all() and (p_host >= 0).all()
                  and (p_host <= 1 + 1e-9).all()
                  and np.array_equal(p_host[rows[0]:rows[1]] if world > 1 else p_host,
                                     p_dev.cpu().numpy()))

    if rank == 0:
        kernel_s = ms_per_step * 1e-3                 # one launch per step and per GPU
        site_updates_local = pts_local * w["site_updates_per_point"]
        achieved = FLOPS_PER_SITE_UPDATE * site_updates_local / kernel_s / 1e12
        ctr = kernel_counters(plan)
        if "fp64_thread_inst_per_site_update" in ctr:
            ex, ex_src = ctr["fp64_thread_inst_per_site_update"], ctr["counters_source"]
        else:
            ex, ex_src = source_fp64_per_site_update(plan), (
                "counted in the kernel source / SASS (tools/sass_loops.py); no ncu opcode "
                "counters committed for this kernel")
        traffic, traffic_note = None, ("no ncu --set full capture of this kernel at this launch "
                                       "size in profiles/kernel_counters.json")
        if ctr.get("dram_launch_points") == pts_local:
            traffic, traffic_note = ctr["dram_bytes_per_launch"], ctr["dram_source"]
        lane_rate = plan["sm_count"] * 64 * (clocks["sm_mhz"] or 1965.0) * 1e6    # FP64 lanes/s
        pipe_util = ex * site_updates_local / kernel_s / lane_rate
        line = {
            "metric": METRIC, "value": value, "unit": "site-updates/s",
            "grid_points_per_sec": pts_total / (ms_per_step * 1e-3),
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": w["scaling"],
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": _config(w, world),
            "kernel": plan, "points_per_gpu": pts_local,
            "l2": "flushed (256 MiB write) between timed iterations; state is register-resident",
            "results_sane_and_e2e_equals_device": ok,
            "roofline": {"bound": "fp64", "achieved": achieved, "peak": fp64_peak,
                         "unit": "TFLOP/s", "frac": achieved / fp64_peak,
                         "traffic": traffic, "traffic_note": traffic_note,
                         "algorithmic_bytes_per_launch": 16 * pts_local + 8 * (
                             w["n_time"] + int(w["beta_array"].size)),
                         "peak_source": "qw_fp64_peak_probe: dependent-free DFMA stream "
                                        "measured in this process, best of 8/16/24 warps per "
                                        "SM (MEASURED_PEAKS.json has no FP64 entry)",
                         "peak_nominal": 148 * 64 * 2 * 1.965e-3,
                         "frac_of_nominal": achieved / (148 * 64 * 2 * 1.965e-3),
                         "probe_clocks": probe_clock,
                         "algorithmic_flops_per_site_update": FLOPS_PER_SITE_UPDATE,
                         "executed_fp64_instructions_per_site_update": ex,
                         "executed_fp64_source": ex_src,
                         "fp64_pipe_utilisation": pipe_util,
                         "note": "achieved counts the ALGORITHMIC 52 flops of classic RK4 per "
                                 "site-update (SURVEY.md 8d); the nested-form kernels issue "
                                 "fewer FP64 instructions for the same step, so frac = 1 is "
                                 "not the ceiling: frac <= 52 / (2 * executed instructions)",
                         "max_frac_of_fma_peak": 52.0 / (2.0 * ex)},
            "e2e": {"value": e2e_value, "unit": "site-updates/s", "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "gather": "none (one GPU)" if world == 1 else
                    "all_gather_into_tensor over NCCL, one device-to-host copy on rank 0",
                    "grid_points_per_sec": pts_total / (e2e_ms * 1e-3)},
            "gpu_launches": launches, "clocks": clocks,
        }
        if not args.no_mirror:
            line["symmetry_reduced"] = mirror_measure(qw, torch, w, d_time, d_beta, rows, flush,
                                                      pts_local, p_dev)
        if world == 1 and not args.no_stream:
            line["streaming"] = stream_measure(qw, torch, dev)
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_port_rate(w, args.cpu_seconds)[0]
            line["cpu_baseline"]["other_rows"] = [cpu_numpy_row(w), literal_reference_row(w)]
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c5", choices=sorted(WORKLOADS))
    ap.add_argument("--rows-per-gpu", type=int, default=None)
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: fixed rows per GPU (default); strong: the workload's whole grid "
                         "split over the GPUs (c5: 1024 x 1024 at every N)")
    ap.add_argument("--no-mirror", action="store_true",
                    help="skip the side measurement of the symmetry-reduced kernels")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-stream", action="store_true",
                    help="skip the N=65536 HBM-streaming side measurement")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        print("note: warmup < 3 does not meet the timing rules", file=sys.stderr)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
