#!/usr/bin/env python
"""Run every BASELINE.json config at full size on one GPU and print a markdown table
(SURVEY.md section 8d): wall time (CUDA events), site-updates/s, grid-points/s, the physics
read-outs (best cell, tau, I, robustness) and a spot check against the C oracle (the parity
tests proper -- every cell of C2, all 1e5 realisations of C4, strided sub-grids of C3 / C5 -- are
tests/test_gpu_fullsize.py, part of pytest -m gpu).

    python tests/run_configs.py [--skip-long] > profiles/rNN_configs.md

Lives under tests/ because it uses the oracle as the checker.
"""
import argparse
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qwtdh_b200 as qw  # noqa: E402
from oracle import c_oracle  # noqa: E402  (checker only)

DEV = torch.device("cuda")


def timed(fn, reps=1):
    fn()                                           # warm-up (module load, pool)
    best, out = 1e30, None
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best, out


def spot_check(prob_kw, t, b, p, k=6, seed=1):
    rng = np.random.default_rng(seed)
    jj, ii = rng.integers(0, len(b), k), rng.integers(0, len(t), k)
    po, _, _ = c_oracle.rk4_batch(prob_kw["n"], prob_kw["topology"], prob_kw["schedule"],
                                  t[ii], b[jj], n_steps=prob_kw["n_steps"],
                                  gamma_on=prob_kw.get("gamma_on", "oracle"))
    return float(np.abs(p[jj, ii] - po).max())


def heatmap_row(name, kw, t, b, reps=1, check=True):
    prob = qw.Problem(**kw)
    dt, db = torch.from_numpy(t).to(DEV), torch.from_numpy(b).to(DEV)
    ms, (p, nrm) = timed(lambda: qw.rk4_heatmap(prob, dt, db), reps)
    pts = len(t) * len(b)
    su = pts * kw["n"] * kw["n_steps"]
    summ = qw.heatmap_summary(p, dt, t_min=(np.pi / 2) * np.sqrt(kw["n"]))
    ph = p.cpu().numpy()
    err = spot_check(kw, t, b, ph) if check else float("nan")
    pl = qw.plan(prob)
    j, i = summ["argmax"]
    tj, ti = summ["tau_argmin"]
    print("| %s | %s R=%d x %d thr | %d | %.1f | %.3e | %.3e | p_max=%.6f at (gamma=%.4g, T=%.4g), "
          "I=%.2f, tau_min=%.4g at (gamma=%.4g, T=%.4g), norm defect max %.1e | %.1e |"
          % (name, pl["path"], pl["sites_per_thread"], pl["threads_per_block"], pts, ms,
             su / (ms * 1e-3), pts / (ms * 1e-3), summ["p_max"], b[j], t[i], summ["iterations"],
             summ["tau_min"], b[tj] if tj >= 0 else float("nan"),
             t[ti] if ti >= 0 else float("nan"),
             float((nrm - 1).abs().max()), err), flush=True)
    return ph


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--skip-long", action="store_true")
    a = ap.parse_args()
    print("# BASELINE configs at full size, one B200 (%s)" % torch.cuda.get_device_name(0))
    print()
    print("Device-resident inputs, CUDA events, one warm-up run; spot check = max |dp| against the C "
          "oracle on 6 random cells of the same grid (full-size parity: tests/test_gpu_fullsize.py). "
          "Rows marked 'half ring' / 'pair' use the exact symmetry reductions the drop-in modules "
          "default to (QW_FLAG_MIRROR); their site-updates/s count FULL-ring site-updates. "
          "FP64 peak measured: %.2f TFLOP/s."
          % qw.fp64_peak_tflops(300.0))
    print()
    print("| config | kernel | points | ms | site-updates/s | grid-points/s | read-out | spot check |")
    print("|---|---|---|---|---|---|---|---|")

    # C1: cycle N=16, single point, n_steps 1024
    prob = qw.Problem(n=16, topology="circular", schedule="lin", n_steps=1024)
    T1, g1 = torch.tensor([16.0], dtype=torch.float64, device=DEV), torch.tensor(
        [1.0], dtype=torch.float64, device=DEV)
    ms, (p, nrm) = timed(lambda: qw.rk4_batch(prob, T1, g1), reps=5)
    host = []
    for _ in range(101):                           # the first call creates the staging buffers
        t0 = time.perf_counter()
        ph, _ = qw.rk4_batch(prob, [16.0], [1.0])
        host.append((time.perf_counter() - t0) * 1e6)
    host_us = float(np.median(host[1:]))
    print("| C1 cycle N=16 (T,gamma)=(16,1) lin S=1024 | parallel-in-time: 23 chunk propagators "
          "+ combine | 1 | %.3f | "
          "%.3e | %.3e | p=%.14f (reference RK4 n=1024: 0.42038739054526); %.0f us per point "
          "device, %.0f us through the host API (median of 100 calls) | %.1e |"
          % (ms, 16 * 1024 / (ms * 1e-3), 1 / (ms * 1e-3), ph[0], ms * 1e3, host_us,
             abs(ph[0] - 0.4203873905452587)), flush=True)

    # C2: cycle N=64, 100x100, lin / cerf_3 / static
    t, b = np.linspace(0.1, 64, 100), np.linspace(0, 2.5, 100)
    for sched in ("lin", "cerf_3", "static"):
        heatmap_row("C2 cycle N=64 100x100 %s S=1024" % sched,
                    dict(n=64, topology="circular", schedule=sched, n_steps=1024), t, b, reps=3)

    # C2, time-independent side with the exact propagator (the reference's expm comparator)
    prob = qw.Problem(n=64, topology="circular", schedule="static")
    dt, db = torch.from_numpy(t).to(DEV), torch.from_numpy(b).to(DEV)
    ms, (p, nrm) = timed(lambda: qw.ti_exact_heatmap(prob, dt, db), 3)
    summ = qw.heatmap_summary(p, dt, t_min=(np.pi / 2) * 8.0)
    from oracle import rk4_numpy as rk
    jj, ii = [5, 50, 99], [7, 60, 99]
    err = max(abs(float(p[j, i]) - rk.ti_exact(64, t[i], b[j])[0]) for j, i in zip(jj, ii))
    print("| C2 cycle N=64 100x100 time-independent, exact Chebyshev propagator | cheb_cycle R=8 x "
          "32 thr | 10000 | %.2f | n/a | %.3e | p_max=%.6f at (gamma=%.4g, T=%.4g), I=%.2f, "
          "tau_min=%.4g, norm defect max %.1e | %.1e (vs dense diagonalisation) |"
          % (ms, 1e4 / (ms * 1e-3), summ["p_max"], b[summ["argmax"][0]], t[summ["argmax"][1]],
             summ["iterations"], summ["tau_min"], float((nrm - 1).abs().max()), err), flush=True)

    # the thesis' production sweep, BCB_latest.py:309-320: 4 schedules x odd N in [3, 71] x
    # (75 beta x 4 T), then Robustness.py over the files it wrote
    import importlib
    import tempfile
    bl = importlib.import_module(qw.__name__ + ".BCB_latest")
    rob = importlib.import_module(qw.__name__ + ".Robustness")
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:
        os.chdir(tmp)
        try:
            for label, kw in (("RK4 step_size=1e-3 (QW_Search.py:153)", dict(integrator="rk4")),
                              ("RK45 rtol=atol=1e-6 (as shipped)", dict(integrator="rk45"))):
                bl.integrator, bl.n_steps, bl.step_size = kw["integrator"], None, 1e-3
                bl.type_of_qw = "dependent"
                import contextlib
                import io
                t0 = time.perf_counter()
                with contextlib.redirect_stdout(io.StringIO()):
                    bl.production_sweep()
                    torch.cuda.synchronize()
                    for dim in range(7, 72, 2):
                        rob.routine(dim, 1)
                wall = time.perf_counter() - t0
                nfiles = len([f for f in os.listdir(tmp) if f.endswith(".npy")])
                print("| thesis production sweep (BCB_latest.py:309-320 + Robustness.py:56-66), %s | "
                      "drop-in modules, 140 heatmaps of 75x4 | 42000 | %.0f | n/a | %.3e | %d .npy "
                      "files written and read back by Robustness.routine | see tests |"
                      % (label, wall * 1e3, 42000 / wall, nfiles), flush=True)
        finally:
            os.chdir(cwd)
            bl.integrator, bl.step_size = "rk4", 1e-3

    # C3: complete N=4096, 256x256, gamma on the Laplacian
    n = 4096
    t, b = np.linspace(0.1, 512, 256), np.linspace(0.5 / n, 2.5 / n, 256)
    for sched in ("cerf_3", "lin"):
        heatmap_row("C3 complete N=4096 256x256 %s S=2048 gamma-on-L" % sched,
                    dict(n=n, topology="complete", schedule=sched, gamma_on="laplacian",
                         n_steps=2048), t, b)
    heatmap_row("C3 complete N=4096 256x256 cerf_3 S=2048 gamma-on-L, (psi_w, psi_other) pair "
                "(drop-in default)",
                dict(n=n, topology="complete", schedule="cerf_3", gamma_on="laplacian",
                     n_steps=2048, flags=8192), t, b)

    # C4: robustness ensemble, cycle N=256, 1e5 noisy realisations
    n = 256
    T0 = (np.pi / 2) * np.sqrt(n)
    gam = np.linspace(0, 2.5, 101)
    prob = qw.Problem(n=n, topology="circular", schedule="lin", n_steps=512)
    pc, _ = qw.rk4_batch(prob, np.full(101, T0), gam)
    g0, p0 = gam[int(np.argmax(pc))], float(pc.max())
    rng = np.random.default_rng(20200912)
    g = np.clip(g0 * (1 + 0.05 * rng.standard_normal(100000)), 0, None)
    T = np.clip(T0 * (1 + 0.05 * rng.standard_normal(100000)), 0, None)
    dT, dg = torch.from_numpy(T).to(DEV), torch.from_numpy(g).to(DEV)
    ms, (p, nrm) = timed(lambda: qw.rk4_batch(prob, dT, dg), reps=3)
    st = qw.ensemble_stats(p)
    po, _, _ = c_oracle.rk4_batch(n, "circular", "lin", T[:8], g[:8], n_steps=512)
    pl = qw.plan(prob)
    kname = "%s R=%d x %d thr" % (pl["path"], pl["sites_per_thread"], pl["threads_per_block"])
    print("| C4 cycle N=256 1e5 noisy (T,gamma), sigma=5%%, S=512 | " + kname + " | "
          "100000 | %.1f | %.3e | %.3e | centre (T0=%.4f, gamma0=%.3f, p0=%.6f); mean %.6f std "
          "%.6f, R=p0-mean=%.6f, frac>=0.8/0.9/0.95/0.99 = %s | %.1e |"
          % (ms, 1e5 * n * 512 / (ms * 1e-3), 1e5 / (ms * 1e-3), T0, g0, p0, st["mean"], st["std"],
             p0 - st["mean"], "/".join("%.4f" % v for v in st["fraction_ge"].values()),
             float(np.abs(p[:8].cpu().numpy() - po).max())), flush=True)
    # the same ensemble with the drop-in default: half ring, four points per warp
    probm = qw.Problem(n=n, topology="circular", schedule="lin", n_steps=512, flags=8192)
    msm, (pm, _) = timed(lambda: qw.rk4_batch(probm, dT, dg), reps=3)
    stm = qw.ensemble_stats(pm)
    plm = qw.plan(probm)
    print("| C4 cycle N=256 1e5 noisy (T,gamma), half ring (drop-in default) | %s R=%d x %d thr, "
          "%d points per warp | 100000 | %.1f | %.3e | %.3e | mean %.6f std %.6f; max |dp| vs the "
          "full-ring kernel %.1e | %.1e |"
          % (plm["path"], plm["sites_per_thread"], plm["threads_per_block"], plm["steps_per_pass"],
             msm, 1e5 * n * 512 / (msm * 1e-3), 1e5 / (msm * 1e-3), stm["mean"], stm["std"],
             float((pm - p).abs().max()), float(np.abs(pm[:8].cpu().numpy() - po).max())),
          flush=True)
    if not a.skip_long:
        # localization ensemble: T0 = N^2/2, S = 131072
        T0l = n * n / 2.0
        Tl = np.clip(T0l * (1 + 0.05 * rng.standard_normal(20000)), 0, None)
        gl = np.clip(g0 * (1 + 0.05 * rng.standard_normal(20000)), 0, None)
        probl = qw.Problem(n=n, topology="circular", schedule="lin", n_steps=131072)
        dT, dg = torch.from_numpy(Tl).to(DEV), torch.from_numpy(gl).to(DEV)
        ms, (p, nrm) = timed(lambda: qw.rk4_batch(probl, dT, dg))
        st = qw.ensemble_stats(p)
        print("| C4 localization: cycle N=256, T0=N^2/2=32768, S=131072, 2e4 realisations | "
              + kname + " | 20000 | %.1f | %.3e | %.3e | mean %.6f std %.6f, "
              "frac>=0.8/0.9/0.95/0.99 = %s | n/a |"
              % (ms, 2e4 * n * 131072 / (ms * 1e-3), 2e4 / (ms * 1e-3), st["mean"], st["std"],
                 "/".join("%.4f" % v for v in st["fraction_ge"].values())), flush=True)

    # C5: cycle N=1024, 1024x1024 (full on one GPU unless --skip-long: 128 rows)
    rows = 128 if a.skip_long else 1024
    t, b = np.linspace(0.1, 1024, 1024), np.linspace(0, 2.5, rows)
    heatmap_row("C5 cycle N=1024 %dx1024 lin S=4096" % rows,
                dict(n=1024, topology="circular", schedule="lin", n_steps=4096), t, b)
    heatmap_row("C5 cycle N=1024 %dx1024 lin S=4096, half ring (drop-in default)" % rows,
                dict(n=1024, topology="circular", schedule="lin", n_steps=4096, flags=8192), t, b)
    # C5 streaming variant
    t, b = np.linspace(0.1, 64, 32), np.linspace(0, 2.5, 32)
    for flag, kb in ((8, 1), (24, 4), (0, 8), (8192, 8)):
        heatmap_row("C5-stream cycle N=65536 32x32 lin S=256, %d step(s)/pass%s"
                    % (kb, ", half ring" if flag == 8192 else ""),
                    dict(n=65536, topology="circular", schedule="lin", n_steps=256, flags=flag),
                    t, b, reps=2, check=True)
    # above one thread block: rings across a cluster / split complete graphs
    t, b = np.linspace(0.1, 50, 64), np.linspace(0.2, 2.5, 37)
    for n in (5000, 8192):
        heatmap_row("cycle N=%d 37x64 cerf_3 S=512 (thread-block cluster)" % n,
                    dict(n=n, topology="circular", schedule="cerf_3", n_steps=512), t, b, reps=2)
    for n in (8192, 16384):
        heatmap_row("complete N=%d 37x64 cerf_3 S=512 gamma-on-L (split over blocks)" % n,
                    dict(n=n, topology="complete", schedule="cerf_3", gamma_on="laplacian",
                         n_steps=512), np.linspace(0.1, 100, 64), np.linspace(0.5 / n, 2.5 / n, 37),
                    reps=2)


if __name__ == "__main__":
    main()
