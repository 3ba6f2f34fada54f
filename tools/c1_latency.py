#!/usr/bin/env python
"""BASELINE config 1 (cycle N=16, one (T, gamma) point, 1024 steps): device time and end-to-end
latency through the public API (host list in -> numpy out), next to one host core of the C oracle.

    python tools/c1_latency.py
"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qwtdh_b200 as qw  # noqa: E402
from oracle import c_oracle  # noqa: E402  (the CPU side of the comparison)


def device_us(prob, reps=200):
    T = torch.tensor([16.0], dtype=torch.float64, device="cuda")
    g = torch.tensor([1.0], dtype=torch.float64, device="cuda")
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        p, _ = qw.rk4_batch(prob, T, g)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best * 1e3, float(p[0])


def host_us(fn, reps=300):
    fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return 1e6 * float(np.median(ts)), 1e6 * float(np.min(ts))


ref = 0.42038739054526
for name, kw in (("n_steps=1024", dict(n_steps=1024)), ("step_size=1e-3 (16000 steps)", dict(step_size=1e-3))):
    for flags, label in ((0, "default"), (512, "warp-packed kernel (QW_FLAG_NO_PACK off PIT)")):
        prob = qw.Problem(n=16, topology="circular", schedule="lin", flags=flags, **kw)
        d_us, p_dev = device_us(prob)
        med, mn = host_us(lambda: qw.rk4_batch(prob, [16.0], [1.0]))
        p_api = float(qw.rk4_batch(prob, [16.0], [1.0])[0][0])
        print("%s, %s [%s]: device %.1f us; API host->host median %.1f us (min %.1f); p = %.14f"
              % (name, label, qw.plan(prob)["path"], d_us, med, mn, p_api))
    okw = dict(n_steps=1024) if "n_steps" in kw else dict(step_size=1e-3)
    med, mn = host_us(lambda: c_oracle.rk4_batch(16, "circular", "lin", [16.0], [1.0], n_threads=1, **okw),
                      reps=50)
    print("%s, C oracle on one host core incl. ctypes: median %.1f us (min %.1f)" % (name, med, mn))
print("reference value (SURVEY.md Appendix A): %.14f" % ref)
