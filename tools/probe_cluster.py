#!/usr/bin/env python
"""Where the cluster kernel waits (library built with -DQW_PROBE): python tools/probe_cluster.py [n]"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qwtdh_b200 as qw  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
lib = qw._lib.load()
dev = torch.device("cuda")
pts = 592
T = torch.linspace(0.1, 50.0, pts, dtype=torch.float64, device=dev)
g = torch.linspace(0.2, 2.5, pts, dtype=torch.float64, device=dev)
prob = qw.Problem(n=n, topology="circular", schedule="cerf_3", n_steps=256)
qw.rk4_batch(prob, T, g)
torch.cuda.synchronize()
buf = (ctypes.c_ulonglong * 4)()
lib.qw_debug_cluster_probe(buf)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
qw.rk4_batch(prob, T, g)
e1.record()
torch.cuda.synchronize()
lib.qw_debug_cluster_probe(buf)
w, lv, loop, steps = list(buf)
print("n=%d: %.2f ms; left-edge wait %.0f cycles per level (%d levels); loop %.0f cycles per step "
      "(thread 0 of every CTA)" % (n, e0.elapsed_time(e1), w / max(1, lv), lv, loop / max(1, steps)))
