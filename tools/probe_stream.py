#!/usr/bin/env python
"""Phase timing of the bulk-copy streaming kernel (library built with -DQW_PROBE):
    python tools/probe_stream.py [kb]"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qwtdh_b200 as qw  # noqa: E402

kb = int(sys.argv[1]) if len(sys.argv) > 1 else 8
flag = {1: 8, 2: 16, 4: 24, 8: 0}[kb]
lib = qw._lib.load()
dev = torch.device("cuda")
T = torch.linspace(0.1, 64.0, 1024, dtype=torch.float64, device=dev)
g = torch.linspace(0.0, 2.5, 1024, dtype=torch.float64, device=dev)
prob = qw.Problem(n=65536, topology="circular", schedule="lin", n_steps=64, flags=flag)
qw.rk4_batch(prob, T, g)
torch.cuda.synchronize()
buf = (ctypes.c_ulonglong * 8)()
lib.qw_debug_probe(buf, 1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
qw.rk4_batch(prob, T, g)
e1.record()
torch.cuda.synchronize()
lib.qw_debug_probe(buf, 1)
v = list(buf)
n = max(1, v[5])
names = ["mbarrier wait", "LDS + mailbox + barrier", "steps", "stage + fence + barrier",
         "store / load issue + index", "windows"]
print("kb=%d: %.2f ms, %d windows (thread 0 of every block)" % (kb, e0.elapsed_time(e1), n))
tot = sum(v[:5])
for i in range(5):
    print("  %-28s %9.0f cycles/window  %5.1f %%" % (names[i], v[i] / n, 100.0 * v[i] / tot))

import numpy as np
bb = (ctypes.c_ulonglong * (3 * 1024))()
lib.qw_debug_probe_blocks(bb)
a = np.array(list(bb), dtype=np.int64).reshape(1024, 3)[:592]
t0 = a[:, 1].min()
start, end = (a[:, 1] - t0) / 1e3, (a[:, 2] - t0) / 1e3
print("last pass: block start us  min %.1f  median %.1f  max %.1f" % (start.min(), np.median(start), start.max()))
print("           block end   us  min %.1f  median %.1f  max %.1f" % (end.min(), np.median(end), end.max()))
late = start > 20.0
print("           blocks starting later than 20 us: %d; blocks per SM: %s" % (
    late.sum(), np.bincount(np.bincount(a[:, 0].astype(int), minlength=148))))
dur = end - start
print("           duration us: min %.1f median %.1f max %.1f" % (dur.min(), np.median(dur), dur.max()))
