#!/usr/bin/env python
"""Tiny run of every nested-form kernel for compute-sanitizer (memcheck / racecheck):
    compute-sanitizer --tool racecheck python tools/sanitize_run.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qwtdh_b200 as qw  # noqa: E402

T = np.array([3.0, 5.5, 0.0, 7.25])
g = np.array([1.1, 0.4, 2.0, 0.0])
for kw in (dict(n=1024, topology="circular", n_steps=70),                 # cycle, 2 warps
           dict(n=2048, topology="circular", n_steps=40),                 # cycle, 4 warps
           dict(n=264, topology="circular", step_size=0.1),               # idle threads
           dict(n=1024, topology="complete", n_steps=100),                # folded, 2 warps, 4 chunks
           dict(n=512, topology="complete", n_steps=100),                 # folded, 1 warp, one map buffer
           dict(n=2048, topology="complete", n_steps=70),                 # folded, 4 warps
           dict(n=1000, topology="complete", n_steps=70),                 # folded, ragged
           dict(n=1024, topology="complete", n_steps=100, flags=128),     # one auxiliary warp
           dict(n=4096, topology="complete", n_steps=70),                 # scalar warp + table warp
           dict(n=3000, topology="complete", n_steps=70),                 # same, ragged
           dict(n=4096, topology="complete", n_steps=70, flags=128),
           dict(n=8192, topology="circular", n_steps=9, flags=0),         # stream, nested
           dict(n=8192, topology="circular", n_steps=5, flags=8)):
    scale = 1.0 if kw["topology"] == "circular" else 1.0 / kw["n"]
    p, nrm = qw.rk4_batch(qw.Problem(schedule="cerf_3", **kw), T * scale, g / scale)
    print(qw.plan(qw.Problem(schedule="cerf_3", **kw))["path"], kw["n"], p)
Tb = np.linspace(1.0, 6.0, 700)
p, nrm = qw.rk4_batch(qw.Problem(n=256, n_steps=20), Tb, np.full(700, 1.3))   # packed nested
print(qw.plan(qw.Problem(n=256, n_steps=20))["path"], p[:3])
gap, cpl, e0 = qw.adiabatic_batch(qw.Problem(n=51, n_steps=1), np.linspace(0.1, 0.9, 9), 1.0)
print("adiabatic", gap[:3])
