#!/usr/bin/env python
"""Multi-GPU check of the drop-in parallel_routine (BCB.py:262-315) over NCCL:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29533 tests/run_parallel_routine.py

Every rank integrates its block of beta rows, one all_gather, rank 0 writes the reference's
.npy files; rank 0 then recomputes the whole heatmap on its own GPU and with the C oracle and
compares.  Lives under tests/ because it uses the oracle as the checker."""
import importlib
import os
import sys
import tempfile

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qwtdh_b200 as qw  # noqa: E402
from oracle import c_oracle  # noqa: E402  (checker only)


def main():
    rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    bcb = importlib.import_module(qw.__name__ + ".BCB")
    tmp = tempfile.mkdtemp() if rank == 0 else None
    box = [tmp]
    dist.broadcast_object_list(box, src=0)
    os.chdir(box[0])
    ok = True
    # 256: the half ring with four points per warp (rk4_cycle_mirror_packed, column mode)
    for dim, sched, topo in ((51, "lin", "circular"), (64, "cerf_3", "circular"),
                             (256, "cerf_3", "circular"), (1024, "sqrt", "circular"),
                             (300, "lin", "complete")):
        bcb.dimension, bcb.step_function, bcb.topology = dim, sched, topo
        bcb.type_of_qw, bcb.n_steps, bcb.step_size = "dependent", None, 0.02
        ub_t, ub_b = (12.0, 2.0) if topo == "circular" else (12.0 / dim, 2.0 * dim)
        bcb.parallel_routine(0.1 * ub_t / 12.0, ub_t, 0, ub_b)
        dist.barrier()
        if rank == 0:
            p = np.load("%d_%s_probability_%s.npy" % (dim, topo, sched))
            t = np.load("%d_circ_time.npy" % dim)
            b = np.load("%d_circ_beta.npy" % dim)
            prob = qw.Problem(n=dim, topology=topo, schedule=sched, step_size=0.02,
                              flags=8192 | 32768)      # the drop-in's flags: mirror, fixed geometry
            p1, _ = qw.rk4_heatmap(prob, t, b)
            po, _ = c_oracle.heatmap(dim, topo, sched, t, b, step_size=0.02)
            e1, e2 = float(np.abs(p - p1).max()), float(np.abs(p - po).max())
            good = p.shape == (24, 24) and e1 == 0.0 and e2 <= 1e-10   # bitwise the one-GPU heatmap
            ok = ok and good
            print("N=%d %s %s: shape %s, |dp| vs one GPU %.1e, vs oracle %.1e -> %s"
                  % (dim, topo, sched, p.shape, e1, e2, "ok" if good else "FAIL"), flush=True)
    # the reference's shipped integrator (SciPy RK45 semantics) and the time-independent
    # comparator over the same process group: the shards stay on the GPUs through the collective
    import numpy as _np
    for kind, integ, ti in (("dependent", "rk45", "exact"), ("independent", "rk4", "exact"),
                            ("independent", "rk4", "rk4")):
        bcb.dimension, bcb.step_function, bcb.topology = 51, "lin", "circular"
        bcb.type_of_qw, bcb.integrator, bcb.ti_method = kind, integ, ti
        bcb.n_steps, bcb.step_size = None, 0.02
        bcb.parallel_routine(0.5, 12.0, 0, 2.0)
        dist.barrier()
        if rank == 0:
            name = "51_circular_probability_%s.npy" % ("lin" if kind == "dependent" else "static")
            p = _np.load(name)
            t, b = _np.load("51_circ_time.npy"), _np.load("51_circ_beta.npy")
            if kind == "dependent":
                p1, _ = qw.rk45_heatmap(qw.Problem(n=51, topology="circular", schedule="lin"), t, b)
            elif ti == "exact":
                p1, _ = qw.ti_exact_heatmap(qw.Problem(n=51, topology="circular", schedule="static"), t, b)
            else:
                p1, _ = qw.rk4_heatmap(qw.Problem(n=51, topology="circular", schedule="static",
                                                  step_size=0.02, flags=8192 | 32768), t, b)
            e1 = float(_np.abs(p - p1).max())
            good = p.shape == (24, 24) and e1 <= 1e-12
            ok = ok and good
            print("N=51 %s / %s / %s: shape %s, |dp| vs one GPU %.1e -> %s"
                  % (kind, integ, ti, p.shape, e1, "ok" if good else "FAIL"), flush=True)
        dist.barrier()
    bcb.type_of_qw, bcb.integrator, bcb.ti_method = "dependent", "rk4", "exact"
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("parallel_routine over %s ranks: %s" % (os.environ["WORLD_SIZE"],
                                                        "ALL OK" if ok else "FAILED"))
        sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
