"""CPU: host-side logic of the drop-in layer that needs no GPU -- dense-matrix helpers
against the reference's own matrices, the robustness arithmetic, row sharding, and the
world_size-2 gather (gloo) with the oracle injected as the compute function."""

import importlib
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

PKG = "quantum-walks-time-dependent-hamiltonians_b200"


def _mod(name):
    return importlib.import_module(PKG + "." + name)


def test_generate_hamiltonian_matches_reference_matrices(golden):
    meta, arr = golden
    bcb, ss = _mod("BCB"), _mod("Schroedinger_Solver")
    for c in meta["hamiltonians"]:
        if c["src"] == "BCB":
            bcb.dimension, bcb.topology, bcb.step_function = c["n"], c["topology"], c["schedule"]
            H = bcb.generate_hamiltonian(c["n"], c["beta"], c["t"], c["T"])
        else:
            ss.dimension = c["n"]
            ss.step_function = {"lin": 1, "sqrt": 2, "cbrt": 3}[c["schedule"]]
            H = ss.generate_hamiltonian(c["n"], c["beta"], c["t"], c["T"])
        assert np.abs(H - arr[c["key"]]).max() <= 1e-15 * max(1, np.abs(arr[c["key"]]).max()), c


def test_time_independent_hamiltonian_and_qw_search_classes():
    bcb, qws = _mod("BCB"), _mod("QW_Search")
    H = bcb.generate_time_independent_hamiltonian(7, 1.5)
    assert H[3, 3] == 2 - 1.5 and H[0, 6] == -1 and H[0, 1] == -1 and H.sum() == -1.5
    HC = qws.Hamiltonian(5, 0, 0, 1, 0, 10)
    HC.build_laplacian()
    HC.build_hamiltonian()
    assert HC.L[1, 1] == 4 and HC.H[1, 1] == 4 - 10 and HC.L[0, 1] == -1   # L not corrupted
    HT = qws.Hamiltonian(6, 1, 1, 2, 1, 0.7)
    HT.build_laplacian()
    HT.build_hamiltonian()
    HT.update_hamiltonian(1.0, 4.0)
    assert HT.H[2, 2] == pytest.approx(0.75 * 2 - 0.25 * 0.7) and HT.H[0, 5] == -0.75
    psi = qws.State(6)
    psi.set_initial(0)
    psi.set_hamiltonian(HT)
    assert psi.bra().shape == (6,) and np.allclose(np.vdot(psi.ket, psi.ket), 1)


def test_robustness_arithmetic(golden, tmp_path, monkeypatch):
    meta, arr = golden
    rob = _mod("Robustness")
    for c in meta["robustness"]:
        assert rob.robustness(arr["rob_prob"], c["beta_index"], c["time_index"],
                              c["variation"]) == c["R"]
    monkeypatch.chdir(tmp_path)
    np.save("9_probability_lin.npy", arr["rob_prob"])
    np.save("9_circ_time.npy", np.arange(5.0))
    np.save("9_circ_beta.npy", np.arange(12.0))
    p, t, b = rob.load_data_file(9, 1)
    assert np.array_equal(p, arr["rob_prob"]) and t.size == 5 and b.size == 12


def test_production_beta_list():
    bl = _mod("BCB_latest")
    assert len(bl.PRODUCTION_BETA) == 75 and len(bl.PRODUCTION_BETA) % 3 == 0
    assert bl.PRODUCTION_BETA[0] == 0.1 and bl.PRODUCTION_BETA[-1] == 3.95
    assert 1.05 not in bl.PRODUCTION_BETA and 1.1 in bl.PRODUCTION_BETA


def test_shard_rows_partitions_exactly(qw):
    for n in (0, 1, 7, 24, 75, 1024):
        for world in (1, 2, 3, 4, 8):
            blocks = [qw.shard_rows(n, world, r) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1
    assert [qw.shard_rows(24, 4, r) for r in range(4)] == [(0, 6), (6, 12), (12, 18), (18, 24)]


def test_problem_validation_on_host(qw):
    with pytest.raises(ValueError):
        qw.Problem(n=8).c_struct()
    with pytest.raises(ValueError):
        qw.Problem(n=8, n_steps=4, step_size=0.1).c_struct()
    with pytest.raises(ValueError):
        qw.Problem(n=8, schedule="quartic", n_steps=4).c_struct()
    with pytest.raises(ValueError):
        qw.Problem(n=8, topology="star", n_steps=4).c_struct()
    s = qw.Problem(n=8, topology="complete", schedule="independent", gamma_on="laplacian",
                   step_size=1e-3).c_struct()
    assert (s.topology, s.schedule, s.gamma_on, s.target, s.step_size) == (0, 6, 1, -1, 1e-3)
    assert list(qw.longest_first_order([3, 9, 9, 1])) == [1, 2, 0, 3]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _oracle_compute(problem, time_array, beta_array, rows):
    from oracle import c_oracle
    return c_oracle.heatmap(problem.n, problem.topology, problem.schedule, time_array,
                            beta_array[rows[0]:rows[1]], n_steps=problem.n_steps, n_threads=2)


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    qw = importlib.import_module(PKG)
    prob = qw.Problem(n=12, topology="circular", schedule="cerf_3", n_steps=64)
    t, b = np.linspace(0.1, 12, 5), np.linspace(0, 2.5, 7)       # 7 rows: uneven blocks
    p, nrm = qw.heatmap_sharded(prob, t, b, compute=_oracle_compute)
    mine, _ = qw.heatmap_sharded(prob, t, b, compute=_oracle_compute, gather=False)
    np.save(os.path.join(out_dir, "p%d.npy" % rank), p)
    np.save(os.path.join(out_dir, "n%d.npy" % rank), nrm)
    np.save(os.path.join(out_dir, "m%d.npy" % rank), mine)
    dist.destroy_process_group()


def test_sharded_heatmap_gloo_world_size_2(tmp_path):
    from oracle import c_oracle
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    t, b = np.linspace(0.1, 12, 5), np.linspace(0, 2.5, 7)
    ref_p, ref_n = c_oracle.heatmap(12, "circular", "cerf_3", t, b, n_steps=64)
    for r in range(world):
        assert np.array_equal(np.load(tmp_path / ("p%d.npy" % r)), ref_p)     # every rank
        assert np.array_equal(np.load(tmp_path / ("n%d.npy" % r)), ref_n)
    assert np.array_equal(np.load(tmp_path / "m0.npy"), ref_p[:4])            # 4 + 3 rows
    assert np.array_equal(np.load(tmp_path / "m1.npy"), ref_p[4:])
