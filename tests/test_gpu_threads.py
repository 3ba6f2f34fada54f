"""Re-entrancy of the C-ABI library ("Threading"): several host threads,
each on its own CUDA stream, drive different kernels of the library at the same time.  The
results must be bitwise those of the same calls issued one after the other."""

import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CASES = [
    dict(n=64, topology="circular", schedule="lin", n_steps=300),                  # packed
    dict(n=71, topology="circular", schedule="sqrt", step_size=0.01),              # packed, any N
    dict(n=256, topology="circular", schedule="cerf_3", n_steps=200),              # packed nested
    dict(n=1024, topology="circular", schedule="lin", n_steps=150),                # nested, block
    dict(n=1000, topology="circular", schedule="cbrt", n_steps=120, flags=8192),   # mirror warp
    dict(n=2048, topology="complete", schedule="lin", gamma_on="laplacian", n_steps=200),
    dict(n=4096, topology="complete", schedule="cerf_5", gamma_on="laplacian", n_steps=300,
         flags=8192),                                                              # symmetric pair
    dict(n=16384, topology="circular", schedule="lin", n_steps=24),                # streaming
    dict(n=5001, topology="complete", schedule="lin", gamma_on="laplacian", n_steps=40),  # generic
]


def _inputs(kw, seed):
    rng = np.random.default_rng(seed)
    P = 600 if kw["n"] <= 256 else 24
    scale = 1.0 / kw["n"] if kw["topology"] == "complete" else 1.0
    T = rng.uniform(0.5, 6.0, P)
    g = rng.uniform(0.1, 2.4, P) * scale
    if kw["topology"] == "complete":
        T = T * 8.0                       # gamma-on-Laplacian: spectrum O(1), T of a few tens
    return T, g


def test_concurrent_host_threads_match_sequential(qw):
    import torch

    dev = torch.device("cuda:0")
    inputs = [_inputs(kw, 100 + i) for i, kw in enumerate(CASES)]
    seq = []
    for kw, (T, g) in zip(CASES, inputs):
        p, nrm = qw.rk4_batch(qw.Problem(**kw), T, g)
        assert np.isfinite(p).all() and (p >= 0).all() and (p <= 1 + 1e-12).all(), kw
        seq.append((p, nrm))

    rounds = 6
    results = [[None] * rounds for _ in CASES]
    errors = []
    gate = threading.Barrier(len(CASES))

    def worker(i):
        try:
            kw, (T, g) = CASES[i], inputs[i]
            stream = torch.cuda.Stream(device=dev)
            dT = torch.tensor(T, device=dev)
            dG = torch.tensor(g, device=dev)
            torch.cuda.synchronize(dev)
            gate.wait(timeout=60)
            with torch.cuda.stream(stream):
                for r in range(rounds):
                    p, nrm = qw.rk4_batch(qw.Problem(**kw), dT, dG)
                    results[i][r] = (p, nrm)
            stream.synchronize()
        except Exception as exc:                          # noqa: BLE001 (reported below)
            errors.append((i, repr(exc)))

    threads = [threading.Thread(target=worker, args=(i,)) for i in range(len(CASES))]
    for th in threads:
        th.start()
    for th in threads:
        th.join(timeout=180)
    assert not any(th.is_alive() for th in threads), "a worker thread hangs"
    assert not errors, errors
    for i, kw in enumerate(CASES):
        for r in range(rounds):
            p, nrm = results[i][r]
            assert np.array_equal(p.cpu().numpy(), seq[i][0]), (kw, r)
            assert np.array_equal(nrm.cpu().numpy(), seq[i][1]), (kw, r)


def test_last_error_is_per_thread(qw):
    """qw_last_error() is thread local: a failing call in one thread does not disturb the
    message (or the result) of another."""
    import torch

    from qwtdh_b200 import _lib

    lib = _lib.load()
    seen = {}
    try:                                  # give this thread a message of its own
        qw.rk4_batch(qw.Problem(n=5, topology="circular", n_steps=-3), [1.0], [1.0])
    except ValueError:
        pass
    mine = lib.qw_last_error().decode()
    assert "step_size" in mine or "n_steps" in mine, mine

    def bad():
        try:
            qw.rk4_batch(qw.Problem(n=2, topology="circular", n_steps=4), [1.0], [1.0])
        except ValueError as exc:
            seen["bad"] = str(exc)
        seen["bad_msg"] = lib.qw_last_error().decode()

    def good():
        p, _ = qw.rk4_batch(qw.Problem(n=16, topology="circular", n_steps=64), [2.0], [1.0])
        seen["good"] = float(p[0])

    tb, tg = threading.Thread(target=bad), threading.Thread(target=good)
    tb.start(); tb.join(60)
    tg.start(); tg.join(60)
    torch.cuda.synchronize()
    assert "cycle graph needs dimension >= 3" in seen["bad"], seen
    assert "dimension" in seen["bad_msg"]
    assert 0.0 < seen["good"] < 1.0
    assert lib.qw_last_error().decode() == mine           # untouched by the other threads' errors


def test_two_devices_in_one_process(qw):
    """One process driving two GPUs (the `device=` argument): kernels that need more than 48 KB
    of shared memory are configured per device (cudaFuncSetAttribute is per device).  Same
    inputs on cuda:0 and cuda:1 -> bitwise the same results."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    for i, kw in enumerate(CASES):
        T, g = _inputs(kw, 300 + i)
        res = [qw.rk4_batch(qw.Problem(**kw), T, g, device=d) for d in (0, 1, 0)]
        for r in res[1:]:
            assert np.array_equal(r[0], res[0][0]) and np.array_equal(r[1], res[0][1]), kw
    T, g = np.array([3.0, 7.5]), np.array([1.1, 0.4])
    prob = qw.Problem(n=1024, topology="circular", schedule="lin")
    a = [qw.rk45_batch(prob, T, g, device=d) for d in (0, 1)]
    assert np.array_equal(a[0][0], a[1][0])
    prob = qw.Problem(n=1024, topology="circular", schedule="static")
    b = [qw.ti_exact_batch(prob, T, g, device=d) for d in (0, 1)]
    assert np.array_equal(b[0][0], b[1][0])
