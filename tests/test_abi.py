"""CPU-only checks of the drop-in boundary: the library loads, exports every symbol that
include/ssi.h declares, and fails loudly (no CPU fallback) when there is no GPU."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]


def _header_symbols():
    text = (ROOT / "include" / "ssi.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ssi_[a-z_0-9]+)\s*\(", text)))


def test_header_declares_expected_entry_points():
    syms = _header_symbols()
    for must in ("ssi_ctx_create", "ssi_set_model", "ssi_set_data", "ssi_set_subspace", "ssi_logpost_batch", "ssi_mh_run",
                 "ssi_rng_replay", "ssi_swa_begin", "ssi_swa_push", "ssi_swa_finish", "ssi_stats"):
        assert must in syms


def test_library_exports_every_declared_symbol(ssi):
    lib = ssi.load()
    from subspaceinference_jl_b200 import _lib
    declared = _header_symbols()
    assert sorted(_lib.SIGNATURES) == declared, "ctypes table and include/ssi.h disagree"
    for name in declared:
        assert hasattr(lib, name), f"{name} not exported by libssi.so"
    assert lib.ssi_version() == 200


def test_rng_replay_matches_oracle_on_host(ssi):
    """ssi_rng_replay is host code (no GPU needed): it must reproduce the oracle's stream."""
    import ssi_oracle as orc
    lib = ssi.load()
    for seed, chain, step, M in [(0, 0, 0, 3), (1234, 7, 9, 5), (2 ** 40 + 17, 65535, 999, 20), (5, 4095, 1, 1)]:
        eps = np.empty(M, np.float32)
        e = C.c_double()
        assert lib.ssi_rng_replay(seed, chain, step, M, C.c_void_p(eps.ctypes.data), C.byref(e)) == 0
        ref = orc.rng_normals(seed, chain, step, M)
        # double-precision Box-Muller rounded once to float: equal up to 1 ulp (libm vs numpy)
        np.testing.assert_allclose(eps, ref, rtol=2.5e-7, atol=1e-9)
        assert abs(e.value - orc.rng_exponential(seed, chain, step)) <= 1e-13 * max(1.0, e.value)
    assert lib.ssi_rng_replay(0, -1, 0, 3, None, None) == -1


def test_no_cpu_fallback(ssi):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(ssi.SsiError) as ei:
        ssi.Engine(0)
    assert ei.value.code == -2 and "no CPU fallback" in str(ei.value)


def test_host_mirror_shapes(ssi):
    """Flux stand-ins: parameter order and DataLoader semantics (host logic only)."""
    import ssi_oracle as orc
    rng = np.random.default_rng(0)
    m = ssi.Chain(ssi.Dense(10, 20, rng=rng), ssi.Dense(20, 20, ssi.relu, rng=rng), ssi.Dense(20, 2, rng=rng))
    assert m.dims == (10, 20, 20, 2) and m.acts == (0, 1, 0)
    w = ssi.extract_params(m)
    assert w.shape == (682,) and w.dtype == np.float32
    X = rng.random((10, 7)).astype(np.float32)
    np.testing.assert_allclose(m(X).detach().numpy(), orc.forward(w, m.dims, m.acts, X), rtol=2e-5, atol=1e-6)
    w2 = rng.standard_normal(682).astype(np.float32)
    ssi.load_params(m, w2)
    np.testing.assert_array_equal(ssi.extract_params(m), w2)
    dl = ssi.DataLoader(X, rng.random((2, 7)), batchsize=3)
    assert len(dl) == 3 and [b[0].shape[1] for b in dl] == [3, 3, 1]
    with pytest.raises(ValueError):
        ssi.sub_inference(m, dl, w, np.zeros((682, 3), np.float32), alg="bogus")
    with pytest.raises(ValueError):
        ssi.subspace_inference(m, None, dl, None, method="nope")


def test_training_step_host_restatement(ssi):
    """The host restatement of `gradient + Flux.update!` that the device trainer is checked against (tests/test_gpu_train.py):
    the loader's index batches are what its iteration slices with, a Descent step is W -= eta * dL/dW with L = mse, and ADAM's
    first step moves every weight by eta * sign(g) (Flux 0.11 rule with bias correction)."""
    import ssi_oracle as orc
    rng = np.random.default_rng(3)
    X = rng.standard_normal((4, 11)).astype(np.float32)
    Y = rng.standard_normal((2, 11)).astype(np.float32)
    dl = ssi.DataLoader(X, Y, batchsize=4, shuffle=True, rng=np.random.default_rng(9))
    dl2 = ssi.DataLoader(X, Y, batchsize=4, shuffle=True, rng=np.random.default_rng(9))
    for sel, (xb, yb) in zip(dl.index_batches(), dl2):
        np.testing.assert_array_equal(xb, X[:, sel])
        np.testing.assert_array_equal(yb, Y[:, sel])
    # single linear layer: dL/dW = 2/(O nb) (W x + b - y) x'
    m = ssi.Chain(ssi.Dense(4, 2, rng=rng))
    w0 = ssi.extract_params(m).astype(np.float64)
    W0, b0 = w0[:8].reshape(4, 2).T, w0[8:]
    cost = lambda mm, x, y: ssi.mse(mm(x), y)
    loss = ssi.train_step(m, cost, ssi.Descent(0.1), X, Y)
    resid = W0 @ X + b0[:, None] - Y
    np.testing.assert_allclose(loss, np.mean(resid ** 2), rtol=1e-5)
    gW, gb = 2.0 / resid.size * resid @ X.T, 2.0 / resid.size * resid.sum(axis=1)
    w1 = ssi.extract_params(m)
    np.testing.assert_allclose(w1[:8].reshape(4, 2).T, W0 - 0.1 * gW, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(w1[8:], b0 - 0.1 * gb, rtol=1e-5, atol=1e-6)
    m2 = ssi.Chain(ssi.Dense(4, 2, rng=rng))
    before = ssi.extract_params(m2).copy()
    ssi.train_step(m2, cost, ssi.ADAM(0.01), X, Y)
    step = ssi.extract_params(m2) - before
    np.testing.assert_allclose(np.abs(step), 0.01, rtol=1e-3)
    assert orc.n_params(m2.dims) == step.shape[0]
