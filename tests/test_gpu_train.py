"""The step before the path: mini-batch training on the device vs the host restatement of
`gradient + Flux.update!` (src/subspace_construction.jl:39-43; subspaceinference_jl_b200/flux.py runs it with torch CPU
autograd in Float32)."""
import copy

import numpy as np
import pytest

import ssi_oracle as orc
from conftest import stable_seed

pytestmark = pytest.mark.gpu


def _model(ssi, dims, acts, rng):
    fn = {0: ssi.identity, 1: ssi.relu, 2: ssi.tanh, 3: ssi.sigmoid}
    return ssi.Chain(*[ssi.Dense(dims[l], dims[l + 1], fn[acts[l]], rng=rng) for l in range(len(acts))])


@pytest.mark.parametrize("dims,acts,N,bs,opt", [
    ((10, 20, 20, 2), (0, 0, 0), 100, 20, "adam"),           # README model and optimiser family
    ((13, 50, 1), (1, 0), 400, 100, "descent"),              # C2 shape
    ((7, 33, 65, 3), (2, 3, 1), 333, 50, "adam"),            # tanh / sigmoid hidden, relu output, ragged last batch
    ((5, 300, 4), (1, 2), 64, 64, "descent"),                # one full-batch step per epoch, width > one GEMM tile
])
def test_training_steps_match_host_autograd(ssi, engine, dims, acts, N, bs, opt):
    rng = np.random.default_rng(stable_seed(dims, N, bs, opt))
    X = rng.standard_normal((dims[0], N)).astype(np.float32)
    Y = rng.standard_normal((dims[-1], N)).astype(np.float32)
    m = _model(ssi, dims, acts, rng)
    cost = lambda mm, x, y: ssi.mse(mm(x), y)
    host_opt = ssi.ADAM(0.01) if opt == "adam" else ssi.Descent(0.05)
    engine.set_model(m.dims, m.acts)
    engine.set_data(X, Y)
    engine.train_begin(ssi.extract_params(m), opt, host_opt.eta, getattr(host_opt, "beta", (0.9, 0.999)))
    loader = ssi.DataLoader(X, Y, batchsize=bs, shuffle=True, rng=np.random.default_rng(3))
    step = 0
    for epoch in range(3):
        for sel in loader.index_batches():
            # alternate between index batches and contiguous windows
            if step % 2 == 1:
                j0 = int(sel[0]) % max(1, N - len(sel))
                sel = np.arange(j0, j0 + len(sel))
                dev_loss = engine.train_step((j0, len(sel)))
            else:
                dev_loss = engine.train_step(sel)
            host_loss = ssi.train_step(m, cost, host_opt, X[:, sel], Y[:, sel])
            np.testing.assert_allclose(dev_loss, host_loss, rtol=2e-5, atol=1e-7, err_msg=f"loss at step {step}")
            step += 1
    Wd, Wh = engine.train_weights(), ssi.extract_params(m)
    scale = np.abs(Wh).max()
    assert np.abs(Wd - Wh).max() <= 2e-4 * scale, np.abs(Wd - Wh).max() / scale
    engine.train_end()
    with pytest.raises(ssi.SsiError):
        engine.train_step((0, 1))                            # no training state any more


def test_device_training_feeds_the_construction(ssi, capsys):
    """subspace_construction(..., device_train=True) == the host loop: same snapshots, same W_swa, same P (up to sign)."""
    rng = np.random.default_rng(11)
    X = rng.standard_normal((13, 600)).astype(np.float32)
    W_true = rng.standard_normal((1, 13)).astype(np.float32)
    Y = (np.tanh(W_true @ X) + 0.1 * rng.standard_normal((1, 600))).astype(np.float32)
    m_host = _model(ssi, (13, 50, 1), (1, 0), rng)
    m_dev = copy.deepcopy(m_host)
    cost = lambda mm, x, y: ssi.mse(mm(x), y)
    out = {}
    for name, mm, dev in (("host", m_host, False), ("device", m_dev, True)):
        data = ssi.DataLoader(X, Y, batchsize=100)
        out[name] = ssi.subspace_construction(mm, cost, data, ssi.ADAM(0.01), T=6, c=2, M=3, print_freq=3, device_train=dev)
    assert capsys.readouterr().out.count("Traing loss") == 4
    (Wh, Ph), (Wd, Pd) = out["host"], out["device"]
    np.testing.assert_allclose(Wd, Wh, rtol=1e-4, atol=1e-4 * np.abs(Wh).max())
    Pd = orc.align_signs(Pd, Ph)
    assert np.abs(Pd - Ph).max() <= 1e-3 * np.abs(Ph).max(), np.abs(Pd - Ph).max() / np.abs(Ph).max()
    np.testing.assert_allclose(ssi.extract_params(m_dev), ssi.extract_params(m_host), rtol=0, atol=2e-4 * np.abs(Wh).max())


def test_device_training_rejects_other_costs(ssi):
    rng = np.random.default_rng(2)
    X = rng.standard_normal((4, 50)).astype(np.float32)
    Y = rng.standard_normal((2, 50)).astype(np.float32)
    m = _model(ssi, (4, 8, 2), (1, 0), rng)
    import torch
    l1 = lambda mm, x, y: torch.mean(torch.abs(mm(x) - y))
    with pytest.raises(ValueError, match="not mse"):
        ssi.subspace_construction(m, l1, ssi.DataLoader(X, Y, batchsize=10), ssi.Descent(0.1), T=2, c=1, M=2, device_train=True)
    with pytest.raises(TypeError):
        ssi.subspace_construction(m, l1, ssi.DataLoader(X, Y, batchsize=10), object(), T=2, c=1, M=2, device_train=True)


def test_whole_pipeline_on_the_device(ssi, capsys):
    """subspace_inference(..., device_train=True): training, snapshots, construction and sampling without the weights
    leaving the device between the steps; every returned lp is the oracle density of the returned weight vector."""
    rng = np.random.default_rng(21)
    X = rng.random((10, 100)).astype(np.float32)
    Y = rng.random((2, 100)).astype(np.float32)
    data = ssi.DataLoader(X, Y, batchsize=20, shuffle=True, rng=np.random.default_rng(1))
    m = _model(ssi, (10, 20, 20, 2), (0, 0, 0), rng)
    chn, lp, W_swa = ssi.subspace_inference(m, lambda mm, x, y: ssi.mse(mm(x), y), data, ssi.ADAM(0.01), itr=8, T=6, c=1, M=3,
                                            seed=2, device_train=True)
    assert len(chn) == 8 and W_swa.shape == (682,) and "Traing loss" in capsys.readouterr().out
    for w, l in zip(chn, lp):
        np.testing.assert_allclose(l, orc.gaussian_loglik(orc.forward(w, m.dims, m.acts, X), Y, 1.0), rtol=1e-5)
