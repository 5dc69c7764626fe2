"""Non-linear subspace operator: density(z) of auto_inference, new_W = W_swa + decoder(z)
(src/space_inference.jl:246-251), against the Float64 oracle on every execution path."""
import numpy as np
import pytest

import ssi_oracle as orc

pytestmark = pytest.mark.gpu


def _decoder(rng, dims, scale=0.3):
    """Flux.destructure-ordered parameters of a Dense chain dims[0] -> ... -> dims[-1]."""
    parts = []
    for l in range(len(dims) - 1):
        parts += [(scale * rng.standard_normal((dims[l + 1], dims[l])) / np.sqrt(dims[l])).reshape(-1, order="F"),
                  0.05 * rng.standard_normal(dims[l + 1])]
    return np.concatenate(parts).astype(np.float32)


def _setup(engine, prob, ddims, dacts, theta):
    engine.set_model(prob.dims, prob.acts)
    engine.set_data(prob.X, prob.Y)
    engine.set_decoder(prob.W_swa, ddims, dacts, theta)


CASES = [
    # model dims, acts, N, decoder hidden dims, decoder acts, paths
    ((10, 20, 20, 2), (0, 0, 0), 100, (3, 8), (2, 0), ("fused", "layered", "tensor")),        # README net, z in R^3 -> 8 -> n
    ((13, 50, 1), (1, 0), 700, (5, 16, 12), (1, 2, 0), ("basis", "layered")),                  # two hidden decoder layers
    ((96, 128, 128, 10), (1, 1, 0), 300, (4, 24), (2, 0), ("tensor", "layered")),               # tensor cores, z' in R^24
    ((13, 50, 1), (1, 0), 400, (6,), (0,), ("basis", "fused")),                                   # a single Dense(M, n): z' = z
]


@pytest.mark.parametrize("dims,acts,N,dh,dacts,paths", CASES)
def test_decoder_density_vs_oracle(ssi, engine, dims, acts, N, dh, dacts, paths):
    rng = np.random.default_rng(len(dims) * 1000 + N)
    n = orc.n_params(dims)
    X = rng.standard_normal((dims[0], N)).astype(np.float32)
    Y = rng.standard_normal((dims[-1], N)).astype(np.float32)
    prob = orc.Problem(dims, acts, X, Y, orc.glorot_flat(rng, dims), np.zeros((n, 1), np.float32))
    ddims = tuple(dh) + (n,)
    theta = _decoder(rng, ddims, 0.2)
    _setup(engine, prob, ddims, dacts, theta)
    B = 9
    Z = rng.standard_normal((ddims[0], B)).astype(np.float32)
    for mask in (1, 3, 7):
        ref = np.array([orc.density_decoder(prob, theta, ddims, dacts, Z[:, b], 0.7, 1.3, 0.9, mask) for b in range(B)])
        for path in paths:
            engine.set_option("path", {"fused": 1, "layered": 2, "tensor": 3, "basis": 4}[path])
            lp, terms = engine.logpost(Z, 0.7, 1.3, 0.9, mask=mask, return_terms=True)
            assert ssi.PATH_NAMES[engine.stats().last_path] == path
            np.testing.assert_allclose(lp, ref, rtol=1e-5, err_msg=f"{path} mask={mask}")
            np.testing.assert_allclose(terms[0], [orc.density_decoder(prob, theta, ddims, dacts, Z[:, b], 0.7) for b in range(B)], rtol=1e-5)
    engine.set_option("path", 0)
    # map(z -> W_swa + decoder(z), chm)
    W = engine.project(Z[:, :3])
    for b in range(3):
        w_ref = prob.W_swa.astype(np.float64) + orc.decoder_forward(theta, ddims, dacts, Z[:, b])
        np.testing.assert_allclose(W[:, b], w_ref, rtol=0, atol=1e-5 * np.abs(w_ref).max())
    # gradient w.r.t. the user's z: chain rule through the decoder's hidden layers, against central differences of the oracle
    lp, g = engine.logpost_grad(Z[:, :2], 0.7, 1.3, 0.9, mask=7)
    for b in range(2):
        f = lambda zz: orc.density_decoder(prob, theta, ddims, dacts, zz, 0.7, 1.3, 0.9, 7)
        z0 = Z[:, b].astype(np.float64)
        np.testing.assert_allclose(lp[b], f(z0), rtol=1e-5)
        h = 1e-4
        fd = np.array([(f(z0 + h * np.eye(len(z0))[i]) - f(z0 - h * np.eye(len(z0))[i])) / (2 * h) for i in range(len(z0))])
        np.testing.assert_allclose(g[:, b], fd, rtol=0, atol=2e-3 * np.linalg.norm(fd))


def test_decoder_nonlinear_head_and_sampler(ssi, engine):
    """A tanh head cannot be folded into an affine subspace: the weights are materialised per sample (LAYERED path).  The
    sampler works in the user's z either way: teacher-forced RWMH decisions against the oracle's decoder density."""
    rng = np.random.default_rng(77)
    dims, acts, N = (10, 20, 20, 2), (0, 1, 0), 100
    n = orc.n_params(dims)
    prob = orc.Problem(dims, acts, rng.random((10, N), dtype=np.float32), rng.random((2, N), dtype=np.float32),
                       orc.glorot_flat(rng, dims), np.zeros((n, 1), np.float32))
    for dacts in ((2, 2), (2, 0)):
        ddims = (3, 8, n)
        theta = _decoder(rng, ddims, 0.3)
        _setup(engine, prob, ddims, dacts, theta)
        Z = rng.standard_normal((3, 6)).astype(np.float32)
        ref = np.array([orc.density_decoder(prob, theta, ddims, dacts, Z[:, b], 0.5) for b in range(6)])
        np.testing.assert_allclose(engine.logpost(Z, 0.5), ref, rtol=1e-5)
        assert ssi.PATH_NAMES[engine.stats().last_path] == ("layered" if dacts[-1] else "fused")
        if dacts[-1]:
            with pytest.raises(ssi.SsiError):
                engine.logpost(Z, 0.5, mask=3)               # the M-space weight prior needs an affine head
        C, S, seed, sz = 5, 8, 31, 0.2
        zt, lt, at = engine.mh_run(C, S, seed, sigma_z=sz, sigma_m=0.5, chain_offset=2)
        assert zt.shape == (3, C, S)
        for ci in range(C):
            z = zt[:, ci, :].T
            np.testing.assert_array_equal(z[0], orc.propose_f32(np.zeros(3, np.float32), sz, orc.rng_normals(seed, 2 + ci, 0, 3)))
            for t in range(1, S):
                zp = orc.propose_f32(z[t - 1], sz, orc.rng_normals(seed, 2 + ci, t, 3))
                margin = (orc.density_decoder(prob, theta, ddims, dacts, zp, 0.5) - orc.density_decoder(prob, theta, ddims, dacts, z[t - 1], 0.5)
                          + orc.rng_exponential(seed, 2 + ci, t))
                if abs(margin) > 1e-3:
                    assert bool(at[ci, t]) == (margin > 0)
    # a linear subspace replaces the decoder again
    P = (0.1 * rng.standard_normal((n, 3))).astype(np.float32)
    engine.set_subspace(prob.W_swa, P)
    prob2 = orc.Problem(dims, acts, prob.X, prob.Y, prob.W_swa, P)
    np.testing.assert_allclose(engine.logpost(Z, 0.5), orc.logpost_batch(prob2, Z, 0.5)[0], rtol=1e-5)


def test_autoencoder_inference_api(ssi, capsys):
    """autoencoder_inference / auto_encoder_subspace / auto_inference with the reference's signatures
    (src/space_inference.jl:198-210, 238-318; src/subspace_construction.jl:93-143)."""
    rng = np.random.default_rng(5)
    X = rng.random((10, 60)).astype(np.float32)
    Y = rng.random((2, 60)).astype(np.float32)
    data = ssi.DataLoader(X, Y, batchsize=20, shuffle=True, rng=rng)
    m = ssi.Chain(ssi.Dense(10, 20, rng=rng), ssi.Dense(20, 2, rng=rng))
    n = 10 * 20 + 20 + 20 * 2 + 2
    enc = ssi.Chain(ssi.Dense(n, 16, ssi.tanh, rng=rng), ssi.Dense(16, 3, rng=rng))
    dec = ssi.Chain(ssi.Dense(3, 16, ssi.tanh, rng=rng), ssi.Dense(16, n, rng=rng))
    cost = lambda mm, x, y: ssi.mse(mm(x), y)
    before = ssi.extract_params(dec).copy()
    chn, lp = ssi.autoencoder_inference(m, cost, data, ssi.ADAM(0.01), enc, dec, itr=12, T=4, c=1, M=3, alg=":rwmh", σ_z=0.3, seed=9)
    assert "Traing loss" in capsys.readouterr().out
    assert len(chn) == 12 and chn[0].shape == (n,) and lp.shape == (12,)
    assert np.abs(ssi.extract_params(dec) - before).max() > 0          # the auto-encoder was trained
    for w, l in zip(chn, lp):                                            # every lp is the oracle likelihood of the returned weights
        np.testing.assert_allclose(l, orc.gaussian_loglik(orc.forward(w, m.dims, m.acts, X), Y, 1.0), rtol=1e-5)
    with pytest.raises(NotImplementedError):
        ssi.auto_inference(m, data, dec, chn[0], M=3)                    # the reference's default :hmc is a host-side sampler
