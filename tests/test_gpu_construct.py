"""Subspace construction streams vs the oracle: W_swa and P within 1e-4 (P up to column sign)."""
from pathlib import Path

import numpy as np
import pytest

import ssi_oracle as orc

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).parent / "golden"


def _run(engine, snaps, ns, M, install=False):
    n = snaps[0].shape[0]
    engine.swa_begin(n, len(snaps))
    for w, s in zip(snaps, ns):
        engine.swa_push(w, s)
    assert engine.swa_columns() == len(snaps)
    return engine.swa_finish(M, install=install)


def _rel(a, b):
    return np.abs(a - b).max() / np.abs(b).max()


def test_golden_construction(ssi, engine):
    g = np.load(GOLD / "construct_small.npz")
    W_swa, P, s = _run(engine, list(g["snapshots"]), g["n_scalars"], int(g["M"]))
    assert _rel(W_swa, g["W_swa"]) < 1e-4
    assert _rel(orc.align_signs(P.astype(np.float64), g["P"]), g["P"]) < 1e-4
    np.testing.assert_allclose(s[:3], g["s"][:3], rtol=1e-5)


@pytest.mark.parametrize("n,K,M", [(3001, 64, 30), (2050, 70, 40), (1203, 90, 64)])
def test_wide_subspaces(ssi, engine, n, K, M):
    """M > 20: k_form_p_const runs with 2 rows (M <= 32) or 1 row (beyond) per thread.  M strong directions with a 5 % ladder
    of scales, so that every wanted singular vector is well separated."""
    rng = np.random.default_rng(n + K + M)
    w = rng.standard_normal(n) * 0.1
    dirs = rng.standard_normal((M, n)) * (0.95 ** np.arange(M))[:, None]
    snaps, ns = [], []
    for k in range(K):
        w = w + dirs.T @ rng.standard_normal(M) * 0.05 + 1e-5 * rng.standard_normal(n)
        snaps.append(w.astype(np.float32))
        ns.append(float(k + 1))
    W_swa, P, s = _run(engine, snaps, ns, M)
    W_ref, P_ref, s_ref, A = orc.construct_from_snapshots(snaps, ns, M)
    assert _rel(W_swa, W_ref) < 1e-4
    np.testing.assert_allclose(s[:M], s_ref[:M], rtol=1e-4)
    # rank-M projector (insensitive to rotations inside clusters of close singular values), then column by column
    x = rng.standard_normal(n)
    Pd = P.astype(np.float64)
    assert _rel(Pd @ (Pd.T @ x), P_ref @ (P_ref.T @ x)) < 2e-4
    assert _rel(orc.align_signs(Pd, P_ref), P_ref) < 2e-3


@pytest.mark.parametrize("n,K,M", [(682, 15, 3), (1001, 40, 5), (4096, 100, 20), (37, 64, 4), (50000, 33, 20),
                                   (300, 1, 1), (300, 2, 2), (300, 3, 2), (2000, 129, 10), (2000, 160, 16)])   # eigen-solver edges: one pair, a bye, five rows per lane
def test_random_snapshots_vs_oracle(ssi, engine, n, K, M):
    rng = np.random.default_rng(n * 1000 + K)
    w = rng.standard_normal(n) * 0.1
    # a few dominant drift directions + noise, so that the top-M singular values are separated
    dirs = rng.standard_normal((M, n)) * (2.0 ** -np.arange(M))[:, None]
    snaps, ns = [], []
    for k in range(K):
        w = w + dirs.T @ rng.standard_normal(M) * 0.05 + 1e-3 * rng.standard_normal(n)
        snaps.append(w.astype(np.float32))
        ns.append(float(k // 4 + 1))          # several mini-batches share one epoch index (Q2)
    W_swa, P, s = _run(engine, snaps, ns, M)
    W_ref, P_ref, s_ref, A = orc.construct_from_snapshots(snaps, ns, M)
    assert _rel(W_swa, W_ref) < 1e-4
    np.testing.assert_allclose(s[:M], s_ref[:M], rtol=1e-4)
    assert _rel(orc.align_signs(P.astype(np.float64), P_ref), P_ref) < 1e-4
    # P columns orthogonal with norms s_k
    G = P.astype(np.float64).T @ P.astype(np.float64)
    np.testing.assert_allclose(G, np.diag(s_ref[:M] ** 2), atol=2e-4 * s_ref[0] ** 2)


def test_more_columns_than_parameters(ssi, engine):
    """README example: batchsize-1 loader, T=10 -> K = 1000 columns > n = 682 (Q3); here a
    reduced K=300 > n=120 keeps the eigen-solve short."""
    rng = np.random.default_rng(9)
    n, K, M = 120, 300, 3
    w = rng.standard_normal(n)
    dirs = rng.standard_normal((M, n)) * np.array([4.0, 2.0, 1.0])[:, None]
    snaps, ns = [], []
    for k in range(K):
        w = w + dirs.T @ rng.standard_normal(M) * 0.1 + 1e-3 * rng.standard_normal(n)
        snaps.append(w.astype(np.float32))
        ns.append(float(k // 100 + 1))
    W_swa, P, s = _run(engine, snaps, ns, M)
    W_ref, P_ref, s_ref, _ = orc.construct_from_snapshots(snaps, ns, M)
    assert _rel(W_swa, W_ref) < 1e-4
    assert _rel(orc.align_signs(P.astype(np.float64), P_ref), P_ref) < 1e-4
    assert np.all(s[n:] < 1e-3 * s[0])       # rank <= n


def test_rank_error_and_install(ssi, engine):
    rng = np.random.default_rng(4)
    dims, acts = (10, 20, 20, 2), (0, 0, 0)
    n = orc.n_params(dims)
    snaps = [orc.glorot_flat(rng, dims) for _ in range(6)]
    engine.swa_begin(n, 6)
    for k in range(2):
        engine.swa_push(snaps[k], 1.0)
    with pytest.raises(ssi.SsiError) as ei:
        engine.swa_finish(3)                 # U[:,1:M] would throw in the reference
    assert ei.value.code == -4
    for k in range(2, 6):
        engine.swa_push(snaps[k], 2.0)
    with pytest.raises(ssi.SsiError):
        engine.swa_push(snaps[0], 3.0)       # K_max reached
    engine.set_model(dims, acts)
    X = rng.random((10, 50)).astype(np.float32)
    Y = rng.random((2, 50)).astype(np.float32)
    engine.set_data(X, Y)
    W_swa, P, _ = engine.swa_finish(3, install=True)
    Z = rng.standard_normal((3, 4)).astype(np.float32)
    lp = engine.logpost(Z)
    ref, _ = orc.logpost_batch(orc.Problem(dims, acts, X, Y, W_swa, P), Z)
    np.testing.assert_allclose(lp, ref, rtol=1e-5)


@pytest.mark.parametrize("n,K,M,expect_tensor", [
    (70000, 100, 20, None),             # wanted directions reach into the noise floor: either Gram may be chosen
    (131072, 20, 20, False),            # K = M with a nearly degenerate tail: the conditioning check must reject the tensor Gram
    (65536 + 4 * 13 + 3, 37, 5, True),  # well separated spectrum: the tensor Gram must be kept
])
def test_tensor_core_gram_vs_oracle(ssi, engine, n, K, M, expect_tensor):
    """Large n, K <= 128: the Gram runs on the tensor cores (TF32 hi/lo split, FP32 chunks drained to FP64).
    P and W_swa must still match the Float64 oracle to 1e-4; the FP64 SIMT Gram is the cross-check."""
    rng = np.random.default_rng(n + K)
    w = (rng.standard_normal(n) * 0.1).astype(np.float32)
    dirs = rng.standard_normal((M, n)).astype(np.float32) * (1.6 ** -np.arange(M, dtype=np.float32))[:, None]
    snaps, ns = [], []
    for k in range(K):
        w = w + (dirs.T @ rng.standard_normal(M).astype(np.float32)) * 0.05 + 1e-3 * rng.standard_normal(n).astype(np.float32)
        snaps.append(w.copy())
        ns.append(float(k + 1))
    W_ref, P_ref, s_ref, _ = orc.construct_from_snapshots(snaps, ns, M)
    res = {}
    for fp64 in (-1, 0, 1):        # -1: tensor-core Gram unguarded (reported, not asserted), 0: guarded default, 1: FP64 SIMT Gram
        engine.set_option("gram_fp64", fp64)
        W_swa, P, s = _run(engine, snaps, ns, M)
        st = engine.stats()
        err_p = _rel(orc.align_signs(P.astype(np.float64), P_ref), P_ref)
        print(f"n={n} K={K} M={M} gram_fp64={fp64}: gram_path={st.gram_path} risk={st.gram_risk:.3g} sweeps={st.jacobi_sweeps} "
              f"err_P={err_p:.3g} err_s={np.abs(s[:M] - s_ref[:M]).max() / s_ref[0]:.3g}")
        res[fp64] = s
        if fp64 < 0:
            assert st.gram_path == 2
            continue
        assert st.gram_path == (1 if fp64 else st.gram_path)
        assert _rel(W_swa, W_ref) < 1e-4
        assert err_p < 1e-4, f"fp64={fp64}"
        # TF32 hi/lo products + FP32 chunks: the Gram is accurate to ~1e-7 of its largest entry, i.e. singular
        # values to ~1e-5 of s_1 in absolute terms; the FP64 Gram resolves every one of them relatively
        if st.gram_path != 2:
            np.testing.assert_allclose(s[:M], s_ref[:M], rtol=1e-5)
        else:
            np.testing.assert_allclose(s[:M], s_ref[:M], rtol=0, atol=1e-5 * s_ref[0])
        if fp64 == 0 and expect_tensor is not None:
            assert (st.gram_path == 2) == expect_tensor, st.gram_risk
    engine.set_option("gram_fp64", 0)


@pytest.mark.parametrize("n,K,M", [(1003, 12, 4), (150000, 24, 6)])
def test_row_sharded_construction_matches_single_context(ssi, n, K, M):
    """rows sharded over two contexts (what two ranks do), Grams summed (the all-reduce), replicated
    eigen-solve, P rows per shard.  The stacked result must equal the single-context construction and the oracle."""
    import torch
    rng = np.random.default_rng(n)
    w = (rng.standard_normal(n) * 0.1).astype(np.float32)
    dirs = rng.standard_normal((M + 2, n)).astype(np.float32) * (2.0 ** -np.arange(M + 2, dtype=np.float32))[:, None]
    snaps, ns = [], []
    for k in range(K):
        w = w + (dirs.T @ rng.standard_normal(M + 2).astype(np.float32)) * 0.05
        snaps.append(w.copy())
        ns.append(float(k // 3 + 1))
    W_ref, P_ref, s_ref, _ = orc.construct_from_snapshots(snaps, ns, M)
    from subspaceinference_jl_b200.api import shard_rows
    engs = [ssi.Engine(0), ssi.Engine(0)]
    try:
        rows = [shard_rows(n, r, 2) for r in range(2)]
        assert rows[0][0] == 0 and rows[0][1] == rows[1][0] and rows[1][1] == n
        for e, (b, t) in zip(engs, rows):
            e.swa_begin(t - b, K)
            for wk, sk in zip(snaps, ns):
                e.swa_push(wk[b:t], sk)
        out = None
        for exact in (False, True):
            Gs = [torch.empty(K * K, dtype=torch.float64, device="cuda:0") for _ in engs]
            for e, G in zip(engs, Gs):
                e.swa_gram_dev(G.data_ptr(), exact)
                e.sync()
            G = Gs[0] + Gs[1]                                  # the all-reduce
            torch.cuda.synchronize()
            outs = [e.swa_finish_gram(M, G.data_ptr(), gram_exact=exact) for e in engs]
            assert (outs[0] is None) == (outs[1] is None)      # every rank takes the same decision
            if outs[0] is not None:
                out = outs
                break
        W_swa = np.concatenate([o[0] for o in out])
        P = np.concatenate([o[1] for o in out], axis=0)
        np.testing.assert_array_equal(out[0][2], out[1][2])    # replicated spectrum
        assert _rel(W_swa, W_ref) < 1e-4
        assert _rel(orc.align_signs(P.astype(np.float64), P_ref), P_ref) < 1e-4
        np.testing.assert_allclose(out[0][2][:M], s_ref[:M], rtol=1e-4)
    finally:
        for e in engs:
            e.close()
