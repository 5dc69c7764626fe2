"""Tensor-core (tcgen05/TMEM/TMA, split precision: calibrated FP16 planes, BF16 planes as fallback) path vs the Float64 oracle."""
from pathlib import Path

import numpy as np
import pytest

import ssi_oracle as orc
from conftest import stable_seed

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).parent / "golden"
RTOL = 1e-5


def _setup(engine, prob):
    engine.set_model(prob.dims, prob.acts)
    engine.set_data(prob.X, prob.Y)
    engine.set_subspace(prob.W_swa, prob.P)


def _rand_problem(dims, acts, N, M, seed):
    rng = np.random.default_rng(seed)
    n = orc.n_params(dims)
    X = rng.random((dims[0], N), dtype=np.float32)
    Y = rng.standard_normal((dims[-1], N)).astype(np.float32)
    return orc.Problem(dims, acts, X, Y, orc.glorot_flat(rng, dims), (0.05 * rng.standard_normal((n, M))).astype(np.float32)), rng


def test_golden_wide_small_tensor_path(ssi, engine):
    g = np.load(GOLD / "logpost_wide_small.npz")
    prob = orc.Problem(tuple(int(d) for d in g["dims"]), tuple(int(a) for a in g["acts"]), g["X"], g["Y"], g["W_swa"], g["P"])
    _setup(engine, prob)
    engine.set_option("path", ssi.PATH_TENSOR)
    lp, terms = engine.logpost(g["Z"], 1.0, 1.0, 1.0, mask=7, return_terms=True)
    assert engine.stats().last_path == ssi.PATH_TENSOR
    np.testing.assert_allclose(terms, g["terms"], rtol=RTOL)


@pytest.mark.parametrize("dims,acts,N,M,B", [
    ((64, 64, 3), (1, 0), 128, 2, 1),                       # one tile, one k-block, one sample
    ((96, 128, 128, 10), (1, 1, 0), 300, 20, 8),            # ragged N, K padding 96 -> 128
    ((784, 256, 512, 10), (1, 1, 0), 1000, 20, 19),         # BN=256, more than one group (G=16)
    ((50, 192, 64, 320, 12), (2, 1, 3, 0), 513, 5, 3),      # 4 tensor layers (ping-pong), BN=64, tanh/sigmoid
    ((32, 128, 300), (1, 0), 200, 3, 2),                    # wide output layer: O=300 > 256 is rejected, see below
    ((20, 1024, 1), (1, 0), 2500, 3, 5),                    # single hidden layer, scalar output
])
def test_random_shapes_tensor_vs_oracle(ssi, engine, dims, acts, N, M, B):
    prob, rng = _rand_problem(dims, acts, N, M, stable_seed(dims, N))
    Z = rng.standard_normal((M, B)).astype(np.float32)
    _setup(engine, prob)
    engine.set_option("path", ssi.PATH_TENSOR)
    if dims[-1] > 256:
        with pytest.raises(ssi.SsiError) as ei:
            engine.logpost(Z, 0.8)
        assert ei.value.code == -5
        return
    lp = engine.logpost(Z, 0.8)
    ref, _ = orc.logpost_batch(prob, Z, 0.8)
    np.testing.assert_allclose(lp, ref, rtol=RTOL)
    # A-B variant: round 1's BF16x3 planes (both operands BF16 | BF16) instead of the calibrated FP16 planes
    engine.set_option("tc_precision", 0)
    lp_bf16 = engine.logpost(Z, 0.8)
    np.testing.assert_allclose(lp_bf16, ref, rtol=RTOL)
    engine.set_option("tc_precision", 1)
    # the FP16 planes are the default because their products are FP32-grade: held to a third of the bar on these
    # adversarial problems (weights perturbed by several times their size, |lp| up to 1e7); what is left is the tensor
    # core's truncating FP32 accumulation, a systematic shrink of 1e-6..2.4e-6 here (1e-7..4e-7 on the BASELINE configs)
    # that is smooth in z and cancels in MH margins (tests/test_gpu_scale.py)
    np.testing.assert_allclose(lp, ref, rtol=RTOL / 3)
    assert engine.stats().tc_range_fallbacks == 0
    # A-B variant: one CTA per 128-row tile instead of CTA pairs (cta_group::2, 256-row MMAs over the two SMs of a TPC; odd
    # tile counts exercise the pair's dummy row block): the same products in the same order, so the same bits
    engine.set_option("tc_pair", 0)
    np.testing.assert_array_equal(engine.logpost(Z, 0.8), lp)
    engine.set_option("tc_pair", 1)
    # the FP32 SIMT path of the same library must agree with the oracle at least as well
    engine.set_option("path", ssi.PATH_LAYERED)
    lp_simt = engine.logpost(Z, 0.8)
    np.testing.assert_allclose(lp_simt, ref, rtol=RTOL)
    print("max rel err vs oracle: tensor %.2e (BF16x3 planes %.2e), simt %.2e" % (np.abs(lp / ref - 1).max(), np.abs(lp_bf16 / ref - 1).max(),
                                                                               np.abs(lp_simt / ref - 1).max()))


def test_tensor_path_batch_invariance_and_mh(ssi, engine):
    prob, rng = _rand_problem((128, 256, 256, 10), (1, 1, 0), 700, 20, 5)
    _setup(engine, prob)
    engine.set_option("path", ssi.PATH_TENSOR)
    Z = (0.1 * rng.standard_normal((20, 40))).astype(np.float32)
    full = engine.logpost(Z)
    part = np.concatenate([engine.logpost(Z[:, :3]), engine.logpost(Z[:, 3:25]), engine.logpost(Z[:, 25:])])
    np.testing.assert_array_equal(full, part)
    zt, lt, at = engine.mh_run(24, 5, 77, sigma_z=0.05)
    for c in (0, 23):
        for t in (0, 4):
            np.testing.assert_allclose(lt[c, t], orc.density(prob, zt[:, c, t]), rtol=RTOL)


def test_auto_picks_tensor_for_wide(ssi, engine):
    prob, rng = _rand_problem((128, 256, 256, 10), (1, 1, 0), 256, 20, 6)
    _setup(engine, prob)
    engine.set_option("path", ssi.PATH_AUTO)
    engine.logpost(np.zeros((20, 2), np.float32))
    assert engine.stats().last_path == ssi.PATH_TENSOR


def test_narrow_chain_runs_padded_on_tensor_path(ssi, engine):
    """Widths that are not multiples of 64 are zero-padded (13-50-1, sigmoid hidden layer: the
    padded hidden units output 0.5 and must be cancelled by zero weight columns)."""
    for dims, acts in (((13, 50, 1), (1, 0)), ((7, 33, 70, 3), (3, 2, 0))):
        prob, rng = _rand_problem(dims, acts, 333, 4, 11)
        _setup(engine, prob)
        Z = rng.standard_normal((4, 6)).astype(np.float32)
        ref, _ = orc.logpost_batch(prob, Z, 0.5)
        engine.set_option("path", ssi.PATH_TENSOR)
        np.testing.assert_allclose(engine.logpost(Z, 0.5), ref, rtol=RTOL)
        engine.set_option("path", ssi.PATH_AUTO)
        engine.logpost(Z, 0.5)
        # AUTO keeps small nets off the tensor path
        assert engine.stats().last_path == (ssi.PATH_BASIS if len(dims) == 3 else ssi.PATH_FUSED)


@pytest.mark.parametrize("dims,acts,N,M,B", [
    ((96, 128, 128, 10), (1, 1, 0), 300, 20, 8),
    ((784, 256, 512, 10), (1, 1, 0), 1000, 20, 19),
    ((20, 1024, 1), (2, 0), 700, 3, 5),
])
def test_first_layer_gemm_vs_basis_and_unfused_output(ssi, engine, dims, acts, N, M, B):
    """The first layer either runs as a GEMM (dataset tiles shared across the group's samples) or as the
    affine-in-z combination of precomputed bases; the output layer is either fused into the last hidden
    layer's epilogue or its own GEMM.  All four combinations must agree with the oracle."""
    prob, rng = _rand_problem(dims, acts, N, M, 1234)
    prob = orc.Problem(prob.dims, prob.acts, prob.X, prob.Y, prob.W_swa, (0.2 * prob.P).astype(np.float32))
    Z = rng.standard_normal((M, B)).astype(np.float32)
    _setup(engine, prob)
    engine.set_option("path", ssi.PATH_TENSOR)
    ref, _ = orc.logpost_batch(prob, Z, 0.8)
    for nobasis in (0, 1):
        for nofuse in (0, 1):
            engine.set_option("tc_nobasis", nobasis)
            engine.set_option("tc_nofuse", nofuse)
            lp = engine.logpost(Z, 0.8)
            np.testing.assert_allclose(lp, ref, rtol=RTOL, err_msg=f"nobasis={nobasis} nofuse={nofuse}")
    engine.set_option("tc_noorder", 1)
    np.testing.assert_allclose(engine.logpost(Z, 0.8), ref, rtol=RTOL)


@pytest.mark.parametrize("dims,acts,N,M,B,group", [
    ((64, 192, 128, 3), (1, 2, 0), 517, 20, 150, 0),     # two groups of 128 + a ragged one; tanh second layer
    ((40, 128, 10), (3, 0), 260, 31, 40, 0),             # M + 1 = 32: the largest subspace the tensor-core first layer takes
    ((40, 128, 10), (1, 0), 260, 40, 37, 0),             # M + 1 > 32: FP32 SIMT basis layer, groups of 32
    ((96, 128, 128, 10), (1, 1, 0), 300, 20, 70, 48),    # group size that is not a multiple of 32
    ((32, 128, 64, 20), (1, 1, 0), 200, 6, 10, 0),       # O > 12: the output layer is its own GEMM (FINAL epilogue)
    ((24, 64, 128, 64, 192, 5), (1, 2, 3, 1, 0), 130, 9, 6, 0),   # deep chain: HIDDEN epilogues ping-pong between the two buffers
])
def test_first_layer_on_tensor_cores_vs_simt_basis(ssi, engine, dims, acts, N, M, B, group):
    """The first layer as a K = M+1 GEMM over the samples (k_tc_basis_mma) and as the FP32 SIMT combination
    (k_tc_basis_layer) must both match the oracle; groups of up to 128 samples, ragged last group."""
    prob, rng = _rand_problem(dims, acts, N, M, 99)
    prob = orc.Problem(prob.dims, prob.acts, prob.X, prob.Y, prob.W_swa, (0.2 * prob.P).astype(np.float32))
    Z = rng.standard_normal((M, B)).astype(np.float32)
    _setup(engine, prob)
    engine.set_option("path", ssi.PATH_TENSOR)
    if group:
        engine.set_option("group", group)
    ref, _ = orc.logpost_batch(prob, Z, 0.8)
    out = {}
    for simt in (0, 1):
        engine.set_option("tc_simt_basis", simt)
        out[simt] = engine.logpost(Z, 0.8)
        np.testing.assert_allclose(out[simt], ref, rtol=RTOL, err_msg=f"tc_simt_basis={simt}")
        assert engine.stats().last_path == ssi.PATH_TENSOR
    # a sample's value does not depend on the group it lands in
    engine.set_option("tc_simt_basis", 0)
    np.testing.assert_array_equal(engine.logpost(Z[:, 5:9], 0.8), out[0][5:9])


def test_wide_full_size_properties(ssi, engine):
    """Full C3 size (784-1024-1024-10, N = 60000, M = 20): oracle on a few samples, additivity of the likelihood over a
    split of the dataset, datapoint-permutation invariance, and bitwise batch invariance across a group boundary."""
    prob = orc.make_problem("wide")
    assert prob.N == 60000 and prob.dims == (784, 1024, 1024, 10)
    _setup(engine, prob)
    rng = np.random.default_rng(7)
    B = 131                                                    # 128 + 3: crosses a group of samples
    Z = (0.1 * rng.standard_normal((prob.M, B))).astype(np.float32)
    lp = engine.logpost(Z, 1.0)
    assert engine.stats().last_path == ssi.PATH_TENSOR and engine.stats().last_units == B * prob.N
    idx = np.array([0, 127, 130])
    ref, _ = orc.logpost_batch(prob, Z[:, idx], 1.0)
    np.testing.assert_allclose(lp[idx], ref, rtol=RTOL)
    print("full-size wide: max rel err vs oracle %.2e" % np.abs(lp[idx] / ref - 1).max())
    # a sample's value does not depend on its batch
    np.testing.assert_array_equal(engine.logpost(Z[:, 126:131], 1.0), lp[126:131])
    # Gaussian likelihood is additive over datapoints: lp(all) = lp(first 25000) + lp(rest) (the -k/2 log(2 pi sigma^2) terms add too)
    Zs = Z[:, :4]
    cut = 25000
    engine.set_data(prob.X[:, :cut], prob.Y[:, :cut])
    lp_a = engine.logpost(Zs, 1.0)
    engine.set_data(prob.X[:, cut:], prob.Y[:, cut:])
    lp_b = engine.logpost(Zs, 1.0)
    np.testing.assert_allclose(lp_a + lp_b, lp[:4], rtol=2e-6)
    # permuting the datapoints changes only the summation order
    perm = rng.permutation(prob.N)
    engine.set_data(prob.X[:, perm], prob.Y[:, perm])
    np.testing.assert_allclose(engine.logpost(Zs, 1.0), lp[:4], rtol=2e-6)


def test_fp16_planes_range_fallback(ssi, engine):
    """The FP16 planes are scaled from a calibration pass at z = 0 with 64x..128x of headroom.  A sample whose activations
    exceed that (here: z of a few hundred on a subspace that moves every weight) makes the library repeat the call on BF16
    planes; the caller sees correct values and a count in the stats."""
    prob, rng = _rand_problem((96, 128, 128, 10), (1, 1, 0), 300, 20, 4321)
    _setup(engine, prob)
    engine.set_option("path", ssi.PATH_TENSOR)
    Z = rng.standard_normal((20, 6)).astype(np.float32)
    ref, _ = orc.logpost_batch(prob, Z, 0.8)
    np.testing.assert_allclose(engine.logpost(Z, 0.8), ref, rtol=RTOL / 3)
    assert engine.stats().tc_range_fallbacks == 0
    Zbig = Z.copy()
    Zbig[:, 3] *= 3000.0                                   # activations ~1e4 times those at z = 0
    ref_big, _ = orc.logpost_batch(prob, Zbig, 0.8)
    lp = engine.logpost(Zbig, 0.8)
    assert engine.stats().tc_range_fallbacks == 1
    np.testing.assert_allclose(lp, ref_big, rtol=RTOL)
    # the context stays on BF16 planes for this (data, subspace); asynchronous calls report instead of repeating
    np.testing.assert_allclose(engine.logpost(Z, 0.8), ref, rtol=RTOL)
    assert engine.stats().tc_range_fallbacks == 1
    import torch
    engine.set_subspace(prob.W_swa, prob.P)                # a new subspace is calibrated afresh
    dZ = torch.from_numpy(np.ascontiguousarray(Zbig.T)).cuda()
    d_lp = torch.empty(6, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    engine.logpost_dev(dZ.data_ptr(), 6, d_lp.data_ptr(), sigma_m=0.8)
    with pytest.raises(ssi.SsiError) as ei:
        engine.sync()
    assert ei.value.code == -6
    engine.logpost_dev(dZ.data_ptr(), 6, d_lp.data_ptr(), sigma_m=0.8)      # repeated, now on BF16 planes
    engine.sync()
    np.testing.assert_allclose(d_lp.cpu().numpy(), ref_big, rtol=RTOL)
