"""Parity at the sizes BASELINE.json quotes (the configs the headline is measured on), each against the Float64 oracle:

  * Metropolis-Hastings DECISIONS of the tensor-core path on the wide network (configs[2]/[4]), teacher-forced from the
    device's own states with the replayed Philox stream, near ties counted with an ABSOLUTE band on the margin;
  * configs[0] literally: README's DataLoader(X, Y, shuffle=true) has batchsize 1, so T = 10 epochs give K = 1000
    deviation columns for n = 682 parameters;
  * configs[3] at size: n = 10,020,874 parameters, K = T = 100 snapshots, M = 20, against a host Float64 Gram-route oracle.
"""
import time

import numpy as np
import pytest

import ssi_oracle as orc

pytestmark = pytest.mark.gpu

TIE_BAND = 1e-3          # absolute: |lp' - lp + e| below this is a near tie (reported, not compared)
# FP32 arithmetic itself does not resolve margins of 1e-3 at |lp| ~ 1e5..1e6: the library's exact-FP32 CUDA-core path
# (PATH_LAYERED, FP32 FMA chains, FP64 only for the final sum) differs from the Float64 oracle by ~3e-8 |lp| rms in lp' - lp
# (profiles/explore_tc_accuracy.py).  Decisions whose oracle margin is below FP32_BAND * |lp| are therefore counted and
# reported, but only flips OUTSIDE that band fail the test; the same check runs on the FP32 path beside the tensor path.
FP32_BAND = 2.5e-7


def _setup(engine, prob):
    engine.set_model(prob.dims, prob.acts)
    engine.set_data(prob.X, prob.Y)
    engine.set_subspace(prob.W_swa, prob.P)


def _teacher_forced(prob, zt, lt, at, seed, sigma_z, sigma_m, chain_ids, z0=None):
    """From the device's own state at t-1 the oracle (Float64, replayed stream) takes the decision again.
    Returns (decisions, flips, near_ties, max |margin error|, max relative lp error); a flip is a differing decision
    whose oracle margin lies outside the absolute tie band."""
    cache = {}

    def dens(z):
        k = z.tobytes()
        if k not in cache:
            cache[k] = orc.density(prob, z.astype(np.float64), sigma_m)
        return cache[k]

    decisions = flips = flips_fp32 = near = 0
    worst_margin_err, worst_rel = 0.0, 0.0
    for ci, c in enumerate(chain_ids):
        z = np.ascontiguousarray(zt[:, ci, :].T)
        first = z0[:, ci] if z0 is not None else orc.propose_f32(np.zeros(prob.M, np.float32), sigma_z, orc.rng_normals(seed, c, 0, prob.M))
        np.testing.assert_array_equal(z[0], first)
        for t in range(1, z.shape[0]):
            zp = orc.propose_f32(z[t - 1], sigma_z, orc.rng_normals(seed, c, t, prob.M))
            lp_prev, lp_prop = dens(z[t - 1]), dens(zp)
            e = orc.rng_exponential(seed, c, t)
            margin = lp_prop - lp_prev + e
            decisions += 1
            accepted = bool(at[ci, t])
            # the device's own margin, reconstructed from its trace (its lp of the proposal is only visible when accepted)
            if accepted:
                worst_margin_err = max(worst_margin_err, abs((lt[ci, t] - lt[ci, t - 1]) - (lp_prop - lp_prev)))
            worst_rel = max(worst_rel, abs(lt[ci, t] / (lp_prop if accepted else lp_prev) - 1.0))
            if abs(margin) <= TIE_BAND:
                near += 1
                continue
            if accepted != (margin > 0):
                flips += 1
                outside = abs(margin) > FP32_BAND * abs(lp_prev)
                flips_fp32 += outside
                print(f"FLIP chain {c} step {t}: oracle margin {margin:.6g} (|lp| {abs(lp_prev):.3g}), device accepted={accepted}"
                      f"{' OUTSIDE the FP32 band' if outside else ''}")
            expect = zp if accepted else z[t - 1]
            np.testing.assert_array_equal(z[t], expect)
    return decisions, flips, flips_fp32, near, worst_margin_err, worst_rel


@pytest.mark.parametrize("path", ["tensor", "layered"])
@pytest.mark.parametrize("sigma_z", [1e-3, 2e-2])
def test_mh_decisions_wide_reduced(ssi, engine, sigma_z, path):
    """784-1024-1024-10 at N = 6000 (|lp| ~ 1.2e5 with sigma_m = 0.5): 32 chains x 9 steps = 256 decisions per proposal
    scale, on PATH_TENSOR and, beside it, on the exact-FP32 CUDA-core path.  sigma_z = 1e-3 puts the margins at O(0.1..1)
    (the regime of a converged chain, where an error in lp' - lp can actually flip a decision); 2e-2 is the
    far-from-equilibrium regime of BASELINE configs[4]."""
    prob = orc.make_problem("wide", N=6000)
    _setup(engine, prob)
    engine.set_option("path", {"tensor": ssi.PATH_TENSOR, "layered": ssi.PATH_LAYERED}[path])
    C, S, seed, sigma_m = 32, 9, 4242, 0.5
    zt, lt, at = engine.mh_run(C, S, seed, sigma_z=sigma_z, sigma_m=sigma_m, chain_offset=100)
    assert ssi.PATH_NAMES[engine.stats().last_path] == path
    t0 = time.time()
    dec, flips, flips_fp32, near, merr, rel = _teacher_forced(prob, zt, lt, at, seed, sigma_z, sigma_m, range(100, 100 + C))
    print(f"wide N=6000 {path} sigma_z={sigma_z:g}: {dec} decisions, {int(at[:, 1:].sum())} accepted, {flips} flips outside the "
          f"absolute band {TIE_BAND}, {flips_fp32} outside the FP32 band {FP32_BAND:g}*|lp|, {near} near ties, "
          f"max |margin error| {merr:.3g}, max rel lp error {rel:.2e} ({time.time() - t0:.0f} s of oracle)")
    assert dec >= 256
    assert flips_fp32 == 0
    assert rel < 1e-6          # 10x margin to the 1e-5 bar


def test_tensor_path_mh_decisions_full_c3(ssi, engine):
    """Full configs[2] size (N = 60000, |lp| ~ 1e6): 8 chains x 3 steps = 16 teacher-forced decisions."""
    prob = orc.make_problem("wide")
    _setup(engine, prob)
    C, S, seed, sigma_z = 8, 3, 77, 1e-3
    zt, lt, at = engine.mh_run(C, S, seed, sigma_z=sigma_z, sigma_m=1.0, chain_offset=5)
    assert engine.stats().last_path == ssi.PATH_TENSOR
    dec, flips, flips_fp32, near, merr, rel = _teacher_forced(prob, zt, lt, at, seed, sigma_z, 1.0, range(5, 5 + C))
    print(f"wide N=60000: {dec} decisions, {int(at[:, 1:].sum())} accepted, {flips} flips outside {TIE_BAND}, {flips_fp32} outside "
          f"the FP32 band, {near} near ties, max |margin error| {merr:.3g}, max rel lp error {rel:.2e}")
    assert dec >= 16 and flips_fp32 == 0 and rel < 1e-6


def test_gradient_wide_on_tensor_cores(ssi, engine):
    """l_pi_grad (src/space_inference.jl:107) on the wide network at N = 6000: the reverse pass runs its large
    contractions on the tensor cores (ssi_gemm_tc.cu, FP16 planes with per-sample scales) and must match the Float64
    oracle as well as the FP32 CUDA-core pass does (1.7e-5 of the gradient norm here; 8.4e-5 before the forward layers got a
    separate accumulator for their small terms); the value it returns is the density's value."""
    prob = orc.make_problem("wide", N=6000)
    _setup(engine, prob)
    rng = np.random.default_rng(11)
    B, sigma_m = 3, 0.5
    Z = (2e-2 * rng.standard_normal((prob.M, B))).astype(np.float32)
    before = engine.stats().gemm_tc_launches
    lp, g = engine.logpost_grad(Z, sigma_m)
    assert engine.stats().gemm_tc_launches - before >= 5          # 2 forward layers, 2 weight gradients, 1 back-propagated delta
    engine.set_option("gemm_simt", 1)
    lp_s, g_s = engine.logpost_grad(Z, sigma_m)
    engine.set_option("gemm_simt", 0)
    lp_d = engine.logpost(Z, sigma_m)
    worst = worst_s = 0.0
    for b in range(B):
        lp_ref, g_ref = orc.density_and_grad(prob, Z[:, b].astype(np.float64), sigma_m)
        nrm = np.linalg.norm(g_ref)
        worst = max(worst, np.linalg.norm(g[:, b] - g_ref) / nrm)
        worst_s = max(worst_s, np.linalg.norm(g_s[:, b] - g_ref) / nrm)
        assert abs(lp[b] - lp_ref) <= 1e-6 * abs(lp_ref)
    print(f"wide N=6000 gradient: tensor cores {worst:.2e}, CUDA cores {worst_s:.2e} of the gradient norm; "
          f"lp vs density call {np.max(np.abs(lp - lp_d) / np.abs(lp_d)):.1e}")
    assert worst < 5e-5 and worst < 2.5 * worst_s
    np.testing.assert_allclose(lp, lp_d, rtol=1e-6)
    # a sample's gradient does not depend on what else is in the batch (per-sample scales)
    lp1, g1 = engine.logpost_grad(Z[:, 1:2], sigma_m)
    np.testing.assert_array_equal(g1[:, 0], g[:, 1])
    assert lp1[0] == lp[1]


def test_mala_decisions_wide_reduced(ssi, engine):
    """The :mala branch (src/space_inference.jl:117-120) on the wide network at N = 6000, value on the tensor path and
    gradient through the tensor-core GEMM: teacher-forced from the device's own states, the Float64 oracle on the replayed
    stream must propose the same point (up to the FP32-grade gradient in the drift) and take the same decision outside the
    tie bands of the RWMH test above."""
    prob = orc.make_problem("wide", N=6000)
    _setup(engine, prob)
    C, S, seed, sigma_z, sigma_m = 4, 4, 99, 4e-3, 0.5
    before = engine.stats().gemm_tc_launches
    zt, lt, at = engine.mala_run(C, S, seed, sigma_z=sigma_z, sigma_m=sigma_m, chain_offset=7)
    assert engine.stats().gemm_tc_launches > before
    t0 = time.time()
    dec = flips = flips_fp32 = near = 0
    for ci in range(C):
        c = 7 + ci
        z = zt[:, ci, :].T
        for t in range(1, S):
            lp_prev, g_prev = orc.density_and_grad(prob, z[t - 1].astype(np.float64), sigma_m)
            zp = orc.mala_propose_f32(z[t - 1], g_prev, sigma_z, orc.rng_normals(seed, c, t, prob.M))
            lp_prop, g_prop = orc.density_and_grad(prob, zp.astype(np.float64), sigma_m)
            margin = orc.mala_log_alpha(z[t - 1], zp, lp_prev, lp_prop, g_prev, g_prop, sigma_z) + orc.rng_exponential(seed, c, t)
            assert abs(lt[ci, t - 1] - lp_prev) <= 1e-6 * abs(lp_prev)
            dec += 1
            if abs(margin) <= TIE_BAND:
                near += 1
            if bool(at[ci, t]) != (margin > 0):
                flips += abs(margin) > TIE_BAND
                flips_fp32 += abs(margin) > FP32_BAND * abs(lp_prev)
                continue
            expect = zp if margin > 0 else z[t - 1]
            drift = 0.5 * sigma_z ** 2 * np.abs(g_prev).max()
            np.testing.assert_allclose(z[t], expect, rtol=0, atol=2e-4 * drift + np.abs(expect).max() * 2.4e-7 + 1e-12)
    print(f"wide N=6000 MALA: {dec} decisions, {int(at[:, 1:].sum())} accepted, {flips} flips outside the absolute band, "
          f"{flips_fp32} outside the FP32 band, {near} near ties ({time.time() - t0:.0f} s of oracle)")
    assert dec >= 12 and flips_fp32 == 0


def test_readme_literal_batchsize1_K1000(ssi):
    """BASELINE configs[0] as literally configured (README.md:59,74-79): DataLoader(X, Y, shuffle=true) is batchsize 1,
    so T = 10, c = 1 collect K = 1000 deviation columns for n = 682 parameters; opt = ADAM(0.1), M = 3."""
    def build(seed):
        rng = np.random.default_rng(seed)
        X = rng.random((10, 100)).astype(np.float32)
        Y = rng.random((2, 100)).astype(np.float32)
        data = ssi.DataLoader(X, Y, shuffle=True, rng=rng)            # batchsize defaults to 1, as Flux's
        m = ssi.Chain(ssi.Dense(10, 20, rng=rng), ssi.Dense(20, 20, rng=rng), ssi.Dense(20, 2, rng=rng))
        return m, data, ssi.ADAM(0.1)

    cost = lambda mm, x, y: ssi.mse(mm(x), y)
    m, data, opt = build(7)
    assert len(data) == 100
    t0 = time.time()
    W_swa, P = ssi.subspace_construction(m, cost, data, opt, T=10, c=1, M=3, print_freq=100)
    t_dev = time.time() - t0
    assert W_swa.shape == (682,) and P.shape == (682, 3)
    # the same training run again on the host, snapshots handed to the oracle
    m2, data2, opt2 = build(7)
    snaps, ns = [], []
    for i in range(1, 11):
        for x, y in data2:
            ssi.train_step(m2, cost, opt2, x, y)
            snaps.append(ssi.extract_params(m2).copy())
            ns.append(float(i))
    assert len(snaps) == 1000
    W_ref, P_ref, s_ref, _ = orc.construct_from_snapshots(snaps, ns, 3)
    err_w = np.abs(W_swa - W_ref).max() / np.abs(W_ref).max()
    err_p = np.abs(orc.align_signs(P.astype(np.float64), P_ref) - P_ref).max() / np.abs(P_ref).max()
    print(f"README literal: K=1000 > n=682, construction {t_dev:.1f} s, W_swa err {err_w:.2e}, P err {err_p:.2e}, s[:4] = {s_ref[:4]}")
    assert err_w < 1e-4 and err_p < 1e-4


def test_construction_c4_at_size(ssi, engine):
    """BASELINE configs[3] at size: n = 10,020,874 (784-2048-2048-2048-10), K = T = 100 snapshots of a random walk,
    M = 20, against a host Float64 restatement of src/subspace_construction.jl:31,44-52,61-65 by the Gram route."""
    n, K, M = 10_020_874, 100, 20
    rng = np.random.default_rng(4)
    w = (0.05 * rng.standard_normal(n, dtype=np.float32))
    mean = np.zeros(n)                                   # W_swa = zeros (:31), Float64 as in the reference
    A = np.empty((K, n), np.float64)                     # deviation columns, one per row here
    engine.swa_begin(n, K)
    t_push = 0.0
    for t in range(1, K + 1):
        w = w + np.float32(1e-3) * rng.standard_normal(n, dtype=np.float32)
        t0 = time.time()
        engine.swa_push(w, float(t))
        t_push += time.time() - t0
        w64 = w.astype(np.float64)
        mean = (t * mean + w64) / (t + 1.0)              # (:46-47) with n = i/c = t
        A[t - 1] = w64 - mean                            # against the UPDATED mean (:51)
    t0 = time.time()
    W_swa, P, s = engine.swa_finish(M)
    t_fin = time.time() - t0
    st = engine.stats()
    G = np.zeros((K, K))
    for c0 in range(0, n, 1 << 20):
        blk = A[:, c0:c0 + (1 << 20)]
        G += blk @ blk.T
    lam, V = np.linalg.eigh(G)
    order = np.argsort(lam)[::-1]
    lam, V = lam[order], V[:, order]
    s_ref = np.sqrt(np.maximum(lam, 0.0))
    P_ref = np.empty((n, M))
    for c0 in range(0, n, 1 << 20):
        P_ref[c0:c0 + (1 << 20)] = A[:, c0:c0 + (1 << 20)].T @ V[:, :M]
    err_w = np.abs(W_swa - mean).max() / np.abs(mean).max()
    sgn = np.sign(np.einsum("ij,ij->j", P, P_ref))
    err_p = np.abs(P * sgn - P_ref).max() / np.abs(P_ref).max()
    err_s = np.abs(s[:M] / s_ref[:M] - 1).max()
    print(f"C4 at size: push {t_push:.1f} s (host copies included), finish {t_fin:.2f} s, gram_path {st.gram_path}, risk {st.gram_risk:.2e}, "
          f"sweeps {st.jacobi_sweeps}; W_swa err {err_w:.2e}, P err {err_p:.2e}, s err {err_s:.2e}")
    assert err_w < 1e-4 and err_p < 1e-4 and err_s < 1e-5
