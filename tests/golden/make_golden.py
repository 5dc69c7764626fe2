"""Generates the committed golden fixtures from the CPU oracle (seeded, deterministic).

The reference ships no golden vectors and cannot run here (no Julia), so these are produced
by the restatement in oracle/ssi_oracle.py and pinned by the independent checks in
tests/test_oracle.py.  Run:  python tests/golden/make_golden.py
"""
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parents[1] / "oracle"))
import ssi_oracle as orc  # noqa: E402


def save_logpost(name, prob, Z, sm, sp, sz):
    _, terms = orc.logpost_batch(prob, Z, sm, sp, sz, mask=7)
    np.savez_compressed(HERE / f"logpost_{name}.npz", dims=np.array(prob.dims), acts=np.array(prob.acts), X=prob.X, Y=prob.Y,
                        W_swa=prob.W_swa, P=prob.P, Z=Z.astype(np.float32), sigma_m=sm, sigma_p=sp, sigma_z=sz, terms=terms)


def main():
    rng = np.random.default_rng(20261018)
    # README config (C1)
    prob = orc.make_problem("readme")
    Z = rng.standard_normal((prob.M, 16)).astype(np.float32)
    save_logpost("readme", prob, Z, 1.0, 1.0, 1.0)
    # UCI-style (C2) at a size the oracle finishes instantly
    prob = orc.make_problem("uci", N=1000)
    Z = (0.1 * rng.standard_normal((prob.M, 24))).astype(np.float32)
    save_logpost("uci_small", prob, Z, 0.1, 2.0, 0.1)
    # wide-style (C3) shrunk: 96-128-128-10 relu, N=300
    dims, acts, M, N = (96, 128, 128, 10), (1, 1, 0), 20, 300
    r2 = np.random.default_rng(31337)
    n = orc.n_params(dims)
    W_swa = orc.glorot_flat(r2, dims)
    X = r2.random((dims[0], N), dtype=np.float32)
    lab = r2.integers(0, dims[-1], size=N)
    Y = -np.ones((dims[-1], N), np.float32)
    Y[lab, np.arange(N)] = 1.0
    P = orc.synthetic_subspace(r2, n, M, 0.5 ** np.arange(1, M + 1))
    prob = orc.Problem(dims, acts, X, Y, W_swa, P)
    Z = (0.1 * rng.standard_normal((M, 8))).astype(np.float32)
    save_logpost("wide_small", prob, Z, 1.0, 1.0, 1.0)

    # RWMH on the README config: 4 chains x 10 steps (itr=10 as README)
    prob = orc.make_problem("readme")
    n_chains, n_steps, seed = 4, 10, 1234
    zt, lt, at = [], [], []
    for c in range(n_chains):
        z, lp, acc, _ = orc.rwmh_chain(prob, n_steps, seed, c, 1.0, 1.0)
        zt.append(z); lt.append(lp); at.append(acc)
    np.savez_compressed(HERE / "mh_readme.npz", dims=np.array(prob.dims), acts=np.array(prob.acts), X=prob.X, Y=prob.Y,
                        W_swa=prob.W_swa, P=prob.P, n_chains=n_chains, n_steps=n_steps, seed=seed, sigma_z=1.0, sigma_m=1.0,
                        z_trace=np.stack(zt), lp_trace=np.stack(lt), accept=np.stack(at))

    # construction: n=682 (README net), 3 epochs x 5 mini-batches, c=1 -> K=15, M=3
    r3 = np.random.default_rng(77)
    n, epochs, nb, M = 682, 3, 5, 3
    w = orc.glorot_flat(r3, (10, 20, 20, 2)).astype(np.float64)
    snaps, ns = [], []
    for i in range(1, epochs + 1):
        for _ in range(nb):
            w = w + 0.05 * r3.standard_normal(n) - 0.01 * w
            snaps.append(w.astype(np.float32))
            ns.append(i / 1)
    W_swa, P, s, _ = orc.construct_from_snapshots(snaps, ns, M)
    np.savez_compressed(HERE / "construct_small.npz", snapshots=np.stack(snaps), n_scalars=np.array(ns), M=M,
                        W_swa=W_swa, P=P, s=s)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
