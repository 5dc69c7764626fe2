"""Synthetic inputs of the BASELINE.json configs, NumPy only.  bench.py builds its workloads from here, so
that the timed arm does not touch oracle/ (the oracle is the checker: it is imported by bench.py only for the post-timing
sanity check and for the CPU baseline legs).  The generators are seeded and produce exactly the arrays
oracle/ssi_oracle.make_problem produces (tests/test_oracle.py checks that), so tests and bench describe the same problems."""
from __future__ import annotations

import math
from types import SimpleNamespace

import numpy as np

CONFIGS = {
    # name: (dims, acts, M, N, seed)          acts: 0 identity, 1 relu, 2 tanh, 3 sigmoid
    "readme": ((10, 20, 20, 2), (0, 0, 0), 3, 100, 1234),          # README.md:52-79
    "uci": ((13, 50, 1), (1, 0), 5, 10000, 2024),                  # BASELINE configs[1]
    "wide": ((784, 1024, 1024, 10), (1, 1, 0), 20, 60000, 31337),  # BASELINE configs[2] / [4]
}


def n_params(dims) -> int:
    return int(sum(dims[l] * dims[l + 1] + dims[l + 1] for l in range(len(dims) - 1)))


def _glorot_flat(rng, dims) -> np.ndarray:
    """Flux's default Dense init (glorot_uniform weights, zero bias) in Flux.destructure order: vec(W_l) column-major, b_l."""
    parts = []
    for l in range(len(dims) - 1):
        din, dout = dims[l], dims[l + 1]
        lim = math.sqrt(6.0 / (din + dout))
        parts += [rng.uniform(-lim, lim, size=(dout, din)).reshape(-1, order="F"), np.zeros(dout)]
    return np.concatenate(parts).astype(np.float32)


def _forward(w, dims, acts, X):
    h, off = np.asarray(X, np.float64), 0
    w = np.asarray(w, np.float64)
    for l, act in enumerate(acts):
        din, dout = dims[l], dims[l + 1]
        W = w[off:off + din * dout].reshape(dout, din, order="F")
        off += din * dout
        h = W @ h + w[off:off + dout][:, None]
        off += dout
        h = (h, np.maximum(h, 0.0), np.tanh(h), 1.0 / (1.0 + np.exp(-h)))[act]
    return h


def make(name: str, N: int | None = None, seed: int | None = None) -> SimpleNamespace:
    dims, acts, M, n_def, seed_def = CONFIGS[name]
    N = n_def if N is None else N
    rng = np.random.default_rng(seed_def if seed is None else seed)
    n = n_params(dims)
    W_swa = _glorot_flat(rng, dims)
    if name == "readme":
        X = rng.random((dims[0], N), dtype=np.float32)
        Y = rng.random((dims[-1], N), dtype=np.float32)
        scales = 10.0 ** (-np.arange(M) / M)
    elif name == "uci":
        X = rng.standard_normal((dims[0], N)).astype(np.float32)
        w_true = _glorot_flat(rng, dims)
        Y = (_forward(w_true, dims, acts, X) + 0.1 * rng.standard_normal((dims[-1], N))).astype(np.float32)
        W_swa = (w_true + 0.01 * rng.standard_normal(n)).astype(np.float32)
        scales = np.array([1.0, 0.5, 0.25, 0.12, 0.06])
    else:
        X = rng.random((dims[0], N), dtype=np.float32)
        lab = rng.integers(0, dims[-1], size=N)
        Y = -np.ones((dims[-1], N), np.float32)
        Y[lab, np.arange(N)] = 1.0
        scales = 0.5 ** np.arange(1, M + 1)
    Q, _ = np.linalg.qr(rng.standard_normal((n, M)))
    P = (Q * np.asarray(scales)[None, :]).astype(np.float32)
    return SimpleNamespace(dims=dims, acts=acts, X=X, Y=Y, W_swa=W_swa, P=P, M=M, N=N, n=n)
