"""Host-side stand-ins for the Flux objects the reference's API takes as arguments
(`Chain`, `Dense`, `DataLoader`, `ADAM`, `Descent`) so that calls read like the reference's
README (README.md:46-80).  They describe the model to the CUDA library; they never run the
hot path.  The SGD step inside `subspace_construction` (src/subspace_construction.jl:39-43,
Zygote + Flux.update!) is outside the accelerated path and is delegated to torch autograd
on the host, the same way the reference delegates it to Zygote.
"""
from __future__ import annotations

import math
from typing import Callable, Iterator, Sequence

import numpy as np
import torch

from ._lib import ACT_IDENTITY, ACT_RELU, ACT_SIGMOID, ACT_TANH


def identity(x):
    return x


def relu(x):
    return torch.relu(x)


def tanh(x):
    return torch.tanh(x)


def sigmoid(x):
    return torch.sigmoid(x)


σ = sigmoid
_ACT_CODE = {identity: ACT_IDENTITY, relu: ACT_RELU, tanh: ACT_TANH, sigmoid: ACT_SIGMOID}


class Dense:
    """Flux.Dense(in, out, σ=identity): y = σ.(W*x .+ b), W (out, in) glorot_uniform, b zeros."""

    def __init__(self, din: int, dout: int, act: Callable = identity, *, rng: np.random.Generator | None = None):
        if act not in _ACT_CODE:
            raise ValueError("Error: activation is not supported on the device path")  # String-style error like the reference
        rng = rng or np.random.default_rng()
        lim = math.sqrt(6.0 / (din + dout))
        self.W = torch.tensor(rng.uniform(-lim, lim, size=(dout, din)), dtype=torch.float32, requires_grad=True)
        self.b = torch.zeros(dout, dtype=torch.float32, requires_grad=True)
        self.act = act
        self.din, self.dout = din, dout

    def __call__(self, x):
        return self.act(self.W @ x + self.b[:, None])


class Chain:
    """Flux.Chain of Dense layers."""

    def __init__(self, *layers: Dense):
        if not layers or not all(isinstance(l, Dense) for l in layers):
            raise ValueError("Error: density function is not avaliable for this model")  # src/space_inference.jl:103
        for a, b in zip(layers[:-1], layers[1:]):
            if a.dout != b.din:
                raise ValueError("DimensionMismatch between consecutive Dense layers")
        self.layers = list(layers)

    def __call__(self, x):
        if not torch.is_tensor(x):
            x = torch.as_tensor(np.asarray(x, dtype=np.float32))
        for l in self.layers:
            x = l(x)
        return x

    @property
    def dims(self) -> tuple[int, ...]:
        return (self.layers[0].din,) + tuple(l.dout for l in self.layers)

    @property
    def acts(self) -> tuple[int, ...]:
        return tuple(_ACT_CODE[l.act] for l in self.layers)

    def params(self):
        """Flux.params order: W_1, b_1, W_2, b_2, ..."""
        out = []
        for l in self.layers:
            out += [l.W, l.b]
        return out


def extract_params(model: Chain) -> np.ndarray:
    """Flat Float32 parameter vector in Flux.destructure order: per layer vec(W) (column-major)
    then b (src/libs.jl:19-22)."""
    parts = []
    for l in model.layers:
        parts.append(l.W.detach().numpy().T.reshape(-1))
        parts.append(l.b.detach().numpy().reshape(-1))
    return np.concatenate(parts).astype(np.float32)


def load_params(model: Chain, w: np.ndarray):
    """The `re(W)` half of Flux.destructure (src/libs.jl:56-57), in place."""
    p = 0
    with torch.no_grad():
        for l in model.layers:
            k = l.din * l.dout
            l.W.copy_(torch.as_tensor(np.asarray(w[p:p + k], np.float32).reshape(l.din, l.dout).T.copy()))
            p += k
            l.b.copy_(torch.as_tensor(np.asarray(w[p:p + l.dout], np.float32)))
            p += l.dout


class DataLoader:
    """Flux.Data.DataLoader(X, Y; batchsize=1, shuffle=false): iterates mini-batches along the
    last axis; `.data` is the full (X, Y) tuple that split_data unwraps (src/libs.jl:75-77)."""

    def __init__(self, X, Y, batchsize: int = 1, shuffle: bool = False, rng: np.random.Generator | None = None):
        X, Y = np.asarray(X), np.asarray(Y)
        if X.shape[-1] != Y.shape[-1]:
            raise ValueError("DimensionMismatch: X and Y must have the same number of observations")
        self.data = (X, Y)
        self.batchsize, self.shuffle = int(batchsize), bool(shuffle)
        self._rng = rng or np.random.default_rng()

    def __len__(self):
        n = self.data[0].shape[-1]
        return (n + self.batchsize - 1) // self.batchsize

    def index_batches(self) -> Iterator[np.ndarray]:
        """The observation indices of one epoch's mini-batches (what __iter__ slices with); the on-device training step
        takes these instead of the data."""
        n = self.data[0].shape[-1]
        idx = self._rng.permutation(n) if self.shuffle else np.arange(n)
        for s in range(0, n, self.batchsize):
            yield idx[s:s + self.batchsize]

    def __iter__(self) -> Iterator[tuple[np.ndarray, np.ndarray]]:
        X, Y = self.data
        for sel in self.index_batches():
            yield X[..., sel], Y[..., sel]


def split_data(data: DataLoader):
    """src/libs.jl:75-77."""
    return data.data[0], data.data[1]


class Descent:
    """Flux.Descent(η): p -= η g."""

    def __init__(self, eta: float = 0.1):
        self.eta = eta

    def update(self, params: Sequence[torch.Tensor]):
        with torch.no_grad():
            for p in params:
                if p.grad is not None:
                    p -= self.eta * p.grad
                    p.grad = None


class ADAM:
    """Flux.ADAM(η, (β1, β2)) with ϵ = 1e-8, Flux 0.11 update rule."""

    def __init__(self, eta: float = 0.001, beta=(0.9, 0.999)):
        self.eta, self.beta, self.eps = eta, beta, 1e-8
        self.state: dict[int, tuple] = {}

    def update(self, params: Sequence[torch.Tensor]):
        b1, b2 = self.beta
        with torch.no_grad():
            for p in params:
                if p.grad is None:
                    continue
                mt, vt, bp1, bp2 = self.state.get(id(p), (torch.zeros_like(p), torch.zeros_like(p), b1, b2))
                mt = b1 * mt + (1 - b1) * p.grad
                vt = b2 * vt + (1 - b2) * p.grad * p.grad
                p -= mt / (1 - bp1) / (torch.sqrt(vt / (1 - bp2)) + self.eps) * self.eta
                self.state[id(p)] = (mt, vt, bp1 * b1, bp2 * b2)
                p.grad = None


def mse(pred, y):
    """Flux.Losses.mse."""
    if not torch.is_tensor(y):
        y = torch.as_tensor(np.asarray(y, dtype=np.float32))
    return torch.mean((pred - y) ** 2)


def train_step(model: Chain, cost: Callable, opt, x, y) -> float:
    """gradient(ps) do cost(model, d...) end; Flux.update!(opt, ps, gs)
    (src/subspace_construction.jl:39-43)."""
    loss = cost(model, torch.as_tensor(np.asarray(x, np.float32)), torch.as_tensor(np.asarray(y, np.float32)))
    loss.backward()
    opt.update(model.params())
    return float(loss.detach())
