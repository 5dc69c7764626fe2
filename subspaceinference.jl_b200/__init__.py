"""subspaceinference.jl_b200 — B200-native drop-in for the batched subspace log-posterior
path of efmanu/SubspaceInference.jl.  Importing this package loads libssi.so (hand-written
sm_100a CUDA behind the C ABI in include/ssi.h); it raises if the library is not built."""
from ._lib import (ACT_IDENTITY, ACT_RELU, ACT_SIGMOID, ACT_TANH, PATH_AUTO, PATH_FUSED, PATH_LAYERED, PATH_NAMES,
                   PATH_BASIS, PATH_TENSOR, TERM_LL, TERM_PRIOR_W, TERM_PRIOR_Z, SsiError, load)
from .engine import Engine
from .flux import ADAM, Chain, DataLoader, Dense, Descent, extract_params, identity, load_params, mse, relu, sigmoid, tanh, train_step
from .api import (auto_encoder_subspace, auto_inference, autoencoder_inference, inference, predictive, shard_rows, sub_inference,
                  subspace_construction, subspace_inference)

load()   # fail loudly at import time when the CUDA library is missing

__all__ = [
    "Engine", "SsiError", "Chain", "Dense", "DataLoader", "ADAM", "Descent", "identity", "relu", "tanh", "sigmoid",
    "mse", "extract_params", "load_params", "train_step", "subspace_construction", "subspace_inference", "sub_inference", "inference", "predictive", "shard_rows", "auto_encoder_subspace", "auto_inference", "autoencoder_inference",
    "TERM_LL", "TERM_PRIOR_W", "TERM_PRIOR_Z", "PATH_AUTO", "PATH_FUSED", "PATH_LAYERED", "PATH_TENSOR", "PATH_BASIS", "PATH_NAMES",
    "ACT_IDENTITY", "ACT_RELU", "ACT_TANH", "ACT_SIGMOID",
]
