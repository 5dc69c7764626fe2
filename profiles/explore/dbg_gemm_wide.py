import os, sys, numpy as np
R = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, R); sys.path.insert(0, R + '/oracle')
import subspaceinference_jl_b200 as ssi
import ssi_oracle as orc
prob = orc.make_problem("wide", N=6000)
eng = ssi.Engine(0)
eng.set_model(prob.dims, prob.acts); eng.set_data(prob.X, prob.Y); eng.set_subspace(prob.W_swa, prob.P)
rng = np.random.default_rng(11)
B, sigma_m = 2, 0.5
Z = (2e-2 * rng.standard_normal((prob.M, B))).astype(np.float32)
ref = [orc.density_and_grad(prob, Z[:, b].astype(np.float64), sigma_m) for b in range(B)]
gref = np.stack([r[1] for r in ref], axis=1)
def err(g): return np.max(np.linalg.norm(g - gref, axis=0) / np.linalg.norm(gref, axis=0))
eng.set_option("gemm_simt", 1); print("simt", err(eng.logpost_grad(Z, sigma_m)[1])); eng.set_option("gemm_simt", 0)
eng.set_option("gemm_split_acc", 2)
for fc in (32, 16, 8, 4):
    eng.set_option("gemm_fwd_chunk", fc)
    print("split_acc 2 fwd_chunk", fc, err(eng.logpost_grad(Z, sigma_m)[1]))
