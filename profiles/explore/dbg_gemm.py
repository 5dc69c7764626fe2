import sys, numpy as np
import os; R = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, R); sys.path.insert(0, R + '/oracle')
import subspaceinference_jl_b200 as ssi
import ssi_oracle as orc
dims, acts, N, M, B = (96, 256, 192, 8), (1, 2, 0), 3000, 12, 3
rng = np.random.default_rng(7)
n = orc.n_params(dims)
prob = orc.Problem(dims, acts, rng.standard_normal((dims[0], N)).astype(np.float32), rng.standard_normal((dims[-1], N)).astype(np.float32), orc.glorot_flat(rng, dims), (0.3 * rng.standard_normal((n, M))).astype(np.float32))
Z = (0.5 * rng.standard_normal((M, B))).astype(np.float32)
eng = ssi.Engine(0)
eng.set_model(prob.dims, prob.acts); eng.set_data(prob.X, prob.Y); eng.set_subspace(prob.W_swa, prob.P)
ref = [orc.density_and_grad(prob, Z[:, b].astype(np.float64), 0.7, 1.3, 0.9, 1) for b in range(B)]
gref = np.stack([r[1] for r in ref], axis=1); lref = np.array([r[0] for r in ref])
eng.set_option("gemm_simt", 1)
lp_s, g_s = eng.logpost_grad(Z, 0.7, 1.3, 0.9, mask=1)
eng.set_option("gemm_simt", 0)
print("simt vs oracle", np.max(np.linalg.norm(g_s - gref, axis=0) / np.linalg.norm(gref, axis=0)), np.max(np.abs(lp_s - lref) / np.abs(lref)))
for mask, chunk, prec in ((1, 32, 0), (1, 32, 1), (2, 32, 1), (2, 4, 1), (4, 32, 1), (0, 32, 1), (0, 32, 0)):
    if True:
        eng.set_option("gemm_tc_mask", mask); eng.set_option("gemm_chunk", chunk); eng.set_option("gemm_prec", prec)
        l0 = eng.stats().gemm_tc_launches
        lp_t, g_t = eng.logpost_grad(Z, 0.7, 1.3, 0.9, mask=1)
        print("mask", mask, "chunk", chunk, "prec", prec, "launches", eng.stats().gemm_tc_launches - l0, "grad vs oracle", np.max(np.linalg.norm(g_t - gref, axis=0) / np.linalg.norm(gref, axis=0)),
              "vs simt", np.max(np.linalg.norm(g_t - g_s, axis=0) / np.linalg.norm(g_s, axis=0)), "lp", np.max(np.abs(lp_t - lref) / np.abs(lref)))
