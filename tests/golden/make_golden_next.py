"""Golden fixtures for the rows built after the hot path: value+gradient of the log-posterior
(src/space_inference.jl:107), the MALA sampler (:117-120) and the posterior-predictive sweep
(docs/src/nn_example.md:207-216, src/plotting.jl:8-9).  Produced by oracle/ssi_oracle.py (seeded); the reference has no
fixtures of its own.  Run:  python tests/golden/make_golden_next.py
"""
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parents[1] / "oracle"))
import ssi_oracle as orc  # noqa: E402


def main():
    rng = np.random.default_rng(20261019)
    out = {}
    for name, prob, zs, (sm, sp, sz) in (("readme", orc.make_problem("readme"), 0.5, (1.0, 1.0, 1.0)),
                                          ("uci", orc.make_problem("uci", N=600), 0.1, (0.1, 2.0, 0.1))):
        Z = (zs * rng.standard_normal((prob.M, 6))).astype(np.float32)
        lp = np.empty((3, 6)); grad = np.empty((3, prob.M, 6))
        for k, mask in enumerate((1, 3, 7)):
            for b in range(6):
                lp[k, b], grad[k, :, b] = orc.density_and_grad(prob, Z[:, b].astype(np.float64), sm, sp, sz, mask)
        out.update({f"{name}_dims": np.array(prob.dims), f"{name}_acts": np.array(prob.acts), f"{name}_X": prob.X, f"{name}_Y": prob.Y,
                    f"{name}_W_swa": prob.W_swa, f"{name}_P": prob.P, f"{name}_Z": Z, f"{name}_sig": np.array([sm, sp, sz]),
                    f"{name}_lp": lp, f"{name}_grad": grad})
    # MALA on the README problem: 3 chains x 8 steps
    prob = orc.make_problem("readme")
    zt, lt, at, mg = [], [], [], []
    for c in range(3):
        z, lp, acc, margin = orc.mala_chain(prob, 8, 4321, c, sigma_z=0.45, sigma_m=1.0)
        zt.append(z); lt.append(lp); at.append(acc); mg.append(margin)
    out.update(mala_seed=4321, mala_sigma_z=0.45, mala_z=np.stack(zt), mala_lp=np.stack(lt), mala_accept=np.stack(at),
               mala_margin=np.stack(mg))
    # predictive sweep on the README network
    Zp = rng.standard_normal((prob.M, 6)).astype(np.float32)
    Xg = rng.uniform(-1, 2, (prob.dims[0], 11)).astype(np.float32)
    traj, mu, sd = orc.predictive_sweep(prob.dims, prob.acts, prob.W_swa, prob.P, Zp, Xg)
    out.update(pred_Z=Zp, pred_Xg=Xg, pred_traj=traj, pred_mean=mu, pred_std=sd)
    np.savez_compressed(HERE / "next_rows.npz", **out)
    print("written", HERE / "next_rows.npz")


if __name__ == "__main__":
    main()
