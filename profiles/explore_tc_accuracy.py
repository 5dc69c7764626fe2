"""Exploration (not a test): how far is the tensor path's lp, and the MH margin lp' - lp, from the Float64 oracle on the
wide network?  Prints absolute and relative errors for a few proposal scales, for every precision mode the library has.
  python profiles/explore_tc_accuracy.py [--N 6000] [--pairs 24] [--modes 0,1]"""
import argparse
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "oracle"))
import ssi_oracle as orc  # noqa: E402
import subspaceinference_jl_b200 as ssi  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--N", type=int, default=6000)
    ap.add_argument("--pairs", type=int, default=24)
    ap.add_argument("--sigma_m", type=float, default=0.5)
    ap.add_argument("--modes", default="")
    a = ap.parse_args()
    prob = orc.make_problem("wide", N=a.N)
    rng = np.random.default_rng(1)
    eng = ssi.Engine(0)
    eng.set_model(prob.dims, prob.acts)
    eng.set_data(prob.X, prob.Y)
    eng.set_subspace(prob.W_swa, prob.P)
    modes = [int(m) for m in a.modes.split(",")] if a.modes else [None]
    for scale_z in (0.1, 0.02):
        Z = (scale_z * rng.standard_normal((prob.M, a.pairs))).astype(np.float32)
        t0 = time.time()
        ref = np.array([orc.density(prob, Z[:, b].astype(np.float64), a.sigma_m) for b in range(a.pairs)])
        t_or = (time.time() - t0) / a.pairs
        for sz in (1e-2, 1e-3, 1e-4):
            Zp = (Z + sz * rng.standard_normal(Z.shape)).astype(np.float32)
            refp = np.array([orc.density(prob, Zp[:, b].astype(np.float64), a.sigma_m) for b in range(a.pairs)])
            for mode in modes:
                if mode is not None:
                    eng.set_option("tc_precision", mode)
                for path in (ssi.PATH_TENSOR, ssi.PATH_LAYERED):
                    eng.set_option("path", path)
                    lp = eng.logpost(Z, a.sigma_m)
                    lpp = eng.logpost(Zp, a.sigma_m)
                    d_dev, d_ref = lpp - lp, refp - ref
                    print(f"z~{scale_z} sigma_z={sz:g} mode={mode} path={ssi.PATH_NAMES[path]:8s} |lp|~{np.abs(ref).mean():.3g} "
                          f"max|lp err|={np.abs(lp - ref).max():.3g} (rel {np.abs(lp / ref - 1).max():.2e}) "
                          f"margin: |dlp| median {np.median(np.abs(d_ref)):.3g}, max|err|={np.abs(d_dev - d_ref).max():.3g} "
                          f"rms={np.sqrt(np.mean((d_dev - d_ref) ** 2)):.3g}", flush=True)
        print(f"oracle: {t_or:.2f} s per density at N={a.N}")
    eng.close()


if __name__ == "__main__":
    main()
