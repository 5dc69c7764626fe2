"""Value + gradient of the log-posterior on the wide MLP (BASELINE configs[2] shape): the generic reverse pass with its
GEMMs on the tensor cores (ssi_gemm_tc.cu) against the same pass on the SIMT kernel (option gemm_simt = 1), and the
density evaluation for scale.  Not the headline bench; used for profiles/.
  python profiles/bench_grad.py [--B 12] [--N 60000]"""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import subspaceinference_jl_b200 as ssi  # noqa: E402
import workloads  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=12)
    ap.add_argument("--N", type=int, default=60000)
    ap.add_argument("--simt", type=int, default=1, help="also time the SIMT kernel on this many samples (0: skip)")
    ap.add_argument("--opt", action="append", default=[])
    args = ap.parse_args()
    prob = workloads.make("wide", N=args.N)
    eng = ssi.Engine(0)
    for kv in args.opt:
        k, v = kv.split("=")
        eng.set_option(k, int(v))
    eng.set_model(prob.dims, prob.acts)
    eng.set_data(prob.X, prob.Y)
    eng.set_subspace(prob.W_swa, prob.P)
    rng = np.random.default_rng(5)
    Z = (1e-3 * rng.standard_normal((prob.M, args.B))).astype(np.float32)
    sigma_m = 0.1
    out = {"workload": f"wide {'-'.join(map(str, prob.dims))}, N={prob.N}, M={prob.M}", "B": args.B}

    def timed(fn, reps=2):
        fn()
        eng.sync()
        t0 = time.perf_counter()
        for _ in range(reps):
            r = fn()
        eng.sync()
        return (time.perf_counter() - t0) / reps, r

    flops_per_sample = 3.0 * 2.0 * prob.N * sum(prob.dims[l] * prob.dims[l + 1] for l in range(len(prob.dims) - 1))
    t_tc, (lp_t, g_t) = timed(lambda: eng.logpost_grad(Z, sigma_m))
    st = eng.stats()
    out["grad_tc"] = {"ms_per_sample": t_tc * 1e3 / args.B, "units_per_s": args.B * prob.N / t_tc,
                      "tflops_fp32_equiv": flops_per_sample * args.B / t_tc / 1e12, "gemm_tc_launches": int(st.gemm_tc_launches)}
    if args.simt:
        Bs = min(args.simt, args.B)
        eng.set_option("gemm_simt", 1)
        t_s, (lp_s, g_s) = timed(lambda: eng.logpost_grad(Z[:, :Bs], sigma_m), reps=1)
        eng.set_option("gemm_simt", 0)
        out["grad_simt"] = {"ms_per_sample": t_s * 1e3 / Bs, "units_per_s": Bs * prob.N / t_s,
                            "tflops_fp32": flops_per_sample * Bs / t_s / 1e12}
        out["tc_vs_simt"] = {"lp_rel": float(np.max(np.abs(lp_t[:Bs] - lp_s) / np.abs(lp_s))),
                             "grad_rel_to_norm": float(np.max(np.linalg.norm(g_t[:, :Bs] - g_s, axis=0) / np.linalg.norm(g_s, axis=0)))}
    Bd = max(args.B, 128)
    Zd = (1e-3 * rng.standard_normal((prob.M, Bd))).astype(np.float32)
    t_d, lp_d = timed(lambda: eng.logpost(Zd, sigma_m))
    out["density"] = {"ms_per_sample": t_d * 1e3 / Bd, "units_per_s": Bd * prob.N / t_d}
    lp_d12 = eng.logpost(Z, sigma_m)
    out["grad_lp_vs_density_lp_rel"] = float(np.max(np.abs(lp_t - lp_d12) / np.abs(lp_d12)))
    out["grad_over_density"] = out["grad_tc"]["ms_per_sample"] / out["density"]["ms_per_sample"]
    print(json.dumps(out))
    eng.close()


if __name__ == "__main__":
    main()
