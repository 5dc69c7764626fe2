"""Construction-stream measurement (BASELINE config 4): SWA push, Gram + eigen + P on a ~10M-parameter
flat vector.  Prints achieved GB/s against the measured HBM peak.  Not the headline bench (that is
bench.py); used for profiles/ and for ncu launch lists.
  python profiles/bench_streams.py [--n 10020874] [--K 100] [--M 20]"""
import argparse
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import subspaceinference_jl_b200 as ssi  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=10020874)      # 784-2048-2048-2048-10
    ap.add_argument("--K", type=int, default=100)
    ap.add_argument("--M", type=int, default=20)
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {"hbm_gbs": 6650.0}
    dev = torch.device("cuda", 0)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    eng = ssi.Engine(0)
    eng.set_stream(stream.cuda_stream)
    g = torch.Generator(device=dev).manual_seed(4)
    w = 0.05 * torch.randn(a.n, device=dev, generator=g)
    snaps = []
    for t in range(a.K):                       # W_t = W_0 + cumulative 1e-3 N(0,1) steps
        w = w + 1e-3 * torch.randn(a.n, device=dev, generator=g)
        snaps.append(w.clone())
    out = {"n": a.n, "K": a.K, "M": a.M, "hbm_peak_gbs": peaks["hbm_gbs"]}
    best_push, best_fin = 1e9, 1e9
    for rep in range(a.reps):
        eng.swa_begin(a.n, a.K)
        stream.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for t, s in enumerate(snaps):
            eng.swa_push_dev(s.data_ptr(), float(t + 1))
        e1.record(stream)
        stream.synchronize()
        best_push = min(best_push, e0.elapsed_time(e1) / a.K)
        eng._check(eng._lib.ssi_swa_finish(eng._h, a.M, None, None, None, 0))
        best_fin = min(best_fin, eng.stats().last_ms)
    out["swa_push_ms"] = best_push
    out["swa_push_gbs"] = 16.0 * a.n / (best_push * 1e-3) / 1e9
    out["swa_push_frac"] = out["swa_push_gbs"] / peaks["hbm_gbs"]
    out["finish_ms"] = best_fin                       # gram + jacobi + P
    out["finish_bytes_gb"] = (2 * 4.0 * a.n * a.K + 4.0 * a.n * a.M) / 1e9
    out["finish_gbs"] = out["finish_bytes_gb"] / (best_fin * 1e-3)
    out["finish_frac"] = out["finish_gbs"] / peaks["hbm_gbs"]
    st = eng.stats()
    out["gram_path"], out["gram_risk"], out["jacobi_sweeps"] = st.gram_path, st.gram_risk, st.jacobi_sweeps
    print(json.dumps(out))
    eng.close()


if __name__ == "__main__":
    main()
