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


def train_section(eng, stream, dims=(784, 2048, 2048, 2048, 10), N=8192, nb=512, steps=20):
    """One mini-batch step (forward, backward, ADAM, snapshot push) of the C4 model on the device, and the same step
    through the host restatement (torch CPU autograd) for scale."""
    import time
    rng = np.random.default_rng(4)
    acts = (ssi.relu,) * (len(dims) - 2) + (ssi.identity,)
    m = ssi.Chain(*[ssi.Dense(dims[l], dims[l + 1], acts[l], rng=rng) for l in range(len(dims) - 1)])
    X = rng.random((dims[0], N), dtype=np.float32)
    Y = rng.standard_normal((dims[-1], N)).astype(np.float32)
    eng.set_model(m.dims, m.acts)
    eng.set_data(X, Y)
    n = int(ssi.extract_params(m).shape[0])
    eng.swa_begin(n, steps + 4)
    eng.train_begin(ssi.extract_params(m), "adam", 1e-3)
    for w in range(3):
        eng.train_step((w * nb, nb), want_loss=False)
    stream.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for k in range(steps):
        eng.train_step(((k % (N // nb)) * nb, nb), want_loss=False)
        eng.train_snapshot(float(k + 1))
    e1.record(stream)
    stream.synchronize()
    ms = e0.elapsed_time(e1) / steps
    eng.train_end()
    opt = ssi.ADAM(1e-3)
    cost = lambda mm, x, y: ssi.mse(mm(x), y)
    ssi.train_step(m, cost, opt, X[:, :nb], Y[:, :nb])
    t0 = time.perf_counter()
    ssi.train_step(m, cost, opt, X[:, nb:2 * nb], Y[:, nb:2 * nb])
    host_ms = (time.perf_counter() - t0) * 1e3
    flops = 3.0 * nb * 2.0 * sum(dims[l] * dims[l + 1] for l in range(len(dims) - 1))
    return {"model": "-".join(map(str, dims)), "n": n, "batch": nb, "optimiser": "ADAM", "step_plus_snapshot_ms": ms,
            "tflops_fp32": flops / (ms * 1e-3) / 1e12, "host_step_ms": host_ms, "host_threads": torch.get_num_threads(),
            "note": "device: forward/backward GEMMs on the tensor cores where the shape qualifies (ssi_gemm_tc.cu), else SIMT FP32; host: torch CPU autograd + snapshot copy not included"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=10020874)      # 784-2048-2048-2048-10
    ap.add_argument("--K", type=int, default=100)
    ap.add_argument("--M", type=int, default=20)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--opt", action="append", default=[], help="library option key=value (ssi_set_option), repeatable")
    ap.add_argument("--train", action="store_true", help="also time the on-device training step (SURVEY 8(f)-3) on the C4 model")
    a = ap.parse_args()
    peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {"hbm_gbs": 6650.0}
    import os
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:       # row-sharded construction: --n is the TOTAL parameter count, each rank holds n/world rows
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
        r0, r1 = ssi.shard_rows(a.n, rank, world)
        n_total, a.n = a.n, r1 - r0
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    eng = ssi.Engine(local)
    eng.set_stream(stream.cuda_stream)
    for kv in a.opt:
        k, v = kv.split("=")
        eng.set_option(k, int(v))
    g = torch.Generator(device=dev).manual_seed(4 + rank)
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
        if world > 1:
            K = eng.swa_columns()
            G = torch.empty(K * K, dtype=torch.float64, device=dev)
            dist.barrier()
            torch.cuda.synchronize()
            e0.record(stream)
            eng.swa_gram_dev(G.data_ptr(), False)
            dist.all_reduce(G)                                  # on `stream` (torch's current stream): ordered after the Gram
            out_f = eng.swa_finish_gram(a.M, G.data_ptr(), gram_exact=False, want_P=False)
            e1.record(stream)
            stream.synchronize()
            assert out_f is not None, "conditioning check asked for the exact Gram"
            t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            best_fin = min(best_fin, float(t.item()))
        else:
            eng._check(eng._lib.ssi_swa_finish(eng._h, a.M, None, None, None, 0))
            best_fin = min(best_fin, eng.stats().last_ms)
    if world > 1:
        t = torch.tensor([best_push], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        best_push = float(t.item())
        out["n_gpus"], out["n"], out["rows_per_gpu"] = world, n_total, a.n
        a.n = n_total                                           # aggregate bytes over all ranks
    out["swa_push_ms"] = best_push
    out["swa_push_gbs"] = 16.0 * a.n / (best_push * 1e-3) / 1e9
    out["swa_push_frac"] = out["swa_push_gbs"] / (peaks["hbm_gbs"] * world)
    out["finish_ms"] = best_fin                       # gram + jacobi + P
    out["finish_bytes_gb"] = (2 * 4.0 * a.n * a.K + 4.0 * a.n * a.M) / 1e9
    out["finish_gbs"] = out["finish_bytes_gb"] / (best_fin * 1e-3)
    out["finish_frac"] = out["finish_gbs"] / (peaks["hbm_gbs"] * world)
    st = eng.stats()
    out["gram_path"], out["gram_risk"], out["jacobi_sweeps"] = st.gram_path, st.gram_risk, st.jacobi_sweeps
    out["stage_ms"] = {"gram": st.finish_gram_ms, "eigen": st.finish_eigen_ms, "form_p": st.finish_p_ms}
    out["gram_frac"] = 4.0 * a.n * a.K / max(st.finish_gram_ms, 1e-9) / 1e6 / peaks["hbm_gbs"] if world == 1 else None
    out["form_p_frac"] = (4.0 * a.n * a.K + 4.0 * a.n * a.M) / max(st.finish_p_ms, 1e-9) / 1e6 / peaks["hbm_gbs"] if world == 1 else None
    if a.train and world == 1:
        out["train"] = train_section(eng, stream)
    if rank == 0:
        print(json.dumps(out))
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
