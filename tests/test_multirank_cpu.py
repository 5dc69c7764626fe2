"""world_size-2 gloo test of the N>1 host logic: chains are sharded by global id with no
data-path collective, lp is all-gathered, and the result equals the single-rank run.  The
per-rank "device" here is the oracle (CPU container: no GPU), so this covers exactly the
plumbing bench.py / the host mirror add around the C ABI: offsets, gather order, RNG keying."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "oracle"))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def shard_range(n_total: int, rank: int, world: int):
    """[begin, end) of the global chain ids rank owns."""
    per = n_total // world
    return rank * per, (rank + 1) * per if rank < world - 1 else n_total


def _worker(rank, world, port, n_chains, n_steps, seed, out_dir):
    import ssi_oracle as orc
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    prob = orc.make_problem("readme")
    b, e = shard_range(n_chains, rank, world)
    lp_local = np.stack([orc.rwmh_chain(prob, n_steps, seed, c)[1] for c in range(b, e)])     # chain_offset = b
    t = torch.from_numpy(lp_local)
    gathered = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(gathered, t)                     # the only collective on the path: gather lp
    ms = torch.tensor([float(rank + 1)])
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)        # bench.py's max-over-ranks timing
    if rank == 0:
        np.save(Path(out_dir) / "lp.npy", torch.cat(gathered).numpy())
        np.save(Path(out_dir) / "ms.npy", ms.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_matches_single_rank(tmp_path):
    import ssi_oracle as orc
    n_chains, n_steps, seed = 6, 5, 42
    port = _free_port()
    mp.spawn(_worker, args=(2, port, n_chains, n_steps, seed, str(tmp_path)), nprocs=2, join=True)
    lp = np.load(tmp_path / "lp.npy")
    prob = orc.make_problem("readme")
    ref = np.stack([orc.rwmh_chain(prob, n_steps, seed, c)[1] for c in range(n_chains)])
    np.testing.assert_array_equal(lp, ref)           # counter-based RNG on the global chain id: bit-identical
    assert np.load(tmp_path / "ms.npy")[0] == 2.0


def test_shard_ranges_cover_everything():
    for n, w in [(65536, 8), (10, 3), (7, 2), (4096, 1)]:
        ids = []
        for r in range(w):
            b, e = shard_range(n, r, w)
            ids += list(range(b, e))
        assert ids == list(range(n))
