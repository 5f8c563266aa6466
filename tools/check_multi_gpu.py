"""Sharded-by-example TRAK scoring under torchrun (NCCL) against the single-GPU result on rank 0's device.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/check_multi_gpu.py

Checks the primal path (one all-reduce of the k x k Gram, all-gather of score slices, unequal shards) and the
dual path (N < k: all-gather of the features, N x N system).  Prints one JSON line on rank 0.
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import gadm_b200 as G
from gadm_b200.distributed import shard_range


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    solo = None
    for r in range(world):  # a world-size-1 group per rank: the same code path without exchanges
        g = dist.new_group([r])
        if r == rank:
            solo = g
    res = {"world": world}
    for name, (n, k, t) in {"primal": (3001, 512, 33), "dual": (301, 1024, 17)}.items():
        gen = torch.Generator(device="cpu").manual_seed(7)
        train = torch.randn(n, k, generator=gen).to(dev)
        genphi = torch.randn(t, k, generator=gen).to(dev)
        lo, hi = shard_range(n, world, rank)
        out, sc = G.trak_scores(train[lo:hi].contiguous(), genphi, lam=0.5, gather=True, return_scorer=True)
        ref, sc1 = G.trak_scores(train, genphi, lam=0.5, gather=True, return_scorer=True, group=solo)
        worst = 0.0
        for key in ref:
            assert out[key].shape == ref[key].shape == (n,), (key, out[key].shape)
            worst = max(worst, float((out[key] - ref[key]).abs().max() / ref[key].abs().max()))
        top_same = bool(torch.equal(torch.argsort(out["trak"], descending=True, stable=True)[:20],
                                    torch.argsort(ref["trak"], descending=True, stable=True)[:20]))
        t_ = torch.tensor([worst], device=dev, dtype=torch.float64)
        dist.all_reduce(t_, op=dist.ReduceOp.MAX)
        res[name] = {"n": n, "k": k, "t": t, "dual": bool(sc.dual), "max_rel_diff_vs_single_gpu": float(t_.item()),
                     "top20_identical": top_same}
        assert sc.dual == (name == "dual") and float(t_.item()) < 2e-4, res
    if rank == 0:
        res["ok"] = True
        print(json.dumps(res), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
