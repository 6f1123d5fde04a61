#!/usr/bin/env python
"""BASELINE.json configs[4]: a batch of 256 synthetic meshes (C3 generator, seeds 0..255), 2048^2 each, 10k sites,
sharded per mesh over the GPUs of the node (32 per GPU at 8 GPUs): no collective, one context + stream per mesh.

    python tools/bench_batch.py [--meshes 256] [--iters 100]                 (one GPU: all meshes it can hold at a time)
    python -m torch.distributed.run --nproc-per-node 8 ... tools/bench_batch.py --meshes 256

Prints one JSON line (rank 0): meshes/s and Lloyd iterations/s over the whole job, timed with CUDA events around the
enqueue of all loops of a rank (max over ranks), inputs resident; and the same with plain launches instead of graphs.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist
    import _inputs as I
    import surface_remesher_b200 as S
    ap = argparse.ArgumentParser()
    ap.add_argument("--meshes", type=int, default=256)
    ap.add_argument("--grid", type=int, default=2048)
    ap.add_argument("--sites", type=int, default=10000)
    ap.add_argument("--iters", type=int, default=100)
    ap.add_argument("--per-gpu", dest="per_gpu", type=int, default=32, help="meshes resident on a GPU at a time")
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n, k = a.grid, a.sites
    mine = S.shard_meshes(a.meshes, world, rank)
    dev = torch.device("cuda", local)
    res = {}
    for graph in (True, False):
        total_ms, done, hashes = 0.0, 0, []
        for lo in range(0, len(mine), a.per_gpu):
            chunk = mine[lo:lo + a.per_gpu]
            with S.BatchLloyd(n, local, graph=graph) as b:
                for m in chunk:
                    d_t, m_t = I.c3_torch(n, dev, seed=m)
                    dens = d_t.cpu().numpy(); mask = m_t.cpu().numpy()
                    vor = np.empty((n, n, 2), np.int16)
                    S.api._ck(S.lib().srm_seed(vor.ctypes.data, dens.ctypes.data, mask.ctypes.data, k, n, None))
                    b.add(dens, mask, vor)
                b.iterate(10); b.synchronize()          # warm-up (graph capture, first-touch)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                b.iterate(a.iters)
                b.synchronize()
                total_ms += (time.perf_counter() - t0) * 1e3
                done += len(chunk)
                if graph:
                    import hashlib
                    hashes += [hashlib.sha1(np.sort(b.sites(i)).tobytes()).hexdigest()[:12] for i in range(len(chunk))]
        t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        res["graph" if graph else "plain"] = {"ms": float(t.item()), "hashes": hashes}
    if world > 1:
        allh = [None] * world
        dist.all_gather_object(allh, res["graph"]["hashes"])
    else:
        allh = [res["graph"]["hashes"]]
    if rank == 0:
        g, p = res["graph"]["ms"], res["plain"]["ms"]
        print(json.dumps({"metric": "batch_meshes_per_s", "workload": f"{a.meshes} meshes (C3 generator, seeds 0..{a.meshes - 1}), {n}x{n}, {k} sites, "
                          f"{a.iters} Lloyd iterations each (BASELINE.json configs[4])", "n_gpus": world, "meshes_per_gpu_resident": a.per_gpu,
                          "value": a.meshes / (g / 1e3), "unit": "meshes/s", "iterations_per_s": a.meshes * a.iters / (g / 1e3), "ms_total": g,
                          "plain_launches": {"meshes_per_s": a.meshes / (p / 1e3), "ms_total": p},
                          "sites_sha1_per_mesh": sum(allh, [])}), flush=True)
    if world > 1:
        dist.barrier(); dist.destroy_process_group()


if __name__ == "__main__":
    main()
