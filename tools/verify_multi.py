#!/usr/bin/env python
"""Correctness of the sharded Lloyd loop across real GPUs/processes (run under torchrun):
every rank runs `iters` iterations on its row band (fused peer all-reduce or NCCL), rank 0 also runs the whole grid in
one context; the gathered band labels, the site lists and omega must be identical.

    torchrun --nproc-per-node 2 tools/verify_multi.py [n] [sites] [iters] [p2p|nccl|py]
"""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, torch.distributed as dist
import bench
import surface_remesher_b200 as S
from surface_remesher_b200.sharded import CudaBandEngine, ShardedLloyd

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
k = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 37
mode = sys.argv[4] if len(sys.argv) > 4 else "p2p"
world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dens, mask, vor = bench.make_inputs(n, k, pinned=False)
bands = S.row_bands_balanced(n, world, np.nonzero(vor[..., 0] != -32768)[0], unit=64, fixed=0.3)   # unequal heights: the general case
r0, r1 = bands[rank]
eng = CudaBandEngine(n, r0, r1, local)
eng.set_inputs(dens, mask, vor)
sl = ShardedLloyd(n, rank, world, eng, dist, bands)
if mode != "py":
    sl.bind_native_collective(mode)
sl.run(iters)
lab = torch.from_numpy(sl.final_labels().copy().view(np.int32).reshape(r1 - r0, n)).cuda()   # short2 -> int32 for NCCL
sites = np.sort(eng.sites())
st = eng.state()
maxr = max(b[1] - b[0] for b in bands)            # bands differ in height: gather padded, cut below
padded = torch.zeros((maxr, n), dtype=lab.dtype, device="cuda")
padded[: r1 - r0] = lab
parts = [torch.empty_like(padded) for _ in range(world)]
dist.all_gather(parts, padded)
parts = [parts[q][: bands[q][1] - bands[q][0]] for q in range(world)]
cs = torch.tensor([int(np.bitwise_xor.reduce(sites.astype(np.int64) * 2654435761 % (1 << 61))), len(sites)], device="cuda")
allcs = [torch.empty_like(cs) for _ in range(world)]
dist.all_gather(allcs, cs)
if rank == 0:
    full = torch.cat(parts, 0).cpu().numpy().view(np.int16).reshape(n, n, 2)
    with S.Context(n) as c:
        c.set_density(dens); c.set_mask(mask); c.set_site_map(vor)
        c.iterate(iters)
        c.label()
        ref = c.get_labels()
        ref_sites = np.sort(c.get_sites())
        st1 = c.state()
    out = {"n": n, "world": world, "mode": mode, "iters": iters, "band_rows": [b[1] - b[0] for b in bands],
           "label_mismatches": int((full != ref).any(axis=2).sum()),
           "site_lists_identical_across_ranks": bool(all((a == allcs[0]).all().item() for a in allcs)),
           "sites_equal_single_gpu": bool(np.array_equal(sites, ref_sites)),
           "omega_equal": st["omega"] == st1["omega"], "energy_equal": st["energy"] == st1["energy"],
           "p2p_state": st}
    print(json.dumps(out))
dist.barrier(); dist.destroy_process_group()
