"""Child of tests/test_gpu_multi.py: run under torch.distributed.run with >= 2 ranks (one GPU each).
Row-band Lloyd loop through libsrm with the native collective; every rank compares ITS band of the final labels, the
replicated site list, the iteration count and omega with the single-process CPU oracle."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist
    import _inputs as I
    import _oracle as O
    import surface_remesher_b200 as S
    from surface_remesher_b200.sharded import CudaBandEngine, ShardedLloyd

    mode, n, k, iters, bands_kind = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), sys.argv[5]
    world = int(os.environ["WORLD_SIZE"]); rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dens = I.density_c3(n)
    mask = I.mask_c3(dens)
    seeds, _, _ = O.seed(dens, mask, k)
    if bands_kind == "unequal":
        h = n // world
        cuts = [0] + [min(n - 64 * (world - r), max(64 * r, r * h + (64 if r % 2 else -64))) for r in range(1, world)] + [n]
        bands = [(cuts[r], cuts[r + 1]) for r in range(world)]
    else:
        bands = S.row_bands(n, world)
    r0, r1 = bands[rank]
    eng = CudaBandEngine(n, r0, r1, local)
    sl = ShardedLloyd(n, rank, world, eng, dist, bands)
    if mode == "nccl":
        eng.set_inputs(dens, mask, seeds)     # full arrays on every rank (replicated upload)
    else:
        sl.set_inputs(dens, mask, seeds)      # sharded upload: own rows + exchange of the non-zero bitmap slices
    if mode != "py":
        sl.bind_native_collective(mode)
    sl.run(iters)
    lab = sl.final_labels()
    sites = eng.sites()
    st = eng.state()
    torch.cuda.synchronize()
    exp, it, en, om = O.gcvt(seeds, dens, mask, iters, stop_rule=0)
    got_set = set(zip((sites & 0xFFFF).tolist(), (sites >> 16).tolist()))
    res = {"rank": rank, "rows": [r0, r1], "label_mismatches": int((lab != exp[r0:r1]).any(axis=2).sum()),
           "sites_equal": got_set == I.site_set(exp), "iterations": st["iterations"], "oracle_iterations": it,
           "omega_equal": bool(st["omega"] == np.float32(om)), "num_sites": st["num_sites"]}
    ok = res["label_mismatches"] == 0 and res["sites_equal"] and res["iterations"] == it and res["omega_equal"]
    print(json.dumps(res), flush=True)
    flag = torch.tensor([0 if ok else 1], device="cuda")
    dist.all_reduce(flag)
    eng.close()
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 0 else 1)


if __name__ == "__main__":
    main()
