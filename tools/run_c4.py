#!/usr/bin/env python
"""BASELINE.json configs[3]: 32768^2 grid, 1M sites (C3 generator).  Functional + timing run on ONE GPU
(the reference cannot run this size at all: no schedule-table entry, gcvt.cu:842-847).
Checks: sampled exact-distance property against a KD-tree, every site labels itself, mass conservation."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import bench
import surface_remesher_b200 as S
from surface_remesher_b200.sharded import _CudaArray

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
k = int(sys.argv[2]) if len(sys.argv) > 2 else 1000000
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 20
t0 = time.time()
dens, mask, vor = bench.make_inputs(n, k, pinned=False)
print(f"inputs {time.time()-t0:.1f}s", file=sys.stderr)
out = {"n": n, "sites": k}
with S.Context(n) as c:
    t0 = time.time(); c.set_density(dens); c.set_mask(mask); c.set_site_map(vor); c.synchronize()
    out["upload_s"] = time.time() - t0
    c.iterate(3); c.synchronize()
    st = c.iterate_profiled(iters)
    out["ms_per_iteration"] = st["iteration"] / iters
    out["it_per_s"] = iters / (st["iteration"] / 1e3)
    out["stages_us"] = {a: round(b / iters * 1e3, 1) for a, b in st.items()}
    c.label()
    runs, ovf = c.debug_counts()
    out["runs"], out["robust_rows"] = runs, ovf
    c.accumulate(False)
    c.synchronize()
    ptr, cnt = c.acc_buffer()
    acc = torch.as_tensor(_CudaArray(ptr, cnt), device="cuda").cpu().numpy()
    sites = S.api.unpack_sites(c.get_sites()).astype(np.int64)
    out["live_sites"] = len(sites)
    # properties on a band of rows (dense labels of the whole grid would be 4 GiB on the host)
    lab = c.get_labels()
from scipy.spatial import cKDTree
rng = np.random.default_rng(0)
py, px = rng.integers(0, n, 300000), rng.integers(0, n, 300000)
d_true, _ = cKDTree(sites).query(np.stack([px, py], 1))
l = lab[py, px].astype(np.int64)
d_lab = (l[:, 0] - px) ** 2 + (l[:, 1] - py) ** 2
out["exact_distance_samples_ok"] = bool(np.array_equal(d_lab, np.rint(d_true ** 2).astype(np.int64)))
out["sites_label_themselves"] = bool(np.array_equal(lab[sites[:, 1], sites[:, 0]].astype(np.int64), sites))
kc = (cnt - 4) // 4
W = acc[0:4 * kc:4].sum()
out["mass_rel_err"] = float(abs(W - dens.astype(np.float64).sum()) / dens.astype(np.float64).sum())
print(json.dumps(out))
