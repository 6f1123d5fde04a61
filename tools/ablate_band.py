#!/usr/bin/env python
"""Ablation of the band kernel's accumulate half on the headline workload, in one process: band_fused time of the
profiled loop with the kernel's measurement switches (srm_set_option "dbg_stats": 2 = no atomics, 4 = synthetic site
ids instead of the idmap lookup, 8 = no fp64 prefix loads; results are then meaningless, timings only), and the
label-only kernel (no accumulation at all).

    python tools/ablate_band.py
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# the ablation switches exist only in the -DSRM_MEASURE build: tools/build_variant.sh measure "-DSRM_MEASURE"
os.environ.setdefault("SRM_LIB", os.path.join(ROOT, "build", "variants", "libsrm_measure.so"))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench                      # noqa: E402
import surface_remesher_b200 as S  # noqa: E402

n, k = 8192, 100000
dens, mask, vor = bench.make_inputs(n, k, pinned=False)
out = {}
for dbg in (0, 2, 4, 8, 12, 14):
    with S.Context(n) as c:
        c.set_density(dens); c.set_mask(mask); c.set_site_map(vor)
        c.iterate(30)                      # relax first with the real kernel
        c.set_option("dbg_stats", dbg)
        st = c.iterate_profiled(50)
        out[f"dbg={dbg}"] = round(st["band_fused"] / 50 * 1e3, 1)
with S.Context(n) as c:
    c.set_density(dens); c.set_mask(mask); c.set_site_map(vor)
    c.iterate(30)
    c.label(); c.synchronize()
    t0 = time.perf_counter()
    for _ in range(50):
        c.label()
    c.synchronize()
    out["label_only_call_us (bits + carry + band without accumulate + row kernel)"] = round((time.perf_counter() - t0) / 50 * 1e6, 1)
print(json.dumps(out))
