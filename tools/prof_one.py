#!/usr/bin/env python
"""Short run of the Lloyd loop on the headline workload, for ncu captures:
    ncu --set full --clock-control none --import-source on -k regex:k_band -s 30 -c 2 -o gpurun_out/prof python tools/prof_one.py [iters]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench                      # noqa: E402
import surface_remesher_b200 as S  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 40
n, k = 8192, 100000
dens, mask, vor = bench.make_inputs(n, k, pinned=False)
with S.Context(n) as c:
    c.set_density(dens); c.set_mask(mask); c.set_site_map(vor)
    c.set_option("dbg_stats", 1)
    c.iterate(iters)
    c.synchronize()
    print(c.state(), c.debug_counts(), "band list max/sum/bands", c.debug_get(0), c.debug_get(1), c.debug_get(2))
