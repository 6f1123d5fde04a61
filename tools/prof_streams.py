#!/usr/bin/env python
"""One launch of every streaming kernel on the headline grid, for ncu captures:
    ncu --set full --clock-control none --import-source on -k regex:'^k_(prefix|expand|centroid)' -o gpurun_out/prof_streams python tools/prof_streams.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench                      # noqa: E402
import surface_remesher_b200 as S  # noqa: E402

n, k = 8192, 100000
dens, mask, vor = bench.make_inputs(n, k, pinned=False)
with S.Context(n) as c:
    c.set_density(dens); c.set_mask(mask); c.set_site_map(vor)   # k_prefix (default build)
    c.iterate(20, stop_rule=False)
    c.label()
    c.accumulate_dense(None, False)                                # k_expand (default build) + k_centroid_dense
    c.accumulate_dense(None, True)
    c.synchronize()
    print(c.state(), c.debug_counts())
