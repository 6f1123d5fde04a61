#!/usr/bin/env python
"""C4 (32768^2, 10^6 sites) on one GPU: stage times + ncu-friendly short loop."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
import surface_remesher_b200 as S
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 20
n, k = 32768, 1000000
dens, mask, vor = bench.make_inputs(n, k, pinned=False)
with S.Context(n) as c:
    c.set_density(dens); c.set_mask(mask); c.set_site_map(vor)
    c.iterate(10)
    st = c.iterate_profiled(iters)
    print(json.dumps({a: round(b / iters, 4) for a, b in st.items()}), c.debug_counts())
