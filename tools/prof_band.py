"""Device-side phase profile of k_band (clock64 per phase, element counters) at the bench config."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, bench
import surface_remesher_b200 as S
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
k = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
dens, mask, vor = bench.make_inputs(n, k, False)
with S.Context(n) as c:
    c.set_density(dens); c.set_mask(mask); c.set_site_map(vor)
    c.iterate(30)
    c.set_option('dbg_stats', 1)
    c.iterate(10)
    print(c.debug_counts(), file=sys.stderr)
    c.set_option('dbg_stats', 0)
    st = c.iterate_profiled(20)
    print({k_: round(v / 20 * 1000, 1) for k_, v in st.items()}, file=sys.stderr)
