#!/usr/bin/env python
"""A/B of kernel variants in ONE process (one `import torch`, one input generation): for the default libsrm.so and every
build/variants/libsrm_<name>.so given, run the same Lloyd loop on the headline workload, print the per-stage times and
check that the site lists after the run are bit-identical to the default build's; the default build is also checked
against the CPU oracle on a small case.

    python tools/ab_inproc.py [name ...]        # e.g.  p4 p5 p6
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench                      # noqa: E402
import surface_remesher_b200 as S  # noqa: E402
import _inputs as I               # noqa: E402
import _oracle as O               # noqa: E402


def use(name):
    if name == "default":
        os.environ.pop("SRM_LIB", None)
    else:
        os.environ["SRM_LIB"] = os.path.join(ROOT, "build", "variants", f"libsrm_{name}.so")
    S.api._lib = None
    S.lib()


def main(names):
    n, k, warm, steps = 8192, 100000, 30, 200
    dens, mask, vor = bench.make_inputs(n, k, pinned=False)
    d5 = I.density_c3(512); m5 = I.mask_c3(d5); s5, _, _ = O.seed(d5, m5, 2000)
    exp5, it5, _, _ = O.gcvt(s5, d5, m5, 30, stop_rule=1)
    ref_sites = None
    for name in ["default"] + list(names):
        use(name)
        v = s5.copy()
        st5 = S.gCVT(v, d5, m5, 512, 1, 30)
        small_ok = bool(st5["iterations"] == it5 and (v != exp5).sum() == 0)
        S.lib().srm_release_cache()
        with S.Context(n) as c:
            c.set_density(dens); c.set_mask(mask); c.set_site_map(vor)
            c.iterate(warm)
            st = c.iterate_profiled(steps)
            runs, ovf = c.debug_counts()
            sites = np.sort(c.get_sites())
        if ref_sites is None:
            ref_sites = sites
        print(json.dumps({"variant": name, "oracle_512_bit_exact": small_ok,
                          "sites_identical_to_default": bool(np.array_equal(sites, ref_sites)),
                          "stages_us": {a: round(b / steps * 1e3, 1) for a, b in st.items()},
                          "it_per_s_profiled_loop": round(steps / (st["iteration"] / 1e3), 1), "robust_rows": ovf}), flush=True)


if __name__ == "__main__":
    main(sys.argv[1:])
