"""Batch mode (BASELINE.json configs[4]) and the CUDA-graph replay of the loop: parity with the oracle."""
import numpy as np
import pytest

import _inputs as I
import _oracle as O

pytestmark = pytest.mark.gpu


def test_graph_replay_matches_plain_loop_and_oracle():
    """srm_iterate with option "graph": blocks of 10 iterations replayed as one CUDA graph; unaligned head and tail run
    as plain launches.  Labels, iteration count, omega: identical to the oracle, with and without the stopping rule."""
    import surface_remesher_b200 as S
    n, k = 512, 3000
    dens = I.density_c3(n); mask = I.mask_c3(dens)
    seeds, _, _ = O.seed(dens, mask, k)
    for iters, stop in ((47, False), (200, True)):
        exp, it, en, om = O.gcvt(seeds, dens, mask, iters, stop_rule=int(stop))
        with S.Context(n) as c:
            c.set_option("graph", 1)
            c.set_density(dens); c.set_mask(mask); c.set_site_map(seeds)
            c.iterate(3, stop)            # misalign: 3, then graphs from iteration 10 on
            c.iterate(iters - 3, stop)
            c.label()
            st = c.state()
            lab = c.get_labels()
        assert st["iterations"] == it and st["omega"] == np.float32(om)
        assert (lab != exp).sum() == 0


def test_batch_of_meshes_matches_oracle():
    """Four independent problems of different densities / seeds on one GPU, enqueued back to back on their own
    streams: each result equals the oracle's run of that problem."""
    import surface_remesher_b200 as S
    n, k, iters = 512, 1500, 60
    probs = []
    for seed in range(4):
        dens = I.density_c3(n, seed=100 + seed); mask = I.mask_c3(dens)
        seeds, _, _ = O.seed(dens, mask, k, state=seed * 7919)
        probs.append((dens, mask, seeds))
    with S.BatchLloyd(n) as b:
        for d, m, s in probs:
            b.add(d, m, s)
        stats = b.run(iters, stop_rule=True)
        labs = [b.labels(i) for i in range(len(probs))]
    for (d, m, s), st, lab in zip(probs, stats, labs):
        exp, it, en, om = O.gcvt(s, d, m, iters, stop_rule=1)
        assert st["iterations"] == it and st["omega"] == np.float32(om)
        assert (lab != exp).sum() == 0


def test_shard_meshes_partition():
    import surface_remesher_b200 as S
    for total, world in ((256, 8), (10, 4), (3, 8)):
        parts = [S.shard_meshes(total, world, r) for r in range(world)]
        assert sorted(sum(parts, [])) == list(range(total))
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
