"""Option "band_order" of the band kernel (csrc/srm_band.cu "Band order"): the CTAs take the 8-row bands by decreasing
cost of an earlier iteration.  The order must be a permutation whatever the cost profile, must follow the contract of
the CPU model (tests/test_band_order_model.py), and must not change any result: whole runs, band contexts with a
band count that is not a power of two, the graph-replay form and the final labelling are compared with the oracle."""
import numpy as np
import pytest

import _inputs as I
import _oracle as O
from test_band_order_model import band_order_model, check_order

pytestmark = pytest.mark.gpu


def _case(kind, n, k):
    dens = I.density_uniform(n) if kind == "uniform" else I.density_c3(n)
    mask = None if kind == "uniform" else I.mask_c3(dens)
    seeds, _, _ = O.seed(dens, mask, k)
    return dens, mask, seeds


@pytest.mark.parametrize("kind,n,k,iters,stop,graph", [("c3", 512, 3000, 40, True, False), ("uniform", 512, 1000, 25, False, False),
                                                       ("c3", 1024, 5000, 30, True, True), ("c3", 2048, 10000, 23, False, False),
                                                       ("uniform", 2048, 40000, 12, False, True)])
def test_whole_gcvt_bit_exact_with_band_order(kind, n, k, iters, stop, graph):
    import surface_remesher_b200 as S
    dens, mask, seeds = _case(kind, n, k)
    with S.Context(n) as c:
        c.set_option("band_order", 1)
        c.set_option("graph", int(graph))
        c.set_density(dens); c.set_mask(mask); c.set_site_map(seeds)
        st = c.run(iters, stop_rule=stop)
        lab = c.get_labels()
        perm, _ = c.debug_band_order()
    exp, it, en, om = O.gcvt(seeds, dens, mask, iters, stop_rule=int(stop))
    assert st["iterations"] == it
    assert st["omega"] == np.float32(om)
    assert (lab != exp).sum() == 0
    assert sorted(perm.tolist()) == list(range(n // 8))
    if it >= 2:
        assert (perm != np.arange(n // 8)).any()   # the order was rebuilt from a real cost profile


def test_order_follows_the_cost_of_the_previous_iteration():
    """Iteration 0 records the runs per row, iteration 1 rebuilds the order from them (and then overwrites the counts
    with its own): the order read after iteration 1 is the model's order for the costs read after iteration 0."""
    import surface_remesher_b200 as S
    n = 4096
    dens, mask, seeds = _case("c3", n, 20000)
    with S.Context(n) as c:
        c.set_option("band_order", 1)
        c.set_density(dens); c.set_mask(mask); c.set_site_map(seeds)
        perm0, _ = c.debug_band_order()
        assert (perm0 == np.arange(n // 8)).all()
        c.iterate(1)
        perm1, cost0 = c.debug_band_order()
        assert (perm1 == np.arange(n // 8)).all() and cost0.sum() > n
        c.iterate(1)
        perm2, _ = c.debug_band_order()
        check_order(perm2, cost0.astype(np.int64))
        rows = np.zeros(n, np.int64); rows[::8] = cost0          # any rows with these band sums
        assert (perm2 == band_order_model(rows)[0]).all()
        c.iterate(9)   # iterations 2..10: no rebuild
        perm3, _ = c.debug_band_order()
        assert (perm3 == perm2).all()
        c.iterate(1)   # iteration 11 rebuilds
        perm4, _ = c.debug_band_order()
        assert sorted(perm4.tolist()) == list(range(n // 8))
    # the same 12 iterations against the oracle
    with S.Context(n) as c:
        c.set_option("band_order", 1)
        c.set_density(dens); c.set_mask(mask); c.set_site_map(seeds)
        c.iterate(12)
        c.label()
        lab = c.get_labels()
    exp, it, _, _ = O.gcvt(seeds, dens, mask, 12, stop_rule=0)
    assert (lab != exp).sum() == 0


def test_band_contexts_with_band_counts_that_are_not_powers_of_two():
    """Two row-band contexts of one process (72 and 56 bands: padded to 128 and 64 keys in the sort) with the fused
    peer-memory all-reduce, against the oracle."""
    import surface_remesher_b200 as S
    n, cut = 1024, 576
    dens, mask, seeds = _case("c3", n, 5000)
    ctxs = []
    for (r0, r1) in ((0, cut), (cut, n)):
        c = S.Context(n, r0, r1)
        c.set_option("band_order", 1)
        c.set_density(dens); c.set_mask(mask); c.set_site_map(seeds)
        ctxs.append(c)
    blobs = [c.p2p_info() for c in ctxs]
    for r, c in enumerate(ctxs):
        c.p2p_connect(blobs, r, 2)
    iters = 23
    for c in ctxs:
        c.iterate(iters, stop_rule=False)
    labs = []
    for c in ctxs:
        c.label(); labs.append(c.get_labels())
    exp, it, en, om = O.gcvt(seeds, dens, mask, iters, stop_rule=0)
    assert (np.concatenate(labs, 0) != exp).sum() == 0
    for c, nb in zip(ctxs, (cut // 8, (n - cut) // 8)):
        perm, _ = c.debug_band_order()
        assert sorted(perm.tolist()) == list(range(nb))
        assert (perm != np.arange(nb)).any()
    for c in ctxs:
        c.p2p_disconnect()
    for c in ctxs:
        c.close()
