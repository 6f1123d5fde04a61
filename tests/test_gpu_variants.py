"""The builds of each streaming kernel must give identical results (and the oracle's):
  "expand"  runs -> dense labels: 0 = one binary search per 4-pixel group (k_expand), 1 = two-level lookup (k_expand2),
                                  2 = run-start bitmap + popcount prefix (k_expand3)
  "prefix"  fp64 prefix sums:     0 = 128/64-bit stores, 1 = 256-bit stores (STG.E.ENL2.256)
Whichever is the compiled default, both are exercised here; the rest of the suite runs on the default."""
import numpy as np
import pytest

import _inputs as I
import _oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(params=[0, 1, 2], ids=["v0", "v1", "v2"])
def variant(request):
    import surface_remesher_b200 as S
    S.api.set_variant("expand", request.param)
    S.api.set_variant("prefix", min(request.param, 1))
    yield request.param
    S.api.set_variant("expand", -1)
    S.api.set_variant("prefix", -1)


def _labels(seeds, row0=0, row1=None):
    import surface_remesher_b200 as S
    n = seeds.shape[0]
    with S.Context(n, row0, row1 if row1 is not None else n) as c:
        c.set_site_map(np.ascontiguousarray(seeds))
        c.label()
        return c.get_labels()


@pytest.mark.parametrize("n,k,seed", [(256, 1, 0), (256, 4000, 3), (512, 20000, 5), (768, 3000, 7), (2048, 10000, 8)])
def test_expand_variants_random(variant, n, k, seed):
    seeds = I.random_sites(n, k, seed)
    assert (_labels(seeds) != O.label_exact(seeds)).sum() == 0


def test_expand_variants_no_site_and_long_rows(variant):
    """Rows without any run (no site), rows of one run, and rows with more runs than the shared-memory staging area of
    either kernel (a full row of sites: 4096 runs in that row and in its neighbours -> searched in global memory)."""
    n = 4096
    empty = np.full((n, n, 2), I.MARK, np.int16)
    assert (_labels(empty) != empty).sum() == 0
    one = empty.copy(); one[n - 3, 7] = (7, n - 3)
    assert (_labels(one) != O.label_exact(one)).sum() == 0
    seeds = I.random_sites(n, 3000, 21)
    xs = np.arange(n)
    seeds[1000, :, 0] = xs; seeds[1000, :, 1] = 1000          # every pixel of row 1000 is a site
    seeds[3001, ::2, 0] = xs[::2]; seeds[3001, ::2, 1] = 3001  # every second pixel of row 3001
    assert (_labels(seeds) != O.label_exact(seeds)).sum() == 0


def test_expand_variants_band_context_of_a_wide_grid(variant):
    """A 256-row band context of a 16384-wide grid (n / 32 = 512 lookup blocks per row)."""
    n, r0, r1 = 16384, 8192, 8448
    rng = np.random.default_rng(3)
    k = 250000
    idx = rng.choice(n * n, size=k, replace=False)
    ys, xs = np.divmod(idx, n)
    import surface_remesher_b200 as S
    with S.Context(n, r0, r1) as c:
        c.set_sites(((xs & 0xFFFF) | (ys << 16)).astype(np.int32))
        c.label()
        got = c.get_labels()
    exp = O.label_band(np.stack([xs, ys], 1).astype(np.int16), n, r0, r1)
    assert (got != exp).sum() == 0


@pytest.mark.parametrize("kind,n,k,iters", [("c3", 512, 3000, 40), ("uniform", 1024, 2000, 30), ("c3", 2048, 10000, 12)])
def test_whole_gcvt_with_either_prefix_and_expand_build(variant, kind, n, k, iters):
    """Whole gCVT call (prefix sums -> Lloyd loop -> final labelling -> expansion) against the oracle: identical site
    pixels after every iteration's update imply identical prefix sums up to the rounding the update law sees."""
    import surface_remesher_b200 as S
    dens = I.density_uniform(n) if kind == "uniform" else I.density_c3(n)
    mask = None if kind == "uniform" else I.mask_c3(dens)
    seeds, _, _ = O.seed(dens, mask, k)
    exp, it, _, _ = O.gcvt(seeds, dens, mask, iters, stop_rule=1)
    v = seeds.copy()
    st = S.gCVT(v, dens, mask, n, 1, iters)
    S.lib().srm_release_cache()
    assert st["iterations"] == it
    assert (v != exp).sum() == 0


def test_time_kernel_reports_both_streams(variant):
    import surface_remesher_b200 as S
    n = 1024
    dens = I.density_c3(n)
    seeds = I.random_sites(n, 2000, 1)
    with S.Context(n) as c:
        c.set_density(dens)
        c.set_site_map(np.ascontiguousarray(seeds))
        with pytest.raises(S.SrmError):
            c.time_kernel("expand")          # no labelling yet
        c.label()
        assert c.time_kernel("expand", 3) > 0 and c.time_kernel("prefix", 3) > 0
        assert (c.get_labels() != O.label_exact(seeds)).sum() == 0
