"""Parity of the CUDA labelling (through the C ABI) with the CPU oracle: bit-exact (integer work)."""
import numpy as np
import pytest

import _inputs as I
import _oracle as O

pytestmark = pytest.mark.gpu


def _label(seeds, robust=None):
    """Labels through the C ABI; robust=None checks that the fused band kernel and the robust row path agree."""
    import surface_remesher_b200 as S
    n = seeds.shape[0]
    out = []
    for rb in ([False, True] if robust is None else [robust]):
        with S.Context(n) as c:
            c.set_option("robust_only", rb)
            c.set_site_map(np.ascontiguousarray(seeds))
            c.label()
            out.append(c.get_labels())
    if len(out) == 2:
        assert (out[0] != out[1]).sum() == 0, "band kernel and robust path disagree"
    return out[0]


@pytest.mark.parametrize("n,k,seed", [(256, 1, 0), (256, 2, 1), (256, 50, 2), (256, 4000, 3), (512, 500, 4),
                                      (512, 20000, 5), (1024, 2000, 6), (768, 3000, 7), (2048, 10000, 8)])
def test_label_random(n, k, seed):
    seeds = I.random_sites(n, k, seed)
    got = _label(seeds)
    exp = O.label_exact(seeds)
    bad = (got != exp).any(axis=2)
    assert bad.sum() == 0, f"{bad.sum()} mismatching pixels, first at {np.argwhere(bad)[:5]}"


@pytest.mark.parametrize("n,pitch,off", [(256, 8, 4), (256, 2, 0), (512, 16, 7), (512, 64, 31), (1024, 128, 64)])
def test_label_lattice_ties(n, pitch, off):
    seeds = I.lattice_sites(n, pitch, off)
    got = _label(seeds)
    exp = O.label_exact(seeds)
    assert (got != exp).sum() == 0


def test_label_degenerate_layouts():
    n = 256
    cases = []
    v = np.full((n, n, 2), I.MARK, np.int16); v[100, :, 0] = np.arange(n); v[100, :, 1] = 100; cases.append(v)   # one full row
    v = np.full((n, n, 2), I.MARK, np.int16); v[:, 17, 0] = 17; v[:, 17, 1] = np.arange(n); cases.append(v)       # one full column
    v = np.full((n, n, 2), I.MARK, np.int16)
    for i in range(n): v[i, i] = (i, i)
    cases.append(v)                                                                                               # diagonal
    v = np.full((n, n, 2), I.MARK, np.int16); v[0, 0] = (0, 0); v[n - 1, n - 1] = (n - 1, n - 1); cases.append(v)  # corners
    v = np.full((n, n, 2), I.MARK, np.int16)
    ys, xs = np.mgrid[0:n, 0:n]; v[..., 0] = xs; v[..., 1] = ys; cases.append(v)                                  # every pixel a site
    for s in cases:
        got = _label(s)
        exp = O.label_brute(s)
        assert (got != exp).sum() == 0


def test_band_kernel_overflow_falls_back_to_robust_path():
    """Every column live (all sites in a few rows) overflows the band kernel's shared-memory lists:
    those rows must be handed to the robust path and still be exact."""
    import surface_remesher_b200 as S
    n = 1024
    v = np.full((n, n, 2), I.MARK, np.int16)
    for y in (3, 500, 1000):
        v[y, :, 0] = np.arange(n); v[y, :, 1] = y
    with S.Context(n) as c:
        c.set_site_map(v)
        c.label()
        runs, ovf = c.debug_counts()
        got = c.get_labels()
    assert ovf > 0
    assert (got != O.label_exact(v)).sum() == 0


def test_label_row_band_contexts_agree():
    import surface_remesher_b200 as S
    n = 512
    seeds = I.random_sites(n, 3000, 11)
    exp = O.label_exact(seeds)
    for (r0, r1) in S.row_bands(n, 4):
        with S.Context(n, r0, r1) as c:
            c.set_site_map(np.ascontiguousarray(seeds))
            c.label()
            got = c.get_labels()
        assert (got != exp[r0:r1]).sum() == 0


@pytest.mark.parametrize("n,bands", [(512, [(0, 64), (64, 448), (448, 512)]),          # 2, 12, 2 words: generic column sweep
                                     (1024, [(0, 256), (256, 768), (768, 1024)]),      # 8 / 16 words: segmented scan
                                     (2048, [(0, 1280), (1280, 2048)])])               # 40 / 24 words: looped segments
def test_label_unequal_row_bands(n, bands):
    """Work-balanced partitions give bands of any height (multiples of 64 rows): every carry-scan variant, with the
    per-column edge values that summarise the sites above and below the band."""
    import surface_remesher_b200 as S
    seeds = I.random_sites(n, 6 * n, 21)
    seeds[: n // 3] = I.MARK          # an empty region at the top: columns whose nearest site lies in another band
    exp = O.label_exact(seeds)
    for (r0, r1) in bands:
        with S.Context(n, r0, r1) as c:
            c.set_site_map(np.ascontiguousarray(seeds))
            c.label()
            got = c.get_labels()
        assert (got != exp[r0:r1]).sum() == 0


def test_jfa_matches_cpu_jfa_and_error_rate():
    import surface_remesher_b200 as S
    n = 512
    seeds = I.random_sites(n, 2000, 12)
    steps = [1] + [n >> (i + 1) for i in range(int(np.log2(n)))]  # 1+JFA
    with S.Context(n) as c:
        c.set_site_map(np.ascontiguousarray(seeds))
        got = c.label_jfa(steps)
    exp = O.label_jfa(seeds, steps)
    assert (got != exp).sum() == 0
    exact = O.label_exact(seeds)
    ys, xs = np.mgrid[0:n, 0:n]
    d_j = (got[..., 0].astype(np.int64) - xs) ** 2 + (got[..., 1].astype(np.int64) - ys) ** 2
    d_e = (exact[..., 0].astype(np.int64) - xs) ** 2 + (exact[..., 1].astype(np.int64) - ys) ** 2
    assert (d_j < d_e).sum() == 0           # JFA can never beat the exact distance
    assert (d_j > d_e).mean() < 1e-3        # and is wrong on well under 0.1 % of pixels


def test_no_sites_and_single_site():
    """Empty site set: every label is MARKER (like the oracle); one site: every pixel takes it."""
    import surface_remesher_b200 as S
    n = 256
    empty = np.full((n, n, 2), I.MARK, np.int16)
    with S.Context(n) as c:
        c.set_site_map(empty)
        c.label()
        got = c.get_labels()
    assert (got == I.MARK).all()
    one = empty.copy(); one[200, 13] = (13, 200)
    got = _label(one)
    assert (got[..., 0] == 13).all() and (got[..., 1] == 200).all()


def test_gcvt_with_zero_density_everywhere_keeps_sites():
    """density == 0: no free site may move (gcvt.cu:777), constrained sites stay; W = 0 gives NaN centroids that must
    not crash or move anything."""
    import surface_remesher_b200 as S
    n = 256
    dens = np.zeros((n, n), np.float32)
    mask = np.zeros((n, n), np.uint8)
    seeds = I.random_sites(n, 50, 3)
    vor = seeds.copy()
    S.gCVT(vor, dens, mask, n, 1, 12)
    assert I.site_set(vor) == I.site_set(seeds)
    assert (vor != O.label_exact(seeds)).sum() == 0
