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


@pytest.mark.parametrize("mode", [0, 1, 3])
def test_jfa_matches_cpu_jfa_and_error_rate(mode):
    """mode 0: one plain kernel per pass; 1: srm_jfa.cu (fused shared-memory tile kernel for runs of small steps +
    vectorised far passes, the default); 3: the same with the n > 16384 instantiations (explicit empty-label test)."""
    import surface_remesher_b200 as S
    n = 512
    seeds = I.random_sites(n, 2000, 12)
    steps = [1] + [n >> (i + 1) for i in range(int(np.log2(n)))]  # 1+JFA
    with S.Context(n) as c:
        c.set_option("jfa_mode", mode)
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


def _jfa_case(case, n):
    if case == "lattice":      # thousands of exact ties: the (x, y) part of the key decides
        seeds = np.full((n, n, 2), I.MARK, np.int16)
        for yy in range(3, n, 10):
            for xx in range(5, n, 10):
                seeds[yy, xx] = (xx, yy)
        return seeds, [n // 4, 16, 6, 5, 2, 1, 1]          # far (vector), tile [6, 5, 2, 1], tile [1]
    if case == "border":       # sites only on the grid's edge and corners; steps that are not multiples of 4
        seeds = np.full((n, n, 2), I.MARK, np.int16)
        for q in range(0, n, 17):
            seeds[0, q] = (q, 0); seeds[n - 1, q] = (q, n - 1); seeds[q, 0] = (0, q); seeds[q, n - 1] = (n - 1, q)
        seeds[n - 1, n - 1] = (n - 1, n - 1)
        return seeds, [3, 101, 50, 27, 13, 7, 3, 1]        # tile [3], far (scalar), tile [13], tile [7, 3, 1]
    if case == "one":          # a single site: most pixels stay empty through the early passes
        seeds = np.full((n, n, 2), I.MARK, np.int16)
        seeds[n - 56, 13] = (13, n - 56)
        return seeds, [4, 2, 1, 8, 4, 2, 1, n // 2, 2, 2, 2, 2, 2]
    seeds = I.random_sites(n, 5000, 77)
    return seeds, [1] + [n >> (i + 1) for i in range(int(np.log2(n)))] + [2, 1]   # 1+JFA+2


@pytest.mark.parametrize("mode", [0, 1, 3])
@pytest.mark.parametrize("case", ["lattice", "border", "one", "random"])
def test_jfa_schedules_bit_exact_vs_cpu_jfa(case, mode):
    """Arbitrary schedules (fused runs of every length, scalar and vector far passes, ties, empty regions, sites on the
    border) against the CPU JFA of the same schedule and key."""
    import surface_remesher_b200 as S
    n = 1024 if case == "random" else 512
    seeds, steps = _jfa_case(case, n)
    with S.Context(n) as c:
        c.set_option("jfa_mode", mode)
        c.set_site_map(np.ascontiguousarray(seeds))
        got = c.label_jfa(steps)
        ms = c.label_jfa_timed(steps, mode)
    assert (got != O.label_jfa(seeds, steps)).sum() == 0
    # one launch per pass in mode 0; runs of small steps are single launches otherwise
    assert (len(ms) == len(steps) if mode == 0 else 0 < len(ms) < len(steps)) and all(t >= 0 for t in ms)


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


def _packed(xy):
    xy = np.asarray(xy, np.int32)
    return np.ascontiguousarray((xy[:, 0] & 0xFFFF) | (xy[:, 1] << 16), np.int32)


def test_band_kernel_staging_overflow_fallback():
    """k_band Phase A stages a warp's live columns in its element buffer (C/2 = 512 entries at n <= 8192).  A stripe of
    1024 columns that all stay live at band level (sites at alternating heights: the conservative band-level pruning
    keeps them, the row rounds drop every second one) overflows warp 0's staging area while the whole band list
    (1024 + the few other sites) stays below the band-list capacity and every row's envelope (~520 runs) fits its
    buffer: the list is then assembled by the recompute fallback and consumed by the normal rounds — no row may take
    the robust path.  Labels must equal the oracle's."""
    import surface_remesher_b200 as S
    n = 8192
    xs = np.arange(1024)
    stripe = np.stack([xs, 3000 + 7 * (xs & 1)], 1)
    rng = np.random.default_rng(11)
    ox = rng.choice(np.arange(2048, n), size=400, replace=False)   # distinct columns: at most 400 more live columns
    oy = rng.integers(0, n, 400)
    xy = np.concatenate([stripe, np.stack([ox, oy], 1)]).astype(np.int16)
    with S.Context(n) as c:
        c.set_option("dbg_stats", 1)
        c.set_sites(_packed(xy))
        c.label()
        runs, ovf = c.debug_counts()
        fallback_warps = c.debug_get(6)
        maxlist = c.debug_get(0)
        got = c.get_labels()
    assert fallback_warps > 0, "the staging-overflow branch was not taken"
    assert ovf == 0, f"{ovf} rows took the robust path: the fallback-assembled list was not what the rounds consumed"
    assert maxlist <= 2816
    exp = O.label_band(xy, n, 0, n)
    bad = (got != exp).any(axis=2)
    assert bad.sum() == 0, f"{bad.sum()} mismatching pixels, first at {np.argwhere(bad)[:5]}"


def _random_site_list(n, k, seed):
    rng = np.random.default_rng(seed)
    idx = np.unique(rng.integers(0, n * n, size=int(k * 1.02), dtype=np.int64))
    rng.shuffle(idx)
    idx = idx[:k]
    return np.stack([idx % n, idx // n], 1).astype(np.int16)


def _band_sums(lab, dens_band, xy, n, r0):
    """Direct fp64 sums (W, X, Y) per site id over the pixels of a band, from the oracle's labels."""
    key = xy[:, 1].astype(np.int64) * n + xy[:, 0].astype(np.int64)
    order = np.argsort(key)
    lk = lab[..., 1].astype(np.int64) * n + lab[..., 0].astype(np.int64)
    ids = order[np.searchsorted(key[order], lk.ravel())]
    d = dens_band.astype(np.float64).ravel()
    rows, cols = np.divmod(np.arange(d.size, dtype=np.int64), n)
    K = len(xy)
    W = np.bincount(ids, weights=d, minlength=K)
    X = np.bincount(ids, weights=d * cols, minlength=K)
    Y = np.bincount(ids, weights=d * (rows + r0), minlength=K)
    return W, X, Y


@pytest.mark.parametrize("n,k,r0,r1", [(16384, 250000, 8192, 8448), (32768, 1000000, 16384, 16640),
                                       (32768, 1000000, 0, 128)])
def test_large_grid_band_context_vs_oracle(n, k, r0, r1):
    """The kernel instantiations that only n > 8192 selects (k_band<1,1280> / <1,1792>, band_cap 3584 / 6144) in a
    row-band context of BASELINE configs[3]'s geometry: labels of the band bit-exact against the oracle (rule A2 from
    the site list), and the fused accumulation (k_band) as well as the separate one (k_acc) against direct fp64 sums
    over the oracle's labels."""
    import torch
    import surface_remesher_b200 as S
    from surface_remesher_b200.sharded import _CudaArray
    xy = _random_site_list(n, k, 5)
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev); g.manual_seed(3)
    dens_t = torch.rand((n, n), device=dev, generator=g, dtype=torch.float32) + 0.25
    dens_band = dens_t[r0:r1].cpu().numpy()
    exp = O.label_band(xy, n, r0, r1)
    W, X, Y = _band_sums(exp, dens_band, xy, n, r0)
    with S.Context(n, r0, r1) as c:
        c.set_density(dens_t); c.set_mask(None); c.set_sites(_packed(xy))
        del dens_t
        torch.cuda.empty_cache()
        c.label()
        runs, ovf = c.debug_counts()
        got = c.get_labels()
        bad = (got != exp).any(axis=2)
        assert bad.sum() == 0, f"{bad.sum()} mismatching pixels, first at {np.argwhere(bad)[:5]}"
        assert ovf == 0
        for fused in (False, True):
            if fused:
                c.label_accumulate(True)
            else:
                c.accumulate(True)
            c.synchronize()
            ptr, cnt = c.acc_buffer()
            acc = torch.as_tensor(_CudaArray(ptr, cnt), device="cuda").cpu().numpy().copy()
            for name, ref, col in (("W", W, 0), ("X", X, 1), ("Y", Y, 2)):
                a = acc[col:4 * k:4]
                err = np.abs(a - ref).max() / np.abs(ref).max()
                assert err < 1e-11, (fused, name, err)
            if not fused:
                # clear for the fused pass: a band context's accumulators are cleared by the update, which the test skips
                torch.as_tensor(_CudaArray(ptr, cnt), device="cuda").zero_()
                torch.cuda.synchronize()


@pytest.mark.parametrize("n,k", [(16384, 250000), (32768, 1000000)])
def test_large_grid_whole_context_rows_vs_oracle(n, k):
    """Whole-grid contexts at n > 8192 (k_carry_loop<32>: 512 / 1024 word rows per column): the dense labels stay on the
    device; three 64-row slabs (top, middle, bottom) are compared bit for bit with the oracle."""
    import torch
    import surface_remesher_b200 as S
    xy = _random_site_list(n, k, 9)
    with S.Context(n) as c:
        c.set_sites(_packed(xy))
        c.label()
        lab = torch.empty((n, n, 2), dtype=torch.int16, device="cuda")
        c.get_labels(lab)
        runs, ovf = c.debug_counts()
        for r0 in (0, n // 2 - 64, n - 64):
            got = lab[r0:r0 + 64].cpu().numpy()
            exp = O.label_band(xy, n, r0, r0 + 64)
            bad = (got != exp).any(axis=2)
            assert bad.sum() == 0, f"rows {r0}..: {bad.sum()} mismatching pixels, first at {np.argwhere(bad)[:5]}"
    assert ovf == 0


def test_run_length_pool_exhaustion_is_retried():
    """The run-length pool holds rows * min(n, band buffer) entries.  Two of every three pixels are sites here: 1365 runs
    per row at n = 2048 (pool sized for 1024 per row), every row on the robust path.  The labelling must notice the
    exhausted pool, grow it and come out exact."""
    import surface_remesher_b200 as S
    n = 2048
    v = np.full((n, n, 2), I.MARK, np.int16)
    xs = np.arange(n)
    keep = (xs % 3) != 2
    v[:, keep, 0] = xs[keep][None, :]
    v[:, keep, 1] = np.arange(n)[:, None]
    with S.Context(n) as c:
        c.set_site_map(v)
        c.label()
        runs, ovf = c.debug_counts()
        got = c.get_labels()
    assert runs > n * 1024
    exp = O.label_exact(v)
    assert (got != exp).sum() == 0
