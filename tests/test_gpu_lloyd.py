"""Parity of the CUDA Lloyd loop (through the C ABI) with the CPU oracle.
Integer results (labels, site pixels, iteration counts) must be identical; energies agree to fp64
summation-order noise (rel 1e-10 asserted; the float the control law sees is compared exactly)."""
import glob
import os

import numpy as np
import pytest

import _inputs as I
import _oracle as O

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _case(kind, n, k):
    dens = I.density_uniform(n) if kind == "uniform" else I.density_c3(n)
    mask = None if kind == "uniform" else I.mask_c3(dens)
    seeds, _, _ = O.seed(dens, mask, k)
    return dens, mask, seeds


def _packed_set(packed):
    p = np.asarray(packed, np.int32)
    return set(zip((p & 0xFFFF).tolist(), (p >> 16).tolist()))


@pytest.mark.parametrize("kind,n,k,omega", [("uniform", 256, 400, 2.0), ("c3", 512, 3000, 2.0), ("c3", 512, 3000, 1.37),
                                            ("uniform", 1024, 2000, 2.0), ("c3", 1024, 20000, 1.0),
                                            ("c3", 2048, 10000, 2.0),
                                            # BASELINE.json configs[1] and configs[2] at full size (the O(N) C oracle takes seconds)
                                            ("uniform", 4096, 20000, 2.0), ("c3", 8192, 100000, 2.0),
                                            ("c3", 8192, 100000, 1.21)])
def test_single_step_teacher_forced(kind, n, k, omega):
    """One Lloyd step from the same site map, compared with the oracle: labels, the updated site set, the energy.
    Both forms of the accumulation are checked: the fused band kernel (the product loop, srm_label_accumulate) and the
    separate per-run kernel (srm_label + srm_accumulate); above 2048^2 only the fused one (test time)."""
    import surface_remesher_b200 as S
    dens, mask, seeds = _case(kind, n, k)
    elab, eout, e = O.lloyd_step(seeds, dens, mask, omega)
    with S.Context(n) as c:
        c.set_density(dens); c.set_mask(mask)
        for fused in ([True, False] if n <= 2048 else [True]):
            c.set_site_map(seeds)
            c.set_omega(omega)
            if fused:
                c.label_accumulate(True)
            else:
                c.label()
                c.accumulate(True)
            lab = c.get_labels()
            c.update()
            got = _packed_set(c.get_sites())
            st = c.state()
            assert (lab != elab).sum() == 0
            assert got == I.site_set(eout), fused
            assert st["iterations"] == 1 and st["num_sites"] == len(got)
            assert st["energy"] == np.float32(e) or abs(st["energy"] - e) / e < 1e-6


@pytest.mark.parametrize("kind,n,k,iters,stop,robust", [("uniform", 256, 400, 60, True, False), ("c3", 512, 3000, 40, True, False),
                                                         ("c3", 256, 300, 200, True, False), ("uniform", 512, 1000, 25, False, False),
                                                         ("c3", 1024, 5000, 30, True, False), ("c3", 512, 3000, 40, True, True),
                                                         ("uniform", 2048, 40000, 12, False, False)])
def test_whole_gcvt_bit_exact(kind, n, k, iters, stop, robust):
    import surface_remesher_b200 as S
    dens, mask, seeds = _case(kind, n, k)
    with S.Context(n) as c:
        c.set_option("robust_only", robust)
        c.set_density(dens); c.set_mask(mask); c.set_site_map(seeds)
        st = c.run(iters, stop_rule=stop)
        lab = c.get_labels()
    exp, it, en, om = O.gcvt(seeds, dens, mask, iters, stop_rule=int(stop))
    assert st["iterations"] == it
    assert st["omega"] == np.float32(om)
    assert (lab != exp).sum() == 0
    assert st["num_sites"] == len(I.site_set(exp))


def test_drop_in_entry_point_matches_oracle():
    import surface_remesher_b200 as S
    n = 512
    dens, mask, seeds = _case("c3", n, 2000)
    vor = seeds.copy()
    st = S.gCVT(vor, dens, mask, n, 1, 30)
    exp, it, _, _ = O.gcvt(seeds, dens, mask, 30, stop_rule=1)
    assert st["iterations"] == it and (vor != exp).sum() == 0


def test_centroidalVoronoi_seeds_like_reference():
    import surface_remesher_b200 as S
    n = 256
    dens = I.density_c3(n); mask = I.mask_c3(dens)
    vor = np.empty((n, n, 2), np.int16)
    S.centroidalVoronoi(vor, dens, mask, 300, n, 1, 20)
    seeds, _, _ = O.seed(dens, mask, 300)
    exp, _, _, _ = O.gcvt(seeds, dens, mask, 20, stop_rule=1)
    assert (vor != exp).sum() == 0


def test_sites_merge_like_reference():
    """Dense seeding on a small grid forces collisions: K must shrink exactly as in the oracle."""
    import surface_remesher_b200 as S
    n = 256
    dens = I.density_uniform(n)
    seeds = I.random_sites(n, 20000, 5)
    with S.Context(n) as c:
        c.set_density(dens); c.set_mask(None); c.set_site_map(seeds)
        cur = seeds
        for it in range(4):
            c.label(); c.accumulate(it % 10 == 0); c.update()
            _, cur, _ = O.lloyd_step(cur, dens, None, 2.0)
            got = _packed_set(c.get_sites())
            assert got == I.site_set(cur)
    assert len(got) < 20000


def test_constrained_sites_never_move_and_zero_density_is_rejected():
    import surface_remesher_b200 as S
    n = 256
    dens = I.density_c3(n); mask = I.mask_c3(dens, every=4)
    seeds, _, _ = O.seed(dens, mask, 500)
    with S.Context(n) as c:
        c.set_density(dens); c.set_mask(mask); c.set_site_map(seeds)
        c.iterate(15)
        got = _packed_set(c.get_sites())
    mys, mxs = np.nonzero(mask)
    assert set(zip(mxs.tolist(), mys.tolist())) <= got
    for x, y in got:
        assert mask[y, x] or dens[y, x] != 0


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(G, "ref_label_*.npz"))), ids=os.path.basename)
def test_cuda_labels_vs_reference_golden(path):
    import surface_remesher_b200 as S
    z = np.load(path)
    n = int(z["n"])
    s = z["sites"].astype(np.int32)
    packed = np.ascontiguousarray((s[:, 0] & 0xFFFF) | (s[:, 1] << 16), np.int32)
    with S.Context(n) as c:
        c.set_sites(packed)
        c.label()
        lab = c.get_labels()
    assert (lab != z["labels"]).sum() == 0


def test_two_band_contexts_with_manual_allreduce():
    """Single process, two band contexts on one GPU: summing their accumulators and feeding the sum to both
    reproduces the whole-grid iteration (the N>1 data path without NCCL)."""
    import torch
    import surface_remesher_b200 as S
    from surface_remesher_b200.sharded import CudaBandEngine, ShardedLloyd
    n = 512
    dens, mask, seeds = _case("c3", n, 3000)

    class FakeDist:
        def __init__(self): self.engines = []
        def all_reduce(self, t): pass

    engines = []
    for (r0, r1) in S.row_bands(n, 2):
        e = CudaBandEngine(n, r0, r1, 0)
        e.set_inputs(dens, mask, seeds)
        engines.append(e)
    cur = seeds
    for it in range(12):
        for e in engines:
            e.label(); e.accumulate(it % 10 == 0)
        total = engines[0].acc_tensor() + engines[1].acc_tensor()
        for e in engines:
            e.acc_tensor().copy_(total)
            e.update()
    torch.cuda.synchronize()
    exp, _, _, _ = O.gcvt(seeds, dens, mask, 12, stop_rule=0)
    labs = []
    for e in engines:
        e.label(); labs.append(e.labels())
    full = np.concatenate(labs, 0)
    assert (full != exp).sum() == 0
    for e in engines: e.close()


@pytest.mark.parametrize("n,k,kind", [(4096, 20000, "uniform"), (8192, 100000, "c3")])
def test_full_size_properties(n, k, kind):
    """BASELINE.json sizes: properties that do not need an O(N) oracle pass in Python.
    (1) exact-distance: for sampled pixels the label's distance equals the true nearest-site distance (KD-tree);
    (2) every site labels itself; (3) mass conservation of the accumulators: sum W == sum d, sum X == sum x*d;
    (4) the oracle's separable labelling agrees bit-for-bit on a 256-row band."""
    import surface_remesher_b200 as S
    from scipy.spatial import cKDTree
    dens = I.density_uniform(n) if kind == "uniform" else I.density_c3(n)
    mask = None if kind == "uniform" else I.mask_c3(dens)
    import ctypes as C
    vor = np.empty((n, n, 2), np.int16)
    S.api._ck(S.lib().srm_seed(vor.ctypes.data, dens.ctypes.data, None if mask is None else mask.ctypes.data, k, n, None))
    with S.Context(n) as c:
        c.set_density(dens); c.set_mask(mask); c.set_site_map(vor)
        c.iterate(3)
        c.label()
        lab = c.get_labels()
        c.accumulate(True)
        c.synchronize()   # the context runs on its own non-blocking stream: finish k_acc before torch reads the buffer
        ptr, cnt = c.acc_buffer()
        import torch
        from surface_remesher_b200.sharded import _CudaArray
        acc = torch.as_tensor(_CudaArray(ptr, cnt), device="cuda").cpu().numpy()
        sites = S.api.unpack_sites(c.get_sites()).astype(np.int64)
    K = len(sites)
    # (2)
    assert np.array_equal(lab[sites[:, 1], sites[:, 0]].astype(np.int64), sites)
    # (1)
    rng = np.random.default_rng(0)
    py, px = rng.integers(0, n, 200000), rng.integers(0, n, 200000)
    d_true, _ = cKDTree(sites).query(np.stack([px, py], 1))
    l = lab[py, px].astype(np.int64)
    d_lab = (l[:, 0] - px) ** 2 + (l[:, 1] - py) ** 2
    assert np.array_equal(d_lab, np.rint(d_true ** 2).astype(np.int64))
    # (3)
    kc = (cnt - 4) // 4   # accumulator slots = initial list length (merged sites leave holes, ids are stable)
    W = acc[0:4 * kc:4].sum(); X = acc[1:4 * kc:4].sum(); Y = acc[2:4 * kc:4].sum()
    d64 = dens.astype(np.float64)
    assert abs(W - d64.sum()) / d64.sum() < 1e-12
    assert abs(X - (d64.sum(0) * np.arange(n)).sum()) / X < 1e-12
    assert abs(Y - (d64.sum(1) * np.arange(n)).sum()) / Y < 1e-12
    # (4) oracle on the full map is O(N) C code: affordable (seconds)
    seedmap = np.full((n, n, 2), I.MARK, np.int16)
    seedmap[sites[:, 1], sites[:, 0], 0] = sites[:, 0]; seedmap[sites[:, 1], sites[:, 0], 1] = sites[:, 1]
    exp = O.label_exact(seedmap)
    assert (lab != exp).sum() == 0


def test_extract_sites_matches_delaunay_input_scan():
    """f1: the CDT input points, in the order delaunay.h:46-57 produces them from the label map."""
    import surface_remesher_b200 as S
    n = 512
    dens, mask, seeds = _case("c3", n, 2000)
    scale, l, b = 0.37, -1.5, 2.25
    with S.Context(n) as c:
        c.set_density(dens); c.set_mask(mask); c.set_site_map(seeds)
        c.iterate(7)
        c.label()
        lab = c.get_labels()
        pts = c.extract_sites(mask, scale, l, b)
    exp = []
    for i in range(n):          # x outer
        for j in range(n):      # y inner
            if not mask[j, i] and lab[j, i, 0] == i and lab[j, i, 1] == j:
                exp.append((i * scale + l, j * scale + b))
    assert np.array_equal(pts, np.array(exp))


def test_two_band_contexts_fused_peer_allreduce():
    """The fused all-reduce (update kernel pulls the partial sums from peer accumulators, arrival flags, parity
    double buffering) with two band contexts of one process on one GPU: no collective call, no Python per step."""
    import surface_remesher_b200 as S
    n = 512
    dens, mask, seeds = _case("c3", n, 3000)
    ctxs = []
    for (r0, r1) in S.row_bands(n, 2):
        c = S.Context(n, r0, r1)
        c.set_density(dens); c.set_mask(mask); c.set_site_map(seeds)
        ctxs.append(c)
    blobs = [c.p2p_info() for c in ctxs]
    for r, c in enumerate(ctxs):
        c.p2p_connect(blobs, r, 2)
    iters = 23
    for c in ctxs:
        c.iterate(iters, stop_rule=False)      # both loops are enqueued asynchronously and meet on the GPU
    labs = []
    for c in ctxs:
        c.label(); labs.append(c.get_labels())
    exp, it, en, om = O.gcvt(seeds, dens, mask, iters, stop_rule=0)
    full = np.concatenate(labs, 0)
    assert (full != exp).sum() == 0
    for c in ctxs:
        st = c.state()
        assert st["iterations"] == iters and st["omega"] == np.float32(om)
        c.close()
