"""The stand-alone centroid pass over a DENSE label map (srm_centroid.cu, srm_accumulate_dense): north_star's form of
pbaCVDComputeCentroid (gcvt.cu:1008-1023) + pbaCVDCalcEnergy (gcvt.cu:1059-1083).  It is not the loop's path (the band
kernel accumulates from its runs); it must nevertheless give the same per-site sums as the run-based kernel and the
oracle's direct sums, and the same updated sites, bit for bit, when it takes the place of srm_accumulate in a step."""
import numpy as np
import pytest

import _inputs as I
import _oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(params=[0, 1], ids=["v0", "v1"], autouse=True)
def variant(request):
    """Both builds of the kernel (srm_centroid.cu: 0 = k_centroid_dense, 1 = k_centroid_dense2, fewer instructions)."""
    import surface_remesher_b200 as S
    S.api.set_variant("centroid", request.param)
    yield request.param
    S.api.set_variant("centroid", -1)


def _case(kind, n, k):
    dens = I.density_uniform(n) if kind == "uniform" else I.density_c3(n)
    mask = None if kind == "uniform" else I.mask_c3(dens)
    seeds, _, _ = O.seed(dens, mask, k)
    return dens, mask, seeds


def _acc(c, zero=False):
    import torch
    from surface_remesher_b200.sharded import _CudaArray
    c.synchronize()
    ptr, cnt = c.acc_buffer()
    t = torch.as_tensor(_CudaArray(ptr, cnt), device="cuda")
    a = t.cpu().numpy().copy()
    if zero:
        t.zero_()
        torch.cuda.synchronize()
    return a


def _direct_sums(lab, dens, sites_packed, row0=0):
    """fp64 sums of d, x*d, y*d per site id and the energy, straight from a label map (numpy)."""
    rows, n = dens.shape
    key = (lab[..., 0].astype(np.int64) & 0xFFFF) | (lab[..., 1].astype(np.int64) << 16)
    sp = np.asarray(sites_packed, np.int64) & 0xFFFFFFFF
    order = np.argsort(sp)
    pos = np.searchsorted(sp[order], key.ravel())
    pos = np.clip(pos, 0, len(sp) - 1)
    hit = sp[order][pos] == key.ravel()
    ids = order[pos]
    d = dens.astype(np.float64).ravel()
    xs = np.tile(np.arange(n, dtype=np.float64), rows)
    ys = np.repeat(np.arange(row0, row0 + rows, dtype=np.float64), n)
    K = len(sp)
    W = np.bincount(ids[hit], weights=d[hit], minlength=K)
    X = np.bincount(ids[hit], weights=(d * xs)[hit], minlength=K)
    Y = np.bincount(ids[hit], weights=(d * ys)[hit], minlength=K)
    lx = lab[..., 0].astype(np.float64).ravel(); ly = lab[..., 1].astype(np.float64).ravel()
    E = (d * ((lx - xs) ** 2 + (ly - ys) ** 2))[lab[..., 0].ravel() != I.MARK].sum()
    return W, X, Y, E


def _close(a, ref, tol=1e-11):
    return np.abs(a - ref).max() <= tol * max(np.abs(ref).max(), 1e-300)


@pytest.mark.parametrize("kind,n,k", [("uniform", 256, 400), ("c3", 512, 3000), ("c3", 1024, 20000), ("uniform", 2048, 1500),
                                      ("c3", 2048, 10000)])
def test_dense_centroid_matches_run_based_kernel_and_direct_sums(kind, n, k):
    import surface_remesher_b200 as S
    dens, mask, seeds = _case(kind, n, k)
    with S.Context(n) as c:
        c.set_density(dens); c.set_mask(mask); c.set_site_map(seeds)
        sites = c.get_sites()
        K = len(sites)
        c.label()
        lab = c.get_labels()
        c.accumulate(True)
        a_runs = _acc(c, zero=True)
        c.accumulate_dense(None, True)
        a_dense = _acc(c, zero=True)
    W, X, Y, E = _direct_sums(lab, dens, sites)
    for col, ref in ((0, W), (1, X), (2, Y)):
        assert _close(a_dense[col:4 * K:4], ref), col
        assert _close(a_dense[col:4 * K:4], a_runs[col:4 * K:4]), col
    e_idx = len(a_dense) - 4   # the energy slot follows the 4 * capacity per-site doubles (srm_acc_buffer)
    assert abs(a_dense[e_idx] - E) / E < 1e-11 and abs(a_runs[e_idx] - E) / E < 1e-9


@pytest.mark.parametrize("kind,n,k,omega", [("uniform", 256, 400, 2.0), ("c3", 512, 3000, 1.37), ("c3", 1024, 20000, 2.0),
                                            ("c3", 2048, 10000, 2.0), ("uniform", 4096, 20000, 2.0)])
def test_step_through_the_dense_centroid_pass_is_bit_exact(kind, n, k, omega):
    """label -> (expand) -> dense centroid pass -> update against the oracle's step: the rounded pixels of the update law
    are the comparison (identical site sets), as for the run-based accumulation."""
    import surface_remesher_b200 as S
    dens, mask, seeds = _case(kind, n, k)
    elab, eout, e = O.lloyd_step(seeds, dens, mask, omega)
    with S.Context(n) as c:
        c.set_density(dens); c.set_mask(mask); c.set_site_map(seeds)
        c.set_omega(omega)
        c.label()
        c.accumulate_dense(None, True)
        c.update()
        p = np.asarray(c.get_sites(), np.int32)
        st = c.state()
    got = set(zip((p & 0xFFFF).tolist(), (p >> 16).tolist()))
    assert got == I.site_set(eout)
    assert st["iterations"] == 1 and st["num_sites"] == len(got)
    assert st["energy"] == np.float32(e) or abs(st["energy"] - e) / e < 1e-6


def test_dense_centroid_over_an_external_approximate_label_map():
    """A device label map that is NOT the context's exact labelling: the jump-flooding result (a few pixels differ, a
    label may form non-adjacent runs in a row) and a band context.  Sums must equal numpy's over that very map."""
    import torch
    import surface_remesher_b200 as S
    n, k = 1024, 6000
    dens, mask, seeds = _case("c3", n, k)
    steps = [1] + [n >> s for s in range(1, 11)]
    with S.Context(n) as c:
        c.set_density(dens); c.set_mask(mask); c.set_site_map(seeds)
        sites = c.get_sites()
        K = len(sites)
        lab_t = torch.empty((n, n, 2), dtype=torch.int16, device="cuda")
        c.label_jfa(steps, lab_t)
        c.accumulate_dense(lab_t, True)
        a = _acc(c, zero=True)
        lab = lab_t.cpu().numpy()
    W, X, Y, E = _direct_sums(lab, dens, sites)
    for col, ref in ((0, W), (1, X), (2, Y)):
        assert _close(a[col:4 * K:4], ref), col
    # band context: rows 256..512 of the same map
    r0, r1 = 256, 512
    with S.Context(n, r0, r1) as c:
        c.set_density(dens); c.set_mask(mask); c.set_sites(sites)
        band = lab_t[r0:r1].contiguous()
        c.accumulate_dense(band, False)
        a = _acc(c, zero=True)
    W, X, Y, _ = _direct_sums(lab[r0:r1], dens[r0:r1], sites, r0)
    for col, ref in ((0, W), (1, X), (2, Y)):
        assert _close(a[col:4 * K:4], ref), col


def test_dense_centroid_argument_errors():
    import torch
    import surface_remesher_b200 as S
    n = 256
    dens, mask, seeds = _case("uniform", n, 100)
    with S.Context(n) as c:
        c.set_density(dens); c.set_mask(None); c.set_site_map(seeds)
        with pytest.raises(S.SrmError):
            c.accumulate_dense(None, False)          # no labelling yet and no map given
        lab = torch.empty((n * n * 2 + 2,), dtype=torch.int16, device="cuda")
        with pytest.raises(S.SrmError):
            c.accumulate_dense(lab[2:], False)       # 4-byte aligned only
        with pytest.raises(TypeError):
            c.accumulate_dense(np.zeros((n, n, 2), np.int16), False)   # host memory
