"""world_size=2 gloo test of the row-band host logic (surface-remesher_b200/sharded.py) on CPU.
The band engine here is backed by the CPU oracle (tests only); the product engine is CudaBandEngine."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import _inputs as I
import _oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class OracleBandEngine:
    """Same contract as CudaBandEngine: per-band sums in acc (4 doubles per site + energy), replicated update."""

    def __init__(self, n, row0, row1):
        self.n, self.row0, self.row1 = n, row0, row1

    def set_inputs(self, density, mask, site_map):
        self.d, self.m = density, mask
        xy = O.sites_of(site_map)
        self.sites = [tuple(p) for p in xy.tolist()]
        self.cap = len(self.sites)
        self.acc = torch.zeros(4 * self.cap + 4, dtype=torch.float64)
        self.omega, self.lastE, self.E, self.it = np.float32(2.0), np.float32(1e18), np.float32(0), 0

    def _map(self):
        v = np.full((self.n, self.n, 2), I.MARK, np.int16)
        for x, y in self.sites:
            v[y, x] = (x, y)
        return v

    def label(self):
        self.lab = O.label_exact(self._map())

    def accumulate(self, want_energy):
        ids = {s: k for k, s in enumerate(self.sites)}
        a = self.acc.numpy()
        a[:] = 0
        lab = self.lab[self.row0:self.row1]
        for yy in range(self.row1 - self.row0):
            y = self.row0 + yy
            for x in range(self.n):
                k = ids[(int(lab[yy, x, 0]), int(lab[yy, x, 1]))]
                d = float(self.d[y, x])
                a[4 * k] += d; a[4 * k + 1] += x * d; a[4 * k + 2] += y * d
                if want_energy:
                    a[4 * self.cap] += d * ((int(lab[yy, x, 0]) - x) ** 2 + (int(lab[yy, x, 1]) - y) ** 2)
        self.want_energy = want_energy

    def acc_tensor(self):
        return self.acc

    def update(self):
        n = self.n
        a = self.acc.numpy()
        W = np.zeros((n, n)); X = np.zeros((n, n)); Y = np.zeros((n, n))
        for k, (x, y) in enumerate(self.sites):
            W[y, x], X[y, x], Y[y, x] = a[4 * k], a[4 * k + 1], a[4 * k + 2]
        site_lab = self._map()
        out = np.empty_like(site_lab)
        O.lib().orc_update_sites(site_lab.ctypes.data, W.ctypes.data, X.ctypes.data, Y.ctypes.data, self.d.ctypes.data,
                                 None if self.m is None else self.m.ctypes.data, n, float(self.omega), out.ctypes.data)
        if self.want_energy:
            self.E = np.float32(a[4 * self.cap] / (n * n))
        new = O.sites_of(out)
        self.sites = [tuple(p) for p in new.tolist()]
        self.it += 1
        if self.it % 10 == 0:
            diff = np.float32(self.lastE - self.E)
            self.omega = np.float32(min(2.0, 1.0 + float(diff)))
            self.lastE = self.E

    def labels(self):
        return self.lab[self.row0:self.row1]


def _worker(rank, world, port, n, k, iters, q, bands=None):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from surface_remesher_b200.sharded import ShardedLloyd
    from surface_remesher_b200 import row_bands
    dens = I.density_c3(n)
    mask = I.mask_c3(dens)
    seeds, _, _ = O.seed(dens, mask, k)
    r0, r1 = (bands or row_bands(n, world))[rank]
    eng = OracleBandEngine(n, r0, r1)
    eng.set_inputs(dens, mask, seeds)
    sl = ShardedLloyd(n, rank, world, eng, dist, bands)
    sl.run(iters)
    lab = sl.final_labels()
    q.put((rank, r0, r1, lab, sorted(eng.sites)))
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


@pytest.mark.parametrize("bands", [None, [(0, 64), (64, 128)], [(0, 64), (64, 128), (128, 256)]],
                         ids=["equal-2", "explicit-2", "unequal-3"])
def test_two_band_lloyd_equals_single_process(bands):
    """world_size 2 (equal bands, and the same partition passed explicitly) and world_size 3 with bands of unequal
    height (64 / 64 / 128 rows), as row_bands_balanced produces them."""
    world = len(bands) if bands else 2
    n, k, iters = (256, 80, 6) if world == 3 else (128, 60, 12)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, k, iters, q, bands)) for r in range(world)]
    for p in procs: p.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for p in procs: p.join(timeout=60)
    assert all(p.exitcode == 0 for p in procs)
    dens = I.density_c3(n); mask = I.mask_c3(dens)
    seeds, _, _ = O.seed(dens, mask, k)
    exp, it, _, _ = O.gcvt(seeds, dens, mask, iters, stop_rule=0)
    assert it == iters
    res.sort()
    assert all(r[4] == res[0][4] for r in res), "replicated site lists diverged between ranks"
    full = np.concatenate([r[3] for r in res], axis=0)
    assert (full != exp).sum() == 0


def test_row_bands_partition():
    from surface_remesher_b200 import row_bands
    for n, w in [(256, 1), (256, 2), (256, 4), (8192, 8), (32768, 8)]:
        b = row_bands(n, w)
        assert b[0][0] == 0 and b[-1][1] == n
        assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
        assert all((r1 - r0) % 64 == 0 for r0, r1 in b)
    with pytest.raises(ValueError):
        row_bands(256, 8)


def test_row_bands_balanced_partition():
    """Equal-work row bands: a partition (contiguous, multiples of the unit, no empty band), better balanced than
    equal heights on a density with empty regions, identical for identical inputs, and equal to row_bands when
    there is nothing to balance."""
    from surface_remesher_b200 import row_bands, row_bands_balanced
    import _inputs as I
    import _oracle as O
    n = 2048
    dens = I.density_c3(n)
    seeds, _, _ = O.seed(dens, I.mask_c3(dens), 6000)
    ys = np.nonzero(seeds[..., 0] != I.MARK)[0]
    for w in (2, 4, 8):
        b = row_bands_balanced(n, w, ys, unit=64, fixed=0.3)
        assert b == row_bands_balanced(n, w, ys.copy(), unit=64, fixed=0.3)
        assert b[0][0] == 0 and b[-1][1] == n and all(b[i][1] == b[i + 1][0] for i in range(w - 1))
        assert all((r1 - r0) % 64 == 0 and r1 > r0 for r0, r1 in b)
        load = lambda bands: max(((ys >= a) & (ys < c)).sum() for a, c in bands)
        assert load(b) <= load(row_bands(n, w))
    b8 = row_bands_balanced(n, 8, ys, unit=64, fixed=0.3)
    assert max(((ys >= a) & (ys < c)).sum() for a, c in b8) < 0.8 * max(((ys >= a) & (ys < c)).sum() for a, c in row_bands(n, 8))
    assert row_bands_balanced(1024, 1, ys) == [(0, 1024)]
    assert row_bands_balanced(1024, 4, np.arange(1024)) == row_bands(1024, 4)      # uniform: equal heights
    assert all((r1 - r0) % 256 == 0 for r0, r1 in row_bands_balanced(32768, 8, np.random.default_rng(0).integers(8000, 24000, 100000)))
