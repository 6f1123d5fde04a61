"""Seeded synthetic inputs shared by tests, bench.py and the golden generator."""
import numpy as np

MARK = -32768


def splitmix64(seed):
    x = np.uint64(seed)
    mask = np.uint64(0xFFFFFFFFFFFFFFFF)
    while True:
        x = (x + np.uint64(0x9E3779B97F4A7C15)) & mask
        z = x
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & mask
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & mask
        z = z ^ (z >> np.uint64(31))
        yield float(z >> np.uint64(11)) / float(1 << 53)


def density_uniform(n):
    return np.ones((n, n), np.float32)


def density_c3(n, seed=42, rows=None):
    """Curvature-like anisotropic density of BASELINE.md §5 (C3): clamp(0.01*A + kappa), kappa = 16 anisotropic
    Gaussian ridges from splitmix64(seed); zero outside the star domain r < 0.45(1+0.15 cos 5phi).
    rows=(r0,r1) generates a band only."""
    with np.errstate(over="ignore"):
        g = splitmix64(seed)
        ridges = []
        for _ in range(16):
            cx, cy = 0.1 + 0.8 * next(g), 0.1 + 0.8 * next(g)
            th = 2 * np.pi * next(g)
            s1, s2 = 0.15 + 0.25 * next(g), 0.005 + 0.015 * next(g)
            a = 5 + 45 * next(g)
            ridges.append((cx, cy, th, s1, s2, a))
    r0, r1 = rows if rows else (0, n)
    ys = (np.arange(r0, r1, dtype=np.float64) / (n - 1))[:, None]
    xs = (np.arange(n, dtype=np.float64) / (n - 1))[None, :]
    kappa = np.zeros((r1 - r0, n))
    for cx, cy, th, s1, s2, a in ridges:
        dx, dy = xs - cx, ys - cy
        e1 = dx * np.cos(th) + dy * np.sin(th)
        e2 = -dx * np.sin(th) + dy * np.cos(th)
        kappa += a * np.exp(-0.5 * ((e1 / s1) ** 2 + (e2 / s2) ** 2))
    w = 0.01 * 1.0 + 1.0 * kappa
    mi, mx = 1e-3, 50.0  # weighting.h:210-230 clamps to the [min,max] of weights in (eps, inf)
    w = np.clip(w, mi, mx)
    dx, dy = xs - 0.5, ys - 0.5
    r = np.sqrt(dx * dx + dy * dy)
    phi = np.arctan2(dy, dx)
    inside = r < 0.45 * (1 + 0.15 * np.cos(5 * phi))
    return np.where(inside, w, 0.0).astype(np.float32)


def c3_torch(n, device, seed=42, every=16):
    """density_c3 + mask_c3 evaluated with torch on `device` (same formulas, float64 math, band by band).
    Used by bench.py / tools for large grids; values agree with the numpy generator to the last ulp of exp()."""
    import torch
    with np.errstate(over="ignore"):
        g = splitmix64(seed)
        ridges = []
        for _ in range(16):
            cx, cy = 0.1 + 0.8 * next(g), 0.1 + 0.8 * next(g)
            th = 2 * np.pi * next(g)
            s1, s2 = 0.15 + 0.25 * next(g), 0.005 + 0.015 * next(g)
            a = 5 + 45 * next(g)
            ridges.append((cx, cy, th, s1, s2, a))
    dens = torch.empty((n, n), dtype=torch.float32, device=device)
    xs = (torch.arange(n, dtype=torch.float64, device=device) / (n - 1))[None, :]
    step = max(64, min(n, (1 << 24) // n))
    for r0 in range(0, n, step):
        r1 = min(n, r0 + step)
        ys = (torch.arange(r0, r1, dtype=torch.float64, device=device) / (n - 1))[:, None]
        kappa = torch.zeros((r1 - r0, n), dtype=torch.float64, device=device)
        for cx, cy, th, s1, s2, a in ridges:
            dx, dy = xs - cx, ys - cy
            e1 = dx * np.cos(th) + dy * np.sin(th)
            e2 = -dx * np.sin(th) + dy * np.cos(th)
            kappa += a * torch.exp(-0.5 * ((e1 / s1) ** 2 + (e2 / s2) ** 2))
        w = torch.clamp(0.01 + kappa, 1e-3, 50.0)
        dx, dy = xs - 0.5, ys - 0.5
        rr = torch.sqrt(dx * dx + dy * dy)
        phi = torch.atan2(dy, dx)
        inside = rr < 0.45 * (1 + 0.15 * torch.cos(5 * phi))
        dens[r0:r1] = torch.where(inside, w, torch.zeros_like(w)).to(torch.float32)
    ins = dens != 0
    edge = ins.clone()
    edge[1:-1, 1:-1] = ins[1:-1, 1:-1] & ~(ins[:-2, 1:-1] & ins[2:, 1:-1] & ins[1:-1, :-2] & ins[1:-1, 2:])
    idx = torch.nonzero(edge.reshape(-1), as_tuple=False).reshape(-1)[::every]   # scan order, like np.nonzero
    mask = torch.zeros(n * n, dtype=torch.uint8, device=device)
    mask[idx] = 1
    return dens, mask.reshape(n, n)


def mask_c3(density, every=16):
    """Boundary samples of the non-zero domain, one every `every` boundary pixels in scan order
    (mimics generateMask on the mesh border vertices)."""
    ins = density != 0
    edge = ins.copy()
    edge[1:-1, 1:-1] = ins[1:-1, 1:-1] & ~(ins[:-2, 1:-1] & ins[2:, 1:-1] & ins[1:-1, :-2] & ins[1:-1, 2:])
    ys, xs = np.nonzero(edge)
    m = np.zeros(density.shape, np.uint8)
    m[ys[::every], xs[::every]] = 1
    return m


def lattice_sites(n, pitch=8, off=4):
    """Regular lattice: thousands of exact ties, the case that distinguishes tie-break rules."""
    v = np.full((n, n, 2), MARK, np.int16)
    for y in range(off, n, pitch):
        for x in range(off, n, pitch):
            v[y, x] = (x, y)
    return v


def random_sites(n, k, seed):
    rng = np.random.default_rng(seed)
    v = np.full((n, n, 2), MARK, np.int16)
    idx = rng.choice(n * n, size=k, replace=False)
    ys, xs = np.divmod(idx, n)
    v[ys, xs, 0] = xs
    v[ys, xs, 1] = ys
    return v


def site_set(site_map):
    n = site_map.shape[0]
    ok = (site_map[..., 0] == np.arange(n)[None, :]) & (site_map[..., 1] == np.arange(n)[:, None])
    ys, xs = np.nonzero(ok)
    return set(zip(xs.tolist(), ys.tolist()))


def random_mesh(nv_side, seed, jitter=0.3):
    """Jittered grid triangulation of the unit square: (points[P,2], weights[P], triangles[T,3])."""
    rng = np.random.default_rng(seed)
    g = np.linspace(0, 1, nv_side)
    X, Y = np.meshgrid(g, g)
    h = 1.0 / (nv_side - 1)
    J = (rng.random((nv_side, nv_side, 2)) - 0.5) * jitter * h
    J[0, :, 1] = 0; J[-1, :, 1] = 0; J[:, 0, 0] = 0; J[:, -1, 0] = 0
    pts = np.stack([X + J[..., 0], Y + J[..., 1]], -1).reshape(-1, 2)
    wt = 0.5 + rng.random(len(pts)) * 4
    tri = []
    for j in range(nv_side - 1):
        for i in range(nv_side - 1):
            a = j * nv_side + i; b = a + 1; c = a + nv_side; d = c + 1
            if (i + j) % 2: tri += [(a, b, d), (a, d, c)]
            else: tri += [(a, b, c), (b, d, c)]
    return np.ascontiguousarray(pts), np.ascontiguousarray(wt), np.asarray(tri, np.int32)
