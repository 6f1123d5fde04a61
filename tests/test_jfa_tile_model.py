"""CPU model of the fused shared-memory tile kernel of the jump-flooding family (csrc/srm_jfa.cu:k_jfa_tile).

The kernel's claim: a run of passes with steps s_0..s_{c-1}, sum H <= 15, computed per 64 x 64 tile on a staged
(64 + 2H)^2 window — pass p on the tile widened by the steps still to come, pixels outside the grid empty in every
pass — is bit-identical to the same passes run one by one over the whole grid (oracle/srm_oracle.c:orc_label_jfa).
This model restates exactly that schedule in numpy (same 64-bit key: dist^2, then x, then y) and checks it against the
oracle, including the grouping rule of the launcher (srm_launch_jfa).  The CUDA kernel itself is compared with the same
oracle in tests/test_gpu_label.py."""
import numpy as np
import pytest

import _inputs as I
import _oracle as O

TILE, HALO, MAXFUSE = 64, 15, 4
EMPTY = np.uint64(0xFFFFFFFF80008000)


def group_steps(steps):
    """Launch plan of srm_launch_jfa (mode 1): ('tile', [steps...]) for runs with sum <= 15 (at most 4), else ('far', k)."""
    plan, s = [], 0
    while s < len(steps):
        run, tot = [], 0
        while s + len(run) < len(steps) and len(run) < MAXFUSE and tot + steps[s + len(run)] <= HALO:
            run.append(steps[s + len(run)]); tot += run[-1]
        if run:
            plan.append(("tile", run)); s += len(run)
        else:
            plan.append(("far", steps[s])); s += 1
    return plan


def keys_of(ord_labels, gx, gy):
    """64-bit key of every candidate for the pixel at (gx, gy): dist^2 << 32 | ordered label (x high, y low)."""
    x = (ord_labels >> np.uint32(16)).astype(np.int64)
    y = (ord_labels & np.uint32(0xFFFF)).astype(np.int64)
    d = (x - gx) ** 2 + (y - gy) ** 2
    d = np.where(ord_labels == np.uint32(0x80008000), np.int64(0xFFFFFFFF), d)
    return (d.astype(np.uint64) << np.uint64(32)) | ord_labels.astype(np.uint64)


def tile_run(ordmap, n, tx, ty, run):
    """One CTA of k_jfa_tile: returns the 64 x 64 block of ordered labels after the run."""
    H = sum(run)
    P = TILE + 2 * HALO
    gx0, gy0 = tx * TILE - HALO, ty * TILE - HALO
    ly, lx = np.mgrid[0:P, 0:P]
    gx, gy = gx0 + lx, gy0 + ly
    inside = (gx >= 0) & (gx < n) & (gy >= 0) & (gy < n)
    src = np.full((P, P), 0x80008000, np.uint32)
    lo, hi = HALO - H, HALO + TILE + H
    win = inside[lo:hi, lo:hi]
    src[lo:hi, lo:hi][win] = ordmap[gy[lo:hi, lo:hi][win], gx[lo:hi, lo:hi][win]]
    m = H
    for s in run:
        m -= s
        lo, hi = HALO - m, HALO + TILE + m
        best = np.full((hi - lo, hi - lo), EMPTY, np.uint64)
        for j in (-1, 0, 1):
            for i in (-1, 0, 1):
                cand = src[lo + j * s:hi + j * s, lo + i * s:hi + i * s]   # always inside the staged window
                best = np.minimum(best, keys_of(cand, gx[lo:hi, lo:hi], gy[lo:hi, lo:hi]))
        dst = np.full((P, P), 0x80008000, np.uint32)   # unwritten cells are never read by a later pass
        dst[lo:hi, lo:hi] = np.where(inside[lo:hi, lo:hi], (best & np.uint64(0xFFFFFFFF)).astype(np.uint32), np.uint32(0x80008000))
        src = dst
    return src[HALO:HALO + TILE, HALO:HALO + TILE]


def model_jfa(seeds, steps):
    n = seeds.shape[0]
    x = seeds[..., 0].astype(np.int64) & 0xFFFF
    y = seeds[..., 1].astype(np.int64) & 0xFFFF
    ordmap = ((x << 16) | y).astype(np.uint32)
    for kind, arg in group_steps(list(steps)):
        if kind == "far":
            k = arg
            gy, gx = np.mgrid[0:n, 0:n]
            pad = np.full((n + 2 * k, n + 2 * k), 0x80008000, np.uint32)
            pad[k:k + n, k:k + n] = ordmap
            best = np.full((n, n), EMPTY, np.uint64)
            for j in (-1, 0, 1):
                for i in (-1, 0, 1):
                    best = np.minimum(best, keys_of(pad[k + j * k:k + j * k + n, k + i * k:k + i * k + n], gx, gy))
            ordmap = (best & np.uint64(0xFFFFFFFF)).astype(np.uint32)
        else:
            out = np.empty_like(ordmap)
            for ty in range(n // TILE):
                for tx in range(n // TILE):
                    out[ty * TILE:(ty + 1) * TILE, tx * TILE:(tx + 1) * TILE] = tile_run(ordmap, n, tx, ty, arg)
            ordmap = out
    lab = np.empty((n, n, 2), np.int16)
    lab[..., 0] = (ordmap >> np.uint32(16)).astype(np.uint16).view(np.int16)
    lab[..., 1] = (ordmap & np.uint32(0xFFFF)).astype(np.uint16).view(np.int16)
    return lab


def test_launch_plan_groups_small_steps():
    assert group_steps([1, 128, 64, 32, 16, 8, 4, 2, 1]) == [("tile", [1]), ("far", 128), ("far", 64), ("far", 32),
                                                              ("far", 16), ("tile", [8, 4, 2, 1])]
    assert group_steps([3, 100, 6, 5, 2, 1, 1]) == [("tile", [3]), ("far", 100), ("tile", [6, 5, 2, 1]), ("tile", [1])]
    assert group_steps([8, 8]) == [("tile", [8]), ("tile", [8])]
    assert group_steps([2, 2, 2, 2, 2]) == [("tile", [2, 2, 2, 2]), ("tile", [2])]


@pytest.mark.parametrize("case", ["random", "lattice", "border", "empty"])
def test_fused_tile_schedule_equals_pass_by_pass_jfa(case):
    n = 256
    if case == "random":
        seeds = I.random_sites(n, 300, 5)
        steps = [1] + [n >> (i + 1) for i in range(8)]          # 1+JFA: tile [1], far 128..16, tile [8,4,2,1]
    elif case == "lattice":                                     # many exact ties: the (x, y) part of the key decides
        seeds = np.full((n, n, 2), I.MARK, np.int16)
        for yy in range(3, n, 10):
            for xx in range(5, n, 10):
                seeds[yy, xx] = (xx, yy)
        steps = [64, 16, 6, 5, 2, 1, 1]
    elif case == "border":                                      # sites only on the grid's edge and corners
        seeds = np.full((n, n, 2), I.MARK, np.int16)
        for q in range(0, n, 17):
            seeds[0, q] = (q, 0); seeds[n - 1, q] = (q, n - 1); seeds[q, 0] = (0, q); seeds[q, n - 1] = (n - 1, q)
        seeds[n - 1, n - 1] = (n - 1, n - 1)
        steps = [1, 128, 64, 32, 16, 8, 4, 2, 1]
    else:
        seeds = np.full((n, n, 2), I.MARK, np.int16)
        seeds[200, 13] = (13, 200)                              # one site: most pixels stay empty in the early passes
        steps = [4, 2, 1, 8, 4, 2, 1]
    got = model_jfa(seeds, steps)
    exp = O.label_jfa(seeds, steps)
    assert (got != exp).sum() == 0
