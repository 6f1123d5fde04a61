"""Batches of independent meshes (BASELINE.json configs[4]: 256 meshes of 2048^2 / 10k sites over 8 GPUs).

Meshes shard trivially: one libsrm context per mesh, each on its own CUDA stream, `per_gpu` of them on every GPU and no
data-path collective ("replicas").  A 2048^2 Lloyd step is launch / latency bound (5 kernels of 5-40 us), so (1) the
loop of every context is replayed as a CUDA graph of 10 iterations (srm_set_option "graph"), and (2) the loops of all
contexts of a GPU are enqueued back to back without a host synchronisation, so that the streams overlap on the device.
"""
import numpy as np

from .api import Context


class BatchLloyd:
    """Independent gCVT problems of one size on one GPU."""

    def __init__(self, n, device=0, graph=True):
        self.n, self.device, self.graph = int(n), int(device), bool(graph)
        self.ctx = []

    def add(self, density, mask, site_map):
        """Upload one problem (numpy arrays or torch tensors on `device`); returns its index in the batch."""
        c = Context(self.n, 0, self.n, self.device)
        c.set_option("graph", self.graph)
        c.set_density(density)
        c.set_mask(mask)
        c.set_site_map(site_map)
        self.ctx.append(c)
        return len(self.ctx) - 1

    def run(self, max_iter, stop_rule=True):
        """gCVT's loop on every problem (gcvt.cu:1110-1147, single level) + the final labelling; returns the stats."""
        for c in self.ctx:          # enqueue everything first: no call below waits for the device
            c.iterate(max(1, int(max_iter)), stop_rule)
        for c in self.ctx:
            c.label()               # (reads the iteration count back: the device may have stopped early)
        return [c.state() for c in self.ctx]

    def iterate(self, iters):
        for c in self.ctx:
            c.iterate(iters, False)

    def synchronize(self):
        for c in self.ctx:
            c.synchronize()

    def labels(self, i, out=None):
        return self.ctx[i].get_labels(out)

    def sites(self, i):
        return self.ctx[i].get_sites()

    def close(self):
        for c in self.ctx:
            c.close()
        self.ctx = []

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


def shard_meshes(num_meshes, world, rank):
    """Indices of the meshes rank `rank` of `world` owns (contiguous blocks, sizes differ by at most one)."""
    base, extra = divmod(int(num_meshes), int(world))
    lo = rank * base + min(rank, extra)
    return list(range(lo, lo + base + (1 if rank < extra else 0)))
