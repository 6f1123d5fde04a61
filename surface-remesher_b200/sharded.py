"""Row-band sharding of one Lloyd loop over the GPUs of a node.

One process per GPU.  Rank r owns rows [r*n/W, (r+1)*n/W).  Site list, site bitmap, density and mask
are replicated; the labelling of a band needs no halo exchange because every rank scans the replicated
column bitmap above and below its band (csrc/srm_label.cu:k_carry).  The only data-path collective is
one all-reduce (sum, fp64) per iteration over the per-site accumulators (W, X, Y, 0) plus the energy
scalar; the site update is then replicated on every rank, deterministically, so the site lists stay
identical without a broadcast.

The band engine is injected so that the host logic (partition, collective, schedule) is testable on
CPU with the gloo backend (tests/test_dist_gloo.py uses an oracle-backed engine); the product engine
is CudaBandEngine over libsrm.so.
"""
import numpy as np

from .api import Context, row_bands, scan_mask, scan_site_map


class _CudaArray:
    """__cuda_array_interface__ view of a raw device pointer (float64 vector by default)."""

    def __init__(self, ptr, count, typestr="<f8"):
        self.__cuda_array_interface__ = {"shape": (int(count),), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 3, "strides": None}


class CudaBandEngine:
    """libsrm.so context for one row band; all work is enqueued on torch's current CUDA stream."""

    def __init__(self, n, row0, row1, device):
        import torch
        self.torch = torch
        self.device = torch.device("cuda", device)
        torch.cuda.set_device(self.device)
        self.ctx = Context(n, row0, row1, device)
        self.ctx.set_stream(torch.cuda.current_stream(self.device).cuda_stream)
        self._acc = None

    def set_inputs(self, density, mask, site_map):
        self.ctx.set_density(density)
        self.ctx.set_mask(mask)
        self.ctx.set_site_map(site_map)
        ptr, cnt = self.ctx.acc_buffer()
        self._keep = _CudaArray(ptr, cnt)
        self._acc = self.torch.as_tensor(self._keep, device=self.device)

    def set_inputs_sharded(self, density, mask, site_map, dist, bands, rank):
        """Like set_inputs, but this rank uploads only ITS rows of the density (1/world of the bytes); the "density != 0"
        bitmap the replicated site update needs for the whole grid is completed by exchanging the bands' slices between
        the ranks (one broadcast per band over NCCL/NVLink: n*n/8 bytes in total)."""
        n = self.ctx.n
        r0, r1 = bands[rank]
        # the two sparse inputs: every rank scans ITS rows on the host (multi-threaded C scans, srm_host.cu) while its rows
        # of the density go up (both calls release the GIL); the lists of all ranks, concatenated in rank order, are the
        # row-major scan of the whole arrays (the order a single GPU sees, so the site ids agree)
        import threading
        scanned = {}

        def scan():
            try:
                scanned["sites"] = scan_site_map(np.asarray(site_map).reshape(n, n, 2)[r0:r1])
                scanned["mpx"] = scan_mask(mask, n, r0, r1) if mask is not None else np.zeros(0, np.int32)
            except BaseException as e:   # re-raised on the calling thread below
                scanned["error"] = e
        th = threading.Thread(target=scan)
        th.start()
        try:
            self.ctx.set_density_band(density[r0:r1])
            ptr, words = self.ctx.shared_bits(0)
            self._nzkeep = _CudaArray(ptr, words, "<i4")
            nz = self.torch.as_tensor(self._nzkeep, device=self.device)
            self.torch.cuda.current_stream(self.device).synchronize()
            for q, (a, b) in enumerate(bands):
                dist.broadcast(nz[a * n // 32: b * n // 32], src=q)
        finally:
            th.join()
        if "error" in scanned:
            raise scanned["error"]
        sites, mpx = scanned["sites"], scanned["mpx"]
        sites, mpx = self._allgather_lists(dist, [sites, mpx], len(bands))
        if mask is not None:
            self.ctx.set_mask_pixels(mpx)
        else:
            self.ctx.set_mask(None)
        self.ctx.set_sites(sites)
        ptr, cnt = self.ctx.acc_buffer()
        self._keep = _CudaArray(ptr, cnt)
        self._acc = self.torch.as_tensor(self._keep, device=self.device)

    def _allgather_lists(self, dist, lists, world):
        """Concatenation over the ranks (in rank order) of each of `lists` (int32 vectors of different lengths)."""
        torch = self.torch
        cnt = torch.tensor([len(a) for a in lists], dtype=torch.int64, device=self.device)
        allc = [torch.zeros_like(cnt) for _ in range(world)]
        dist.all_gather(allc, cnt)
        allc = torch.stack(allc).cpu().numpy()            # [world, len(lists)]
        out = []
        for j, a in enumerate(lists):
            m = int(allc[:, j].max())
            buf = torch.zeros(max(m, 1), dtype=torch.int32, device=self.device)
            if len(a):
                buf[:len(a)] = torch.from_numpy(np.ascontiguousarray(a)).to(self.device)
            parts = [torch.zeros_like(buf) for _ in range(world)]
            dist.all_gather(parts, buf)
            out.append(np.concatenate([parts[q][:int(allc[q, j])].cpu().numpy() for q in range(world)]).astype(np.int32))
        return out

    def set_sites(self, packed):
        self.ctx.set_sites(packed)
        ptr, cnt = self.ctx.acc_buffer()
        self._keep = _CudaArray(ptr, cnt)
        self._acc = self.torch.as_tensor(self._keep, device=self.device)

    def label(self):
        self.ctx.label()

    def accumulate(self, want_energy):
        self.ctx.accumulate(want_energy)

    def label_accumulate(self, want_energy):
        self.ctx.label_accumulate(want_energy)

    def acc_tensor(self):
        return self._acc

    def update(self):
        self.ctx.update()

    def sites(self):
        return self.ctx.get_sites()

    def labels(self, out=None):
        """Dense labels of this band; `out`: a host array to fill instead of a fresh one (a fresh 134 MB numpy array costs
        ~10 ms of first-touch page faults per call at 8192^2 on two ranks — callers that repeat the call reuse theirs)."""
        return self.ctx.get_labels(out)

    def state(self):
        return self.ctx.state()

    def close(self):
        self.ctx.close()


class ShardedLloyd:
    """The Lloyd loop of gCVT (gcvt.cu:1110-1147, single level) over `world` row bands."""

    def __init__(self, n, rank, world, engine, dist=None, bands=None):
        self.n, self.rank, self.world = n, rank, world
        self.bands = list(bands) if bands else row_bands(n, world)    # e.g. row_bands_balanced(...)
        self.row0, self.row1 = self.bands[rank]
        self.engine = engine
        self.dist = dist
        self.it = 0

    def set_inputs(self, density, mask, site_map):
        """Inputs of this rank's band.  With the CUDA engine and world > 1 every rank uploads only its own rows of the
        density and the ranks exchange the "density != 0" bitmap slices (CudaBandEngine.set_inputs_sharded); otherwise the
        engine takes the full arrays."""
        if self.world > 1 and hasattr(self.engine, "set_inputs_sharded"):
            bands = getattr(self, "bands", None) or row_bands(self.n, self.world)
            self.engine.set_inputs_sharded(density, mask, site_map, self.dist, bands, self.rank)
        else:
            self.engine.set_inputs(density, mask, site_map)
        self.it = 0

    def step(self):
        """label -> per-band accumulate -> all-reduce -> replicated update."""
        want_energy = (self.it % 10) == 0   # gcvt.cu:1116
        if hasattr(self.engine, "label_accumulate"):
            self.engine.label_accumulate(want_energy)   # fused band kernel
        else:
            self.engine.label()
            self.engine.accumulate(want_energy)
        if self.world > 1:
            self.dist.all_reduce(self.engine.acc_tensor())  # sum
        self.engine.update()
        self.it += 1

    def bind_native_collective(self, mode="p2p"):
        """world > 1 with the CUDA engine: hand the all-reduce to libsrm itself, so that run() is one C++ loop without
        per-iteration Python.  mode "p2p": fused all-reduce over peer memory (the update kernel pulls the partial sums
        from the peers over NVLink; CUDA IPC handles gathered through torch.distributed).  mode "nccl": an NCCL
        all-reduce issued by libsrm on its stream (communicator id broadcast through torch.distributed).
        Call after set_inputs()."""
        if self.world == 1 or not hasattr(self.engine, "ctx"):
            return False
        import torch
        dev = self.engine.device
        if mode == "p2p":
            mine = torch.tensor(list(self.engine.ctx.p2p_info()), dtype=torch.uint8, device=dev)
            allb = [torch.zeros(160, dtype=torch.uint8, device=dev) for _ in range(self.world)]
            self.dist.all_gather(allb, mine)
            self.engine.ctx.p2p_connect([bytes(t.cpu().tolist()) for t in allb], self.rank, self.world)
            self.dist.barrier()
            self.native = True
            return True
        if self.rank == 0:
            raw = self.engine.ctx.nccl_unique_id()
            t = torch.tensor(list(raw), dtype=torch.uint8, device=dev)
        else:
            t = torch.zeros(128, dtype=torch.uint8, device=dev)
        self.dist.broadcast(t, 0)
        self.engine.ctx.nccl_init(bytes(t.cpu().tolist()), self.rank, self.world)
        self.native = True
        return True

    def shutdown(self):
        """Collective teardown of the band contexts: close the peer mappings on every rank, meet, then free."""
        if getattr(self, "native", False) and self.world > 1 and hasattr(self.engine, "ctx"):
            self.engine.ctx.p2p_disconnect()
            self.dist.barrier()
        self.engine.close()

    def run(self, iters):
        if getattr(self, "native", False) or (self.world == 1 and hasattr(self.engine, "ctx")):
            self.engine.ctx.iterate(iters, stop_rule=False)   # fused band kernel (+ NCCL all-reduce) + update, in C++
            self.it += iters
            return
        for _ in range(iters):
            self.step()

    def final_labels(self, out=None):
        """Labels of this band for the current sites (gcvt.cu:1149); `out`: host array [rows, n, 2] int16 to fill."""
        self.engine.label()
        return self.engine.labels() if out is None else self.engine.labels(out)
