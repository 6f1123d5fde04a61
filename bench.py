#!/usr/bin/env python
"""bench.py — Lloyd iterations/s of the discrete-CVT hot path at BASELINE.json's headline config
(configs[2]: curvature-like anisotropic density, 8192^2 grid, 100k sites) on N GPUs of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--n 8192] [--sites 100000]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A step = one Lloyd iteration (exact Voronoi labelling -> per-site centroid sums -> site update), i.e. the body
of the loop at reference gcvt.cu:1112-1123.  One JSON line on stdout (rank 0).

  value     whole-job iterations/s with inputs resident in HBM, CUDA events around exactly K steps, max over ranks
  e2e       the same metric through the reference-facing entry point: gCVT(host buffers) with --steps iterations per
            call and PAGEABLE (malloc) buffers, as the reference's caller hands them (main.cpp:203,214-215); H2D of
            density/mask/seed map, the loop, the final labelling and the D2H of the label map are all inside the timed
            region (wall clock around the synchronised call).  e2e_pinned: the same call with pinned buffers.
  parity    sha1 of the sorted site list after 20 steps from the seeds; must be equal at every N and equal to
            oracle_sha1 (the OpenMP CPU port's list after the same 20 steps)
  roofline  the dominant kernel (row envelope) against the measured HBM peak: algorithmic bytes / event time
  cpu_baseline  the OpenMP CPU port of the same loop (oracle/, test infrastructure) on a bounded sample

--impl reference times the UNMODIFIED reference implementation of the path, which is CUDA: oracle/_ref/libsrm_ref.so
(gcvt.cu compiled for sm_100 from /root/reference by oracle/Makefile) on the same GPU, north_star baseline (a), like for
like with our arm: value = its device-resident loop body (inputs already on the GPU), e2e = its own entry point gCVT()
with pageable host buffers and the same --steps iterations per call.  If that library is missing or fails at this
size, the OpenMP CPU port is timed instead.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "lloyd_iterations_per_s"
UNIT = "it/s"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def make_inputs(n, k, pinned):
    """C3 inputs: density via the shared generator (numpy, band by band to bound memory),
    boundary mask, sites from the reference seeding algorithm (srm_seed, host code as in gcvt.h:76-104)."""
    import _inputs as I
    import surface_remesher_b200 as S
    import ctypes as C
    t0 = time.time()

    def alloc(shape, dtype):
        if pinned:
            import torch
            return torch.empty(shape, dtype=getattr(torch, np.dtype(dtype).name), pin_memory=True).numpy()
        return np.empty(shape, dtype)

    dens = alloc((n, n), np.float32)
    mask = alloc((n, n), np.uint8)
    use_gpu = False
    try:
        import torch
        use_gpu = torch.cuda.is_available()
    except Exception:
        pass
    if use_gpu:   # same formulas evaluated on the GPU (seconds instead of a minute at 32768^2)
        import torch
        d_t, m_t = I.c3_torch(n, torch.device("cuda", torch.cuda.current_device()))
        dens[:] = d_t.cpu().numpy(); mask[:] = m_t.cpu().numpy()
        del d_t, m_t
        torch.cuda.empty_cache()
    else:
        step = 512
        for r0 in range(0, n, step):
            dens[r0:r0 + step] = I.density_c3(n, rows=(r0, min(n, r0 + step)))
        mask[:] = I.mask_c3(dens)
    vor = alloc((n, n, 2), np.int16)
    S.api._ck(S.lib().srm_seed(vor.ctypes.data, dens.ctypes.data, mask.ctypes.data, int(k), n, None))
    log(f"[bench] inputs n={n} sites={k}+{int(mask.sum())} mask in {time.time() - t0:.1f}s")
    return dens, mask, vor


class ClockSampler:
    """SM clock and throttle reasons of one GPU while the timed region runs.  The timed region of the default run is a
    few milliseconds, shorter than one `nvidia-smi -lms` period, so the samples are taken in-process through NVML by a
    thread (~1 kHz; the main thread waits in cudaStreamSynchronize with the GIL released); `nvidia-smi -lms 100` runs
    beside it as a fallback.  mark_begin() / mark_end() bracket the timed region: "sm_mhz" is the median of the samples
    inside it (or, if it was too short for any, of the samples since the warm-up started)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    NVML_REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index
        self.nv, self.nv_stop, self.nv_thread, self.nv_max = [], False, None, None
        self.t_begin = self.t_end = None

    def _nvml_handle(self):
        import pynvml as N
        N.nvmlInit()
        try:   # CUDA_VISIBLE_DEVICES may renumber the devices: go through the UUID
            import torch
            uuid = str(torch.cuda.get_device_properties(self.index).uuid)
            uuid = uuid if uuid.startswith("GPU-") else "GPU-" + uuid
            return N, N.nvmlDeviceGetHandleByUUID(uuid.encode() if hasattr(uuid, "encode") else uuid)
        except Exception:
            return N, N.nvmlDeviceGetHandleByIndex(self.index)

    def _nvml_loop(self, N, h):
        while not self.nv_stop:
            try:
                sm = N.nvmlDeviceGetClockInfo(h, N.NVML_CLOCK_SM)
                try:
                    rs = N.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    rs = N.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.nv.append((time.perf_counter(), float(sm), int(rs)))
            except Exception:
                break
            time.sleep(0.0005)

    def start(self):
        try:
            N, h = self._nvml_handle()
            self.nv_max = float(N.nvmlDeviceGetMaxClockInfo(h, N.NVML_CLOCK_SM))
            self.nv_thread = threading.Thread(target=self._nvml_loop, args=(N, h), daemon=True)
            self.nv_thread.start()
        except Exception as e:
            log("[bench] NVML clock sampling unavailable:", e)
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception as e:  # nvidia-smi missing
            log("[bench] nvidia-smi clock sampling unavailable:", e)

    def mark_begin(self):
        self.t_begin = time.perf_counter()

    def mark_end(self):
        self.t_end = time.perf_counter()

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        self.nv_stop = True
        if self.nv_thread:
            self.nv_thread.join(timeout=1.0)
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [nm for j, nm in enumerate(names) if any(len(r) > 3 + j and r[3 + j].startswith("Active") for r in self.rows)]
        out = {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
               "reasons": reasons, "samples": len(sm), "source": "nvidia-smi -lms 100"}
        if self.nv:
            inside = [x for x in self.nv if self.t_begin is not None and self.t_end is not None and self.t_begin <= x[0] <= self.t_end]
            use = inside if inside else self.nv
            bits = 0
            for x in use:
                bits |= x[2]
            out = {"sm_mhz": float(np.median([x[1] for x in use])), "sm_max_mhz": self.nv_max or out["sm_max_mhz"],
                   "reasons": sorted(set(reasons) | {nm for b, nm in self.NVML_REASONS.items() if bits & b}),
                   "samples": len(use), "samples_in_timed_region": len(inside),
                   "source": "NVML, in-process thread" + ("" if inside else " (timed region shorter than one sample: warm-up + timed region)"),
                   "nvidia_smi_samples": len(sm)}
        return out


def sites_sha1(packed):
    """sha1 of the sorted packed site list (x | y << 16, int32): independent of list order and of merged-site holes."""
    import hashlib
    a = np.sort(np.asarray(packed, np.int64).astype(np.uint32))
    return hashlib.sha1(a.tobytes()).hexdigest()


PARITY_STEPS = 20


def workload_name(n, k, nmask):
    which = "configs[2]" if n == 8192 else "configs[3]" if n == 32768 else "generator of configs[2], other size"
    return f"C3 curvature-like anisotropic density {n}x{n}, {k} sites + {nmask} fixed boundary sites (BASELINE.json {which})"


def cpu_baseline(dens, mask, vor, iters=PARITY_STEPS):
    """OpenMP port (oracle/srm_oracle.c:orc_fast_step) on the host cores: bounded sample of the same workload.
    With iters == PARITY_STEPS the final site list is the oracle side of the parity hash (omega stays 2.0 for the first
    20 iterations of the control law, gcvt.cu:1125-1140)."""
    import _oracle as O
    f = O.FastLloyd(vor, dens, mask)
    K0 = f.K
    t0 = time.time()
    f.step(2.0)  # first step also pays page faults / OpenMP pool start: timed separately
    t1 = time.time()
    for _ in range(iters - 1):
        f.step(2.0)
    dt = time.time() - t1
    n = dens.shape[0]
    packed = (f.sx[:f.K].astype(np.int64) & 0xFFFF) | (f.sy[:f.K].astype(np.int64) << 16)
    return {"value": (iters - 1) / dt, "unit": UNIT, "cores": O.lib().orc_num_threads(), "kind": "port",
            "sample": f"{iters - 1} Lloyd iterations of the same {n}x{n} / {K0}-site workload after 1 warm-up iteration "
                      f"({t1 - t0:.2f} s)", "sites_sha1_after": {"steps": iters, "sha1": sites_sha1(packed), "num_sites": int(f.K)}}


# ----------------------------------------------------------------------------------------------- ours

def run_ours(args):
    import torch
    import torch.distributed as dist
    import surface_remesher_b200 as S
    from surface_remesher_b200.sharded import CudaBandEngine, ShardedLloyd

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n, k, K, W = args.n, args.sites, args.steps, args.warmup
    if args.only_c4:   # measurement helper: just the configs[3] leg (e.g. to compare band partitions)
        c4 = run_c4(args, world, rank, local, dist if world > 1 else None)
        if rank == 0:
            print(json.dumps({"c4": c4}), flush=True)
        if world > 1:
            dist.barrier(); dist.destroy_process_group()
        return
    dens, mask, vor = make_inputs(n, k, pinned=False)   # pageable numpy buffers, like the reference's caller
    nmask = int(mask.sum())
    # row bands of equal work (sites per block of rows), from the replicated seed map: identical on every rank
    bands = S.row_bands_balanced(n, world, np.nonzero(vor[..., 0] != -32768)[0]) if args.bands == "balanced" \
        else S.row_bands(n, world)
    r0, r1 = bands[rank]
    launches = S.lib().srm_launch_count

    eng = CudaBandEngine(n, r0, r1, local)
    sl = ShardedLloyd(n, rank, world, eng, dist if world > 1 else None, bands)
    sl.set_inputs(dens, mask, vor)   # N > 1: every rank uploads its own rows; the non-zero bitmap slices are exchanged
    if world > 1 and args.collective != "py":
        sl.bind_native_collective(args.collective)   # all-reduce inside libsrm's C++ loop (peer memory or NCCL)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- device-resident timing: W warm-up steps, then exactly K steps between events
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    sl.run(W)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = launches()
    sampler.mark_begin()
    e0.record()
    sl.run(K)
    e1.record()
    barrier()
    sampler.mark_end()
    gpu_launches = launches() - l0          # kernels of libsrm launched by this rank inside the timed region
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([ms, float(gpu_launches)], device="cuda", dtype=torch.float64)
        dist.all_reduce(t[:1], op=dist.ReduceOp.MAX)
        dist.all_reduce(t[1:], op=dist.ReduceOp.SUM)
        ms, gpu_launches = float(t[0].item()), int(t[1].item())
    st = eng.state()

    # ---- per-stage pass (same loop, events between the stages) for the roofline object
    stage = None
    if world == 1 or args.collective != "py":
        stage = eng.ctx.iterate_profiled(max(K, 10), stop_rule=False)   # every rank takes part in the all-reduce
        stage = {s_: v * K / max(K, 10) for s_, v in stage.items()}
        torch.cuda.synchronize()
        runs, ovf = eng.ctx.debug_counts()

    # ---- parity: 20 steps from the seeds, hash of the site list (replicated: every rank holds the same list)
    sl.set_inputs(dens, mask, vor)
    barrier()
    sl.run(PARITY_STEPS)
    par_sites = eng.sites()
    parity = {"steps": PARITY_STEPS, "sites_sha1": sites_sha1(par_sites), "num_sites": int(len(par_sites))}
    if world > 1:   # all ranks must agree
        h = torch.tensor(list(bytes.fromhex(parity["sites_sha1"])), device="cuda", dtype=torch.int32)
        hs = [torch.zeros_like(h) for _ in range(world)]
        dist.all_gather(hs, h)
        parity["equal_on_all_ranks"] = all(bool((x == h).all().item()) for x in hs)

    # ---- end to end through the reference-facing call: pageable host buffers, K iterations per call
    e2e_iters = args.e2e_iters if args.e2e_iters > 0 else K
    e2e_pinned = None
    if world == 1:
        eng.close()

        def call_gcvt(d_, m_, buf):
            buf[:] = vor
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            est = S.gCVT(buf, d_, m_, n, 1, e2e_iters)
            torch.cuda.synchronize()
            return est, time.perf_counter() - t0

        buf = np.empty((n, n, 2), np.int16)
        call_gcvt(dens, mask, buf)              # first call: context creation (the cached context is part of the design)
        best, est = None, None
        for rep in range(3):
            est, dt = call_gcvt(dens, mask, buf)
            best = dt if best is None else min(best, dt)
        its = max(est["iterations"], 1)
        e2e_val = est["iterations"] / best
        # bytes that cross PCIe per call: the density; the seed map and the mask are scanned on the host (srm_host.cu)
        # and go up as lists of K sites / constraint pixels
        h2d = (dens.nbytes + 4 * (k + nmask) + 4 * nmask) / its
        d2h = buf.nbytes / its
        if n <= 16384:   # same call with pinned buffers (extra key)
            pd = torch.empty((n, n), dtype=torch.float32, pin_memory=True).numpy(); pd[:] = dens
            pm = torch.empty((n, n), dtype=torch.uint8, pin_memory=True).numpy(); pm[:] = mask
            pb = torch.empty((n, n, 2), dtype=torch.int16, pin_memory=True).numpy()
            bp = min(call_gcvt(pd, pm, pb)[1] for _ in range(2))
            e2e_pinned = {"value": est["iterations"] / bp, "unit": UNIT, "ms_per_call": bp * 1e3}
            del pd, pm, pb
    else:
        # each rank: upload its inputs, run the loop, download its band of labels
        best = None
        lab = np.empty((sl.row1 - sl.row0, n, 2), np.int16)   # the caller's output buffer, reused like `buf` at N = 1
        for rep in range(3):
            barrier()
            t0 = time.perf_counter()
            sl.set_inputs(dens, mask, vor)
            sl.run(e2e_iters)
            sl.final_labels(lab)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            t = torch.tensor([dt], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            best = float(t.item()) if best is None else min(best, float(t.item()))
        e2e_val = e2e_iters / best
        # all ranks together: the density once (each rank its rows), the site / constraint lists on every rank
        h2d = (dens.nbytes + world * (4 * (k + nmask) + 4 * nmask)) / e2e_iters
        d2h = world * lab.nbytes / e2e_iters
        sl.shutdown()

    # ---- BASELINE.json configs[3] (32768^2, 10^6 sites), time-boxed: device-resident steps only
    c4 = None
    if args.c4_steps > 0 and n != 32768:
        try:
            c4 = run_c4(args, world, rank, local, dist if world > 1 else None)
        except Exception as e:   # pragma: no cover
            c4 = {"error": str(e)[:300]}

    if rank != 0:
        if world > 1:
            dist.barrier(); dist.destroy_process_group()
        return

    N = n * n
    peak, peak_src = measured_peaks()
    line = {
        "metric": METRIC, "value": K / (ms / 1e3), "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "int32 labels / f64 accumulators", "data": "synthetic",
        "config": {"workload": workload_name(n, k, nmask), "grid": n, "sites": st["num_sites"],
                   "parallelism": f"row bands x{world} ({args.bands}: {[b[1] - b[0] for b in bands]} rows), collective={args.collective}" if world > 1 else "single GPU",
                   "l2": f"per-step working set (density {4 * N / 1e6:.0f} MB streamed or fp64 prefix arrays read at run ends, "
                         f"bitmap + carries {3 * N / 8 / 1e6:.0f} MB) exceeds the 126 MB L2; no explicit flush",
                   "stop_rule": "off (fixed step count); energy every 10th step like the reference"},
        "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "iterations_per_call": e2e_iters, "ms_per_call": best * 1e3, "host_buffers": "pageable (malloc)",
                "api": "gCVT(host buffers)" if world == 1 else "ShardedLloyd(host buffers)"},
        "gpu_launches": gpu_launches,   # counted by libsrm (srm_launch_count) over the timed region, all ranks
        "parity": parity,
        "clocks": clocks,
    }
    if e2e_pinned:
        line["e2e_pinned"] = e2e_pinned
    if c4:
        line["c4"] = c4
    if stage is not None and world > 1:
        line["config"]["stages_ms_per_step_rank0"] = {s_: v / K for s_, v in stage.items()}
    if stage is not None:
        row_ms = stage["band_fused"] / K
        alg, alg_note = band_algorithmic_bytes(n, r1 - r0, runs)
        ach = alg / (row_ms / 1e3) / 1e9
        traffic = None   # dram__bytes_read.sum + dram__bytes_write.sum of k_band per launch, from the committed ncu capture
        tp = os.path.join(ROOT, "profiles", "r2_k_band_traffic.json")
        if n == 8192 and k == 100000 and world == 1 and os.path.exists(tp):
            tj = json.load(open(tp))
            traffic = 0.9 * tj["traffic_normal_step"] + 0.1 * tj["traffic_energy_step"]   # every 10th step computes the energy
        line["roofline"] = {"bound": "hbm", "kernel": "k_band (fused labelling + accumulation; " + alg_note +
                                                      ("; rank 0's band)" if world > 1 else ")"),
                            "runs_per_step": runs, "robust_path_rows": ovf,
                            "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic,
                            "peak_source": peak_src, "ms_per_launch": row_ms,
                            "stages_ms_per_step": {s_: v / K for s_, v in stage.items()},
                            "step_bytes_per_px_equiv_GBs": {"8B_per_px": 8.0 * N / (ms / K / 1e3) / 1e9}}
    if not args.no_cpu and n <= 8192:   # reported baseline + oracle side of the parity hash: rank 0
        try:
            cb = cpu_baseline(dens, mask, vor)
            parity["oracle_sha1"] = cb["sites_sha1_after"]["sha1"]
            parity["oracle_num_sites"] = cb["sites_sha1_after"]["num_sites"]
            parity["equal_to_oracle"] = parity["oracle_sha1"] == parity["sites_sha1"]
            if world == 1:
                line["cpu_baseline"] = {k_: v for k_, v in cb.items() if k_ != "sites_sha1_after"}
        except Exception as e:  # pragma: no cover
            line["cpu_baseline"] = {"error": str(e)}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier(); dist.destroy_process_group()


def run_c4(args, world, rank, local, dist):
    """BASELINE.json configs[3]: C3 generator at 32768^2 with 10^6 sites on `world` row bands.  A few warm-up steps, then
    args.c4_steps timed steps (CUDA events, max over ranks), and the sha1 of the replicated site list, which must be the
    same at every N."""
    import torch
    import surface_remesher_b200 as S
    from surface_remesher_b200.sharded import CudaBandEngine, ShardedLloyd
    n, k = 32768, 1000000
    t0 = time.time()
    dens, mask, vor = make_inputs(n, k, pinned=False)
    nmask = int(mask.sum())
    bands = S.row_bands_balanced(n, world, np.nonzero(vor[..., 0] != -32768)[0]) if args.bands == "balanced" else S.row_bands(n, world)

    def build(bands_):
        r0_, r1_ = bands_[rank]
        eng_ = CudaBandEngine(n, r0_, r1_, local)
        sl_ = ShardedLloyd(n, rank, world, eng_, dist, bands_)
        sl_.set_inputs(dens, mask, vor)
        if world > 1 and args.collective != "py":
            sl_.bind_native_collective(args.collective)
        return eng_, sl_

    eng, sl = build(bands)
    rebalanced_from = None
    if world > 1 and args.c4_bands == "auto" and args.collective != "py":
        # one round of measured rebalancing: band-kernel time of every rank on equal bands -> bands of equal cost
        sl.run(3)
        st0 = eng.ctx.iterate_profiled(6, stop_rule=False)
        tb = torch.tensor([st0["band_fused"] / 6], device="cuda", dtype=torch.float64)
        allt = [torch.zeros_like(tb) for _ in range(world)]
        dist.all_gather(allt, tb)
        times = [float(x.item()) for x in allt]
        new_bands = S.rebalance_bands(bands, times)
        if new_bands != bands:
            rebalanced_from = {"bands": [b[1] - b[0] for b in bands], "k_band_ms_per_rank": [round(x, 4) for x in times]}
            sl.shutdown()
            bands = new_bands
            eng, sl = build(bands)
        else:
            sl.set_inputs(dens, mask, vor)   # same partition: restart from the seeds (the hash counts iterations)
    del dens, vor
    t_setup = time.time() - t0

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    W, K = 5, args.c4_steps
    sl.run(W)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    sl.run(K)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    stage = eng.ctx.iterate_profiled(10, stop_rule=False)
    torch.cuda.synchronize()
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        sb = torch.tensor([stage["band_fused"] / 10], device="cuda", dtype=torch.float64)
        allb = [torch.zeros_like(sb) for _ in range(world)]
        dist.all_gather(allb, sb)
        band_ms = [round(float(x.item()), 4) for x in allb]
    else:
        band_ms = [round(stage["band_fused"] / 10, 4)]
    sites = eng.sites()
    st = eng.state()
    out = {"workload": workload_name(n, k, nmask), "grid": n, "sites": st["num_sites"], "n_gpus": world, "steps": K, "warmup": W,
           "value": K / (ms / 1e3), "unit": UNIT, "ms_per_step": ms / K, "bands": [b[1] - b[0] for b in bands],
           "k_band_ms_per_rank": band_ms, "stages_ms_per_step_rank0": {s_: round(v / 10, 4) for s_, v in stage.items()},
           "sites_sha1_after": {"steps": W + K + 10, "sha1": sites_sha1(sites)},
           "setup_s": round(t_setup, 1)}
    if rebalanced_from:
        out["rebalanced_from"] = rebalanced_from
    sl.shutdown()
    del mask
    return out


def band_algorithmic_bytes(n, rows, runs):
    """Algorithmic bytes of one k_band launch: per 8-row band and column 8 B (bitmap word + up/dn
    carries) = rows * n; per run the 16 B fp64 prefix pair at its end + 24 B of accumulator update + 4 B site id."""
    return 1.0 * rows * n + 44.0 * runs, "rows*n B + 44 B/run"


# ------------------------------------------------------------------------------------------ reference

def _ref_child(n, k, steps, inputs):
    """Runs in a subprocess: the reference's gCVT on the same inputs.  e2e = wall clock around its own entry point
    with pageable host buffers (alloc + H2D + loop + final labelling + D2H + free: what its caller pays);
    device-resident = CUDA events around its loop body only.
    The inputs come from a file written by the parent: this process must not touch torch.cuda — the reference
    writes 4 MB past its pbaMargin allocation, which only goes unnoticed while its own
    cudaMalloc blocks are the neighbours."""
    import _ref as R
    z = np.load(inputs)
    dens, mask, vor = z["dens"], z["mask"], z["vor"]
    R.gcvt(vor, dens, mask, 2)                      # warm-up (context, module load)
    best, it = None, 0
    for _ in range(2):
        t0 = time.perf_counter()
        out, it = R.gcvt(vor, dens, mask, steps)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    loop_ms = R.loop_timed(vor, dens, mask, steps)   # device-resident loop body only
    print(json.dumps({"it": it, "ms": best * 1e3, "loop_ms": loop_ms, "mask": int(mask.sum())}), flush=True)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n, k, K, W = args.n, args.sites, args.steps, args.warmup
    import _ref as R
    dens, mask, vor = make_inputs(n, k, pinned=False)
    nmask = int(mask.sum())
    line = {"impl": "reference", "metric": METRIC, "unit": UNIT, "n_gpus": args.gpus, "steps": K, "warmup": W,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "data": "synthetic",
            "config": {"workload": workload_name(n, k, nmask), "grid": n}}
    res, crashes = None, []
    inputs = None
    if R.available():
        import tempfile
        inputs = os.path.join(tempfile.gettempdir(), f"srm_ref_inputs_{os.getpid()}.npz")
        np.savez(inputs, dens=dens, mask=mask, vor=vor)
        # The reference CUDA is fragile at this size: on the B200 it dies with "illegal memory
        # access" (gpuErrchk at gcvt.cu:1150) for long runs.  Try the requested step count first, then shorter calls.
        for steps in [s_ for s_ in dict.fromkeys([K, 200, 100, 30, 20, 10]) if s_ <= K]:
            try:
                p = subprocess.run([sys.executable, os.path.abspath(__file__), "--_ref_child", "--n", str(n), "--sites", str(k),
                                    "--steps", str(steps), "--_inputs", inputs], capture_output=True, text=True, timeout=1500)
                log(p.stderr[-1500:])
                if p.returncode == 0 and p.stdout.strip():
                    res = json.loads(p.stdout.strip().splitlines()[-1])
                    res["steps"] = steps
                    break
                msg = [l for l in p.stderr.splitlines() if "GPUassert" in l or "rror" in l]
                crashes.append({"steps": steps, "rc": p.returncode, "msg": (msg[-1] if msg else "")[:160]})
            except Exception as e:
                crashes.append({"steps": steps, "msg": str(e)[:160]})
    if crashes:
        line["reference_crashes"] = crashes
    if inputs and os.path.exists(inputs):
        os.remove(inputs)
    if res and res["it"] > 0:
        v_loop = res["steps"] / (res["loop_ms"] / 1e3)      # device-resident loop body: like for like with our `value`
        v_call = res["it"] / (res["ms"] / 1e3)               # whole call with pageable host buffers: like our `e2e`
        line["steps"] = res["it"]
        line["config"]["steps_run"] = res["it"]
        if res["it"] != K:
            line["config"]["steps_note"] = f"the reference crashed at {K} iterations per call; {res['it']} were run"
        line.update({"value": v_loop, "ms_per_step": res["loop_ms"] / res["steps"], "dtype": "short2 labels / f32 sums",
                     "cpu_baseline": {"value": v_loop, "unit": UNIT, "cores": 0, "kind": "reference",
                                      "sample": f"reference CUDA loop body (oracle/_ref/libsrm_ref.so, built from the unmodified "
                                                f"gcvt.cu for sm_100) on the same B200, {res['steps']} iterations, inputs resident; "
                                                "the reference has no CPU implementation of this path"},
                     "e2e": {"value": v_call, "unit": UNIT,
                             "h2d_bytes_per_step": (dens.nbytes + mask.nbytes + vor.nbytes) / res["it"],
                             "d2h_bytes_per_step": vor.nbytes / res["it"], "iterations_per_call": res["it"],
                             "ms_per_call": res["ms"], "host_buffers": "pageable (malloc)",
                             "api": "reference gCVT(host buffers): alloc + H2D + loop + final labelling + D2H + free"}})
    else:
        cb = cpu_baseline(dens, mask, vor, iters=max(2, min(K, args.cpu_iters)))
        cb.pop("sites_sha1_after", None)
        line.update({"value": cb["value"], "ms_per_step": 1e3 / cb["value"], "dtype": "int32 labels / f64 sums",
                     "cpu_baseline": cb, "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0,
                                                 "d2h_bytes_per_step": 0},
                     "note": "reference CUDA library unavailable or failed at this size: OpenMP CPU port timed instead"})
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", "--grid", dest="n", type=int, default=8192,
                    help="grid side (under torch.distributed.run spell it --grid: the launcher's own parser trips over --n)")
    ap.add_argument("--sites", type=int, default=100000)
    ap.add_argument("--e2e-iters", dest="e2e_iters", type=int, default=0,
                    help="iterations per gCVT call of the e2e leg (default 0 = --steps, the same as the reference arm)")
    ap.add_argument("--cpu-iters", dest="cpu_iters", type=int, default=3)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--c4-steps", dest="c4_steps", type=int, default=30,
                    help="timed steps of the BASELINE configs[3] leg (32768^2, 10^6 sites) appended to the line as `c4`; 0 = skip")
    ap.add_argument("--only-c4", dest="only_c4", action="store_true", help="run only the configs[3] leg")
    ap.add_argument("--c4-bands", dest="c4_bands", default="auto", choices=["auto", "fixed"],
                    help="configs[3] leg at N > 1: auto = one round of measured rebalancing of the row bands (default); "
                         "fixed = the partition --bands gives")
    ap.add_argument("--collective", default="p2p", choices=["p2p", "nccl", "py"],
                    help="N>1: p2p = fused all-reduce over peer memory inside the update kernel (default); nccl = NCCL "
                         "all-reduce issued by libsrm; py = torch.distributed all-reduce per step from Python")
    ap.add_argument("--bands", default="equal", choices=["balanced", "equal"],
                    help="N>1: row bands of equal height (default) or of equal estimated work (row_bands_balanced)")
    ap.add_argument("--_ref_child", action="store_true")
    ap.add_argument("--_inputs", default=None)
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args._ref_child:
        _ref_child(args.n, args.sites, args.steps, args._inputs)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
