#!/usr/bin/env python
"""bench.py — Lloyd iterations/s of the discrete-CVT hot path at BASELINE.json's headline config
(configs[2]: curvature-like anisotropic density, 8192^2 grid, 100k sites) on N GPUs of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--n 8192] [--sites 100000]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A step = one Lloyd iteration (exact Voronoi labelling -> per-site centroid sums -> site update), i.e. the body
of the loop at reference gcvt.cu:1112-1123.  One JSON line on stdout (rank 0).

  value     whole-job iterations/s with inputs resident in HBM, CUDA events around exactly K steps, max over ranks
  e2e       the same metric through the reference-facing entry point (gCVT with HOST buffers in pinned memory:
            H2D of density/mask/seed map, the loop, the final labelling, D2H of the label map all inside the timed
            region)
  roofline  the dominant kernel (row envelope) against the measured HBM peak: algorithmic bytes / event time
  cpu_baseline  the OpenMP CPU port of the same loop (oracle/, test infrastructure) on a bounded sample

--impl reference times the UNMODIFIED reference implementation of the path, which is CUDA: oracle/_ref/libsrm_ref.so
(gcvt.cu compiled for sm_100 from /root/reference by oracle/Makefile) through its own entry point gCVT() on the same
GPU, north_star baseline (a).  If that library is missing or fails at this size, the OpenMP CPU port is timed instead.
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
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception as e:  # nvidia-smi missing
            log("[bench] clock sampling unavailable:", e)

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [nm for j, nm in enumerate(names) if any(len(r) > 3 + j and r[3 + j].startswith("Active") for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def cpu_baseline(dens, mask, vor, iters=2):
    """OpenMP port (oracle/srm_oracle.c:orc_fast_step) on the host cores: bounded sample of the same workload."""
    import _oracle as O
    f = O.FastLloyd(vor, dens, mask)
    f.step(2.0)  # warm-up (page faults, OpenMP pool)
    t0 = time.time()
    for _ in range(iters):
        f.step(2.0)
    dt = time.time() - t0
    n = dens.shape[0]
    return {"value": iters / dt, "unit": UNIT, "cores": O.lib().orc_num_threads(), "kind": "port",
            "sample": f"{iters} Lloyd iterations of the same {n}x{n} / {f.K}-site workload after 1 warm-up"}


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
    pinned = n <= 16384   # 32768^2: 9.7 GB of host buffers per rank stay pageable (8 ranks would pin 77 GB)
    dens, mask, vor = make_inputs(n, k, pinned=pinned)
    # row bands of equal work (sites per block of rows), from the replicated seed map: identical on every rank
    bands = S.row_bands_balanced(n, world, np.nonzero(vor[..., 0] != -32768)[0]) if args.bands == "balanced" \
        else S.row_bands(n, world)
    r0, r1 = bands[rank]

    eng = CudaBandEngine(n, r0, r1, local)
    eng.set_inputs(dens, mask, vor)
    sl = ShardedLloyd(n, rank, world, eng, dist if world > 1 else None, bands)
    if world > 1 and args.collective != "py":
        sl.bind_native_collective(args.collective)   # all-reduce inside libsrm's C++ loop (peer memory or NCCL)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- device-resident timing: W warm-up steps, then exactly K steps between events
    sl.run(W)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    sl.run(K)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    st = eng.state()

    # ---- per-stage pass (same loop, events between the stages) for the roofline object; N=1 only
    stage = None
    if world == 1 or args.collective != "py":
        stage = eng.ctx.iterate_profiled(K, stop_rule=False)   # every rank takes part in the all-reduce
        torch.cuda.synchronize()
        runs, ovf = eng.ctx.debug_counts()

    # ---- end to end through the reference-facing call, host buffers in pinned memory
    e2e_iters = args.e2e_iters
    if world == 1:
        eng.close()
        out = vor.copy() if False else None
        import torch as _t
        buf = _t.empty((n, n, 2), dtype=_t.int16, pin_memory=pinned).numpy()
        best = None
        for rep in range(2):  # first call warms the allocator / module load
            buf[:] = vor
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            est = S.gCVT(buf, dens, mask, n, 1, e2e_iters)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        e2e_val = est["iterations"] / best
        its = max(est["iterations"], 1)
        h2d = (dens.nbytes + mask.nbytes + vor.nbytes) / its
        d2h = buf.nbytes / its
    else:
        # each rank: upload its inputs, run the loop, download its band of labels
        eng.close()
        # set-up (like the process group): context, device buffers and the peer mappings / communicator; libsrm keeps
        # the site-indexed buffers across calls of the same size, so the mappings stay valid for the timed call
        eng2 = CudaBandEngine(n, r0, r1, local)
        eng2.set_inputs(dens, mask, vor)
        sl2 = ShardedLloyd(n, rank, world, eng2, dist, bands)
        if args.collective != "py":
            sl2.bind_native_collective(args.collective)
        barrier()
        t0 = time.perf_counter()
        eng2.set_inputs(dens, mask, vor)
        sl2.run(e2e_iters)
        lab = sl2.final_labels()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_val = e2e_iters / float(t.item())
        h2d = (dens.nbytes + mask.nbytes + vor.nbytes) / e2e_iters
        d2h = lab.nbytes / e2e_iters
        eng2.close()

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
        "config": {"workload": f"C3 curvature-like anisotropic density {n}x{n}, {k} sites + {int(mask.sum())} fixed boundary sites "
                               f"(BASELINE.json {'configs[2]' if n == 8192 else 'configs[3]' if n == 32768 else 'generator of configs[2], other size'})", "grid": n, "sites": st["num_sites"],
                   "parallelism": f"row bands x{world} ({args.bands}: {[b[1] - b[0] for b in bands]} rows), collective={args.collective}" if world > 1 else "single GPU",
                   "l2": "fp64 prefix arrays (24 B/px, read at run ends) + site-id map (4 B/px, read per run) "
                         f"= {28 * N / 1e6:.0f} MB > 126 MB L2; no explicit flush",
                   "stop_rule": "off (fixed step count); energy every 10th step like the reference"},
        "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "iterations_per_call": e2e_iters, "api": "gCVT(host buffers)" if world == 1 else "ShardedLloyd(host buffers)"},
        "gpu_launches": (6 + (1 if world > 1 and args.collective == "p2p" else 0)) * K * world,  # k_bits, k_carry, k_band, k_row, k_update_pos, k_update_resolve (+ k_signal) per step and rank
        "clocks": clocks,
    }
    if stage is not None and world > 1:
        line["config"]["stages_ms_per_step_rank0"] = {s: v / K for s, v in stage.items()}
    if stage is not None:
        row_ms = stage["band_fused"] / K
        # algorithmic bytes of one k_band launch (rank 0's band at N > 1): per 8-row band and column 8 B (bitmap word +
        # up/dn carries) = rows * n; per run 8 B written + 16 B fp64 prefix pair + 4 B site id + 24 B accumulator update = 52 B
        alg = 1.0 * (r1 - r0) * n + 52.0 * runs
        ach = alg / (row_ms / 1e3) / 1e9
        traffic = None   # dram__bytes_read.sum + dram__bytes_write.sum of k_band per launch, from the committed ncu capture
        tp = os.path.join(ROOT, "profiles", "r1_k_band_traffic.json")
        if n == 8192 and k == 100000 and world == 1 and os.path.exists(tp):
            tj = json.load(open(tp))
            traffic = 0.9 * tj["traffic_normal_step"] + 0.1 * tj["traffic_energy_step"]   # every 10th step computes the energy
        line["roofline"] = {"bound": "hbm", "kernel": "k_band (fused labelling + accumulation; rows*n B + 52 B/run" +
                                                      ("; rank 0's band)" if world > 1 else ")"),
                            "runs_per_step": runs, "robust_path_rows": ovf,
                            "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic,
                            "peak_source": peak_src, "ms_per_launch": row_ms,
                            "note": "latency / instruction-issue bound, not HBM bound: see DESIGN.md section 4",
                            "stages_ms_per_step": {s: v / K for s, v in stage.items()},
                            "step_bytes_per_px_equiv_GBs": {"8B_per_px": 8.0 * N / (ms / K / 1e3) / 1e9}}
    if not args.no_cpu and world == 1:   # reported baseline: rank 0 at N = 1 only
        try:
            line["cpu_baseline"] = cpu_baseline(dens, mask, vor, iters=args.cpu_iters)
        except Exception as e:  # pragma: no cover
            line["cpu_baseline"] = {"error": str(e)}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier(); dist.destroy_process_group()


# ------------------------------------------------------------------------------------------ reference

def _ref_child(n, k, steps, inputs):
    """Runs in a subprocess: the reference's gCVT on the same inputs, timed with CUDA events around the call.
    The inputs come from a file written by the parent: this process must not touch torch.cuda — the reference
    writes 4 MB past its pbaMargin allocation, which only goes unnoticed while its own
    cudaMalloc blocks are the neighbours."""
    import _ref as R
    z = np.load(inputs)
    dens, mask, vor = z["dens"], z["mask"], z["vor"]
    R.gcvt(vor, dens, mask, 2)                      # warm-up (context, module load)
    out, it, ms = R.gcvt(vor, dens, mask, steps, timed=True)
    loop_ms = R.loop_timed(vor, dens, mask, steps)   # device-resident loop body only
    print(json.dumps({"it": it, "ms": ms, "loop_ms": loop_ms, "mask": int(mask.sum())}), flush=True)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n, k, K, W = args.n, args.sites, args.steps, args.warmup
    import _ref as R
    line = {"impl": "reference", "metric": METRIC, "unit": UNIT, "n_gpus": args.gpus, "steps": K, "warmup": W,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "data": "synthetic",
            "config": {"workload": f"C3 curvature-like anisotropic density {n}x{n}, {k} sites (BASELINE.json configs[2])",
                       "grid": n}}
    res, crashes = None, []
    inputs = None
    if R.available():
        import tempfile
        dens, mask, vor = make_inputs(n, k, pinned=False)
        inputs = os.path.join(tempfile.gettempdir(), f"srm_ref_inputs_{os.getpid()}.npz")
        np.savez(inputs, dens=dens, mask=mask, vor=vor)
        # The reference CUDA is fragile at this size: on the B200 it dies with "illegal memory
        # access" (gpuErrchk at gcvt.cu:1150) for long runs.  Try the requested step count first, then shorter calls.
        for steps in [s for s in dict.fromkeys([K, 200, 100, 30, 10]) if s <= K]:
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
        v = res["it"] / (res["ms"] / 1e3)
        line["steps"] = res["it"]
        line.update({"value": v, "ms_per_step": res["ms"] / res["it"], "dtype": "short2 labels / f32 sums",
                     "cpu_baseline": {"value": v, "unit": UNIT, "cores": 0, "kind": "reference",
                                      "sample": f"reference CUDA gCVT() (oracle/_ref/libsrm_ref.so, built from the unmodified "
                                                f"gcvt.cu for sm_100) on the same B200, one call of {res['it']} iterations with host "
                                                "buffers; the reference has no CPU implementation of this path"},
                     "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                     "device_resident": {"value": res["steps"] / (res["loop_ms"] / 1e3), "unit": UNIT,
                                         "note": "reference loop body only, inputs already on the GPU"}})
    else:
        if not R.available():
            dens, mask, vor = make_inputs(n, k, pinned=False)
        cb = cpu_baseline(dens, mask, vor, iters=max(1, min(K, args.cpu_iters)))
        line.update({"value": cb["value"], "ms_per_step": 1e3 / cb["value"], "dtype": "int32 labels / f64 sums",
                     "cpu_baseline": cb, "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0,
                                                 "d2h_bytes_per_step": 0},
                     "note": "reference CUDA library unavailable or failed at this size: OpenMP CPU port timed instead"})
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", "--grid", dest="n", type=int, default=8192,
                    help="grid side (under torch.distributed.run spell it --grid: the launcher's own parser trips over --n)")
    ap.add_argument("--sites", type=int, default=100000)
    ap.add_argument("--e2e-iters", dest="e2e_iters", type=int, default=100)
    ap.add_argument("--cpu-iters", dest="cpu_iters", type=int, default=2)
    ap.add_argument("--no-cpu", action="store_true")
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
