import sys, os, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, bench
import _oracle as O
import surface_remesher_b200 as S
from surface_remesher_b200.sharded import _CudaArray
n = int(sys.argv[1]); k = int(sys.argv[2])
dens, mask, vor = bench.make_inputs(n, k, False)
if len(sys.argv) > 3 and sys.argv[3] == 'uniform':
    dens[:] = 1.0
for robust in (0, 1):
    with S.Context(n) as c:
        c.set_option("robust_only", robust)
        c.set_density(dens); c.set_mask(mask); c.set_site_map(vor)
        c.label(); runs, ovf = c.debug_counts()
        c.accumulate(True)
        c.synchronize()   # the context runs on its own non-blocking stream: finish k_acc before torch reads the buffer
        ptr, cnt = c.acc_buffer()
        acc = torch.as_tensor(_CudaArray(ptr, cnt), device="cuda").cpu().numpy()
        lab = c.get_labels()
        sites = S.api.unpack_sites(c.get_sites()).astype(np.int64)
    kc = (cnt - 4) // 4
    W = acc[0:4 * kc:4]
    d64 = dens.astype(np.float64)
    # CPU sums from the GPU labels
    ids = np.full(n * n, -1, np.int64); ids[sites[:, 1] * n + sites[:, 0]] = np.arange(len(sites))
    lid = ids[lab[..., 1].astype(np.int64) * n + lab[..., 0].astype(np.int64)]
    Wc = np.bincount(lid.ravel(), weights=d64.ravel(), minlength=kc)
    bad = np.nonzero(np.abs(W - Wc) > 1e-9 * (1 + np.abs(Wc)))[0]
    print(f"n={n} robust={robust} runs={runs} ovf_rows={ovf} mass_err={(W.sum()-d64.sum())/d64.sum():.3e} bad_sites={len(bad)}", file=sys.stderr)
    if len(bad):
        b = bad[:10]
        print(" sites(x,y):", sites[b].tolist(), "gpuW", W[b].tolist(), "cpuW", Wc[b].tolist(), file=sys.stderr)
        # rows: recompute per-row sums for the first bad site from labels
        s = b[0]; rows = np.unique(np.nonzero(lid == s)[0]); print(" first bad site rows", rows[:5], rows[-5:], file=sys.stderr)
    exp = O.label_exact(vor)
    print(" labels vs oracle mismatches:", int((lab != exp).any(axis=2).sum()), file=sys.stderr)
