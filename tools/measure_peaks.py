#!/usr/bin/env python
"""Write-only and read-only stream rates of this GPU next to the copy rate of MEASURED_PEAKS.json: the ceilings for
k_expand (4 B/px written, nothing read per pixel) and k_prefix (4 B/px read, 24 B/px written)."""
import json
import torch
n = 1 << 28   # 1 GiB of float32
a = torch.empty(n, dtype=torch.float32, device="cuda")
b = torch.empty(n, dtype=torch.float32, device="cuda")
def t(fn, reps=10):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
ms_fill = t(lambda: a.fill_(1.0))
ms_zero = t(lambda: torch.cuda.memset(a.data_ptr(), 0, 4 * n) if hasattr(torch.cuda, "memset") else a.zero_())
ms_copy = t(lambda: b.copy_(a))
ms_sum = t(lambda: a.sum())
print(json.dumps({"write_only_fill_GBs": 4 * n / ms_fill / 1e6, "write_only_zero_GBs": 4 * n / ms_zero / 1e6,
                  "copy_read_plus_write_GBs": 8 * n / ms_copy / 1e6, "read_only_sum_GBs": 4 * n / ms_sum / 1e6}))
