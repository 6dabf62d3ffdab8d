#!/usr/bin/env python
"""Write-only HBM bandwidth of the device (the roofline of a store-bound kernel such as calc_ao): torch fill_, cudaMemset
and a copy (read + write) of the same size for comparison.  Prints one JSON line."""
import json, sys
import torch
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 8_000_000_000
dev = torch.device('cuda', 0)
a = torch.empty(n // 8, dtype=torch.float64, device=dev)
b = torch.empty(n // 8, dtype=torch.float64, device=dev)
def best(fn, reps=6):
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts[1:])
t_fill = best(lambda: a.fill_(1.5))
t_zero = best(lambda: a.zero_())
t_copy = best(lambda: b.copy_(a))
print(json.dumps({'bytes': n, 'fill_gbs': n / t_fill / 1e6, 'memset_gbs': n / t_zero / 1e6,
                  'copy_read_plus_write_gbs': 2 * n / t_copy / 1e6}))
