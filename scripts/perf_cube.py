#!/usr/bin/env python
"""Cube output sink (okb_format_cube): device-resident kernel time against the HBM roofline, the host-to-host call, and
the reference's formatting loop (oracle/oracle_out.py = cube.py:86-96) on a bounded sample.  One JSON line.
    python scripts/perf_cube.py [N]      (N^3 values, default 200)"""
import json, os, sys, time
import numpy, torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, 'oracle'))
from orbkit_b200 import _lib
from orbkit_b200.engine import get_engine
N = int(sys.argv[1]) if len(sys.argv) > 1 else 200
eng = get_engine()
dev = torch.device('cuda', eng.device)
stream = torch.cuda.ExternalStream(eng.stream_ptr(), device=dev)
rng = numpy.random.default_rng(0)
host = rng.normal(size=(1, N, N, N)) * 10.0 ** rng.integers(-8, 3, size=(1, N, N, N))
data = torch.from_numpy(host).to(dev)
nbytes = eng.lib.okb_cube_body_bytes(1, N, N, N)
text = torch.empty(nbytes, dtype=torch.uint8, device=dev)
torch.cuda.synchronize()
flags = _lib.OKB_FLAG_IN_DEVICE | _lib.OKB_FLAG_OUT_DEVICE
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
times = []
for it in range(8):
    with torch.cuda.stream(stream):
        flush.fill_(it)                                   # L2 flush between timed launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        _lib.check(eng.lib.okb_format_cube(eng.ctx, data.data_ptr(), 1, N, N, N, text.data_ptr(), nbytes, flags))
        e1.record(stream)
    eng.sync()
    times.append(e0.elapsed_time(e1))
ms = float(numpy.median(times[3:]))
alg = 8.0 * N ** 3 + nbytes
peak = json.load(open(os.path.join(REPO, 'MEASURED_PEAKS.json')))['hbm_gbs']
# host to host (pageable NumPy in, bytes out), as output.cube_body does
from orbkit_b200 import output
output.cube_body(host)
t0 = time.perf_counter()
body = output.cube_body(host)
t_host = time.perf_counter() - t0
# the reference's loop on a sample of rows
import oracle_out
sample = host[:, :1, :80]
t0 = time.perf_counter()
ref = oracle_out.cube_body(sample)
t_ref = time.perf_counter() - t0
assert body.tobytes()[:len(ref)] == ref
print(json.dumps({'what': 'cube text, %d^3 values' % N, 'kernel_ms': ms, 'values_per_s': N ** 3 / ms * 1e3,
                  'roofline': {'bound': 'hbm', 'achieved': alg / ms / 1e6, 'peak': peak, 'unit': 'GB/s',
                               'frac': alg / ms / 1e6 / peak, 'alg_bytes_per_value': alg / N ** 3},
                  'host_to_host_ms': t_host * 1e3, 'host_to_host_values_per_s': N ** 3 / t_host,
                  'cpu_baseline': {'values_per_s': sample.size / t_ref, 'cores': 1, 'kind': 'port',
                                   'sample': '%d values (80 rows)' % sample.size}}))
