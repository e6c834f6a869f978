"""Quick on-box probes: FP64 DMMA/DFMA peaks, LL exchange latency, rough idfact timing (fast mode).
Writes gpurun_out/probe.json.  Not a benchmark: bench.py is."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lowrankapprox.jl_b200"))
import numpy as np
import brapprox
from brapprox._binding import DeviceMatrix

out = {}
ctx = brapprox.Context(0)
out["fp64_peak_tflops"] = brapprox.probe_fp64_peak(ctx)
out["exchange_us"] = {g: brapprox.probe_exchange_latency(g, 2000, ctx) for g in (8, 32, 74, 148)}
print(json.dumps(out), flush=True)

import torch
torch.manual_seed(0)
for n in (int(a) for a in (sys.argv[1:] or ["2048", "8192"])):
    r = 640
    U, _ = torch.linalg.qr(torch.randn(n, r, dtype=torch.float64, device="cuda"))
    V, _ = torch.linalg.qr(torch.randn(n, r, dtype=torch.float64, device="cuda"))
    s = 10.0 ** (-12.0 * torch.arange(r, dtype=torch.float64, device="cuda") / 500.0)
    At = (V * s) @ U.T            # row-major (n x n) == column-major A = U diag(s) V^T
    A = DeviceMatrix(At.data_ptr(), n, n, n, keep=At)
    torch.cuda.synchronize()
    from brapprox._frontend import idfact_device
    for rep in range(3):
        ctx.profile_enable(rep == 2)
        t0 = time.perf_counter()
        inf = idfact_device(A, rtol=1e-12, seed=1, ctx=ctx)
        dt = time.perf_counter() - t0
    prof = ctx.profile_read()
    ctx.profile_enable(False)
    rounds = [(int(inf.orders[t]), int(inf.ks[t]), int(inf.steps[t])) for t in range(inf.rounds)]
    out[f"idfact_n{n}"] = {"wall_ms": dt * 1e3, "k": int(inf.k), "rounds": rounds, "prof_ms": prof}
    print(json.dumps(out[f"idfact_n{n}"]), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "probe.json"), "w"), indent=1)
