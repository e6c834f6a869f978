"""On-box probe: the blocked tall QRCP (qrcp_blocked.cu) against the persistent rank-1 kernel on sketch = :none shapes.
   BRA_QRCP_BLOCKED=0 selects the old kernel.  Not a benchmark."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lowrankapprox.jl_b200"))
import numpy as np
import torch
import brapprox
from brapprox._binding import DeviceMatrix
from brapprox._frontend import idfact_device

ctx = brapprox.Context(0)
torch.manual_seed(0)
for (m, n, r) in [(8192, 8192, 640), (16384, 4096, 640), (4096, 1024, 300)] + ([(65536, 2048, 300)] if os.environ.get("BRA_QRCP_BLOCKED") != "0" else []):
    U, _ = torch.linalg.qr(torch.randn(m, r, dtype=torch.float64, device="cuda"))
    V, _ = torch.linalg.qr(torch.randn(n, r, dtype=torch.float64, device="cuda"))
    s = 10.0 ** (-12.0 * torch.arange(r, dtype=torch.float64, device="cuda") / 500.0)
    At = ((V * s) @ U.T).contiguous()
    A = DeviceMatrix(At.data_ptr(), m, n, m, keep=At)
    torch.cuda.synchronize()
    ms = []
    for rep in range(3):
        ctx.profile_enable(True)
        t0 = time.perf_counter()
        inf = idfact_device(A, rtol=1e-10, sketch="none", ctx=ctx)
        torch.cuda.synchronize()
        ms.append((time.perf_counter() - t0) * 1e3)
        prof = ctx.profile_read()
    steps = int(inf.steps[0])
    ph = ctx.qrcp_phases()
    print(json.dumps({"shape": [m, n], "k": int(inf.k), "steps": steps, "idfact_none_ms": round(min(ms), 3),
                      "qrcp_ms": round(prof["qrcp"][0], 3), "us_per_step": round(prof["qrcp"][0] * 1e3 / max(steps, 1), 2),
                      "kcycles_select_owner_fpass_exchange": list(ph.values())[:4], "GBps_trailing_reads": round(8.0 * m * n * steps / (prof["qrcp"][0] * 1e-3) / 1e9 / 2, 1)}), flush=True)
    del A, At, U, V
