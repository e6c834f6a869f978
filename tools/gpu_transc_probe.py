"""On-box probe: idfact(:n) against idfact(:c) at the C2 shape, and psvdfact of a wide 4096 x 8192 matrix (which takes the
(:left, :c) sketch).  Not a benchmark."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lowrankapprox.jl_b200"))
import brapprox
from brapprox._binding import DeviceMatrix
from brapprox._frontend import idfact_device, psvdfact_device
import torch

ctx = brapprox.Context(0)
torch.manual_seed(0)


def mat(m, n, r=640):
    U, _ = torch.linalg.qr(torch.randn(m, r, dtype=torch.float64, device="cuda"))
    V, _ = torch.linalg.qr(torch.randn(n, r, dtype=torch.float64, device="cuda"))
    s = 10.0 ** (-12.0 * torch.arange(r, dtype=torch.float64, device="cuda") / 500.0)
    At = (V * s) @ U.T                     # n x m row-major == m x n column-major
    return DeviceMatrix(At.data_ptr(), m, n, m, keep=At)


def timed(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ctx.profile_enable(True)
    t0 = time.perf_counter()
    for _ in range(reps):
        inf = fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    prof = ctx.profile_read()
    ctx.profile_enable(False)
    return dt * 1e3, inf, {k: round(v[0] / reps, 3) for k, v in prof.items() if v[0] > 0}


A = mat(8192, 8192)
for tr in ("n", "c"):
    ms, inf, prof = timed(lambda: idfact_device(A, rtol=1e-12, seed=1, trans=tr, ctx=ctx))
    print(json.dumps({"what": f"idfact trans={tr} 8192^2", "ms": round(ms, 3), "k": int(inf.k), "stage_ms": prof}), flush=True)
W = mat(4096, 8192)
ms, inf, prof = timed(lambda: psvdfact_device(W, rtol=1e-12, seed=1, ctx=ctx))
print(json.dumps({"what": "psvdfact 4096x8192 (trans=:c inside)", "ms": round(ms, 3), "k": int(inf.k), "stage_ms": prof}), flush=True)
T = mat(8192, 4096)
ms, inf, prof = timed(lambda: psvdfact_device(T, rtol=1e-12, seed=1, ctx=ctx))
print(json.dumps({"what": "psvdfact 8192x4096 (trans=:n inside)", "ms": round(ms, 3), "k": int(inf.k), "stage_ms": prof}), flush=True)
