"""On-box probe: per-stage CUDA-event timing of psvdfact at the C2 shape + Jacobi sweep count.  Not a benchmark."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lowrankapprox.jl_b200"))
import brapprox
from brapprox import _binding as B
from brapprox._binding import DeviceMatrix
import torch

ctx = brapprox.Context(0)
torch.manual_seed(0)
for n in (int(a) for a in (sys.argv[1:] or ["8192"])):
    r = 640
    U, _ = torch.linalg.qr(torch.randn(n, r, dtype=torch.float64, device="cuda"))
    V, _ = torch.linalg.qr(torch.randn(n, r, dtype=torch.float64, device="cuda"))
    s = 10.0 ** (-12.0 * torch.arange(r, dtype=torch.float64, device="cuda") / 500.0)
    At = (V * s) @ U.T
    A = DeviceMatrix(At.data_ptr(), n, n, n, keep=At)
    torch.cuda.synchronize()
    from brapprox._frontend import psvdfact_device
    for rep in range(3):
        ctx.profile_enable(rep == 2)
        t0 = time.perf_counter()
        inf = psvdfact_device(A, rtol=1e-12, seed=1, ctx=ctx)
        dt = time.perf_counter() - t0
    prof = ctx.profile_read()
    ctx.profile_enable(False)
    B.lib.bra_debug_jacobi_sweeps.restype = int
    import ctypes
    ph = (ctypes.c_int32 * 8)()
    B.lib.bra_debug_jacobi_phases(ctx.handle, ph)
    print(json.dumps({"n": n, "wall_ms": dt * 1e3, "k": int(inf.k), "ksvd": int(inf.ksvd),
                      "jacobi_sweeps": B.lib.bra_debug_jacobi_sweeps(ctx.handle), "jacobi_kcyc_load_rot_store_bar": list(ph),
                      "prof_ms": {k: round(v[0], 3) for k, v in prof.items()}}), flush=True)
