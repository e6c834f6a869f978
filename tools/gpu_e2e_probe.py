"""On-box: where the end-to-end time of psvdfact with a HOST matrix goes (C2 shape)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lowrankapprox.jl_b200"))
import numpy as np, torch, brapprox
from brapprox import _binding as B
from brapprox._frontend import psvdfact_device, _rounds
dev = torch.device("cuda", 0)
ctx = brapprox.Context(0)
torch.manual_seed(0)
n, r = 8192, 640
U, _ = torch.linalg.qr(torch.randn(n, r, dtype=torch.float64, device=dev))
V, _ = torch.linalg.qr(torch.randn(n, r, dtype=torch.float64, device=dev))
s = 10.0 ** (-12.0 * torch.arange(r, dtype=torch.float64, device=dev) / 500.0)
At = ((V * s) @ U.T).contiguous()
Ah = torch.empty((n, n), dtype=torch.float64, pin_memory=True); Ah.copy_(At)
A = Ah.numpy().T
for rep in range(4):
    t0 = time.perf_counter()
    psvdfact_device(A, rtol=1e-12, seed=rep, ctx=ctx)
    t1 = time.perf_counter()
    inf, rounds, steps = _rounds(ctx)
    ks = int(inf.ksvd)
    Uh = ctx.fetch(B.F_U, (n, ks)); t2 = time.perf_counter()
    Sh = ctx.fetch(B.F_S, (ks,)); Vh = ctx.fetch(B.F_VT, (ks, n)); t3 = time.perf_counter()
    print("rep %d: psvdfact(host A) %.2f ms | fetch U %.2f ms | fetch S,Vt %.2f ms | total %.2f" %
          (rep, (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, (t3 - t0) * 1e3), flush=True)
