"""On-box: raw pinned H2D / D2H bandwidth vs the library's staging path for the C2 matrix (537 MB)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lowrankapprox.jl_b200"))
import torch
n = 8192
h = torch.empty((n, n), dtype=torch.float64, pin_memory=True); h.normal_()
d = torch.empty((n, n), dtype=torch.float64, device="cuda")
for _ in range(2): d.copy_(h, non_blocking=True); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): d.copy_(h, non_blocking=True)
e1.record(); e1.synchronize()
ms = e0.elapsed_time(e1) / 5
print("torch pinned H2D: %.2f ms  %.1f GB/s" % (ms, h.numel() * 8 / ms / 1e6))
e0.record()
for _ in range(5): h.copy_(d, non_blocking=True)
e1.record(); e1.synchronize()
ms = e0.elapsed_time(e1) / 5
print("torch pinned D2H: %.2f ms  %.1f GB/s" % (ms, h.numel() * 8 / ms / 1e6))
import brapprox, numpy as np
ctx = brapprox.Context(0)
A = h.numpy().T
for rep in range(3):
    t0 = time.perf_counter()
    V = brapprox.idfact(A, rtol=1e-12, seed=rep, ctx=ctx)
    t1 = time.perf_counter()
print("idfact host A: %.2f ms (device-resident idfact is ~11.8 ms)" % ((t1 - t0) * 1e3))
