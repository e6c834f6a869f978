"""One launch of every hot kernel at its BASELINE shape, for `ncu` captures (no warm-up, no timing).
    ncu --set full --clock-control none --import-source on -k regex:'<kernels>' -o gpurun_out/prof python tools/ncu_targets.py"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lowrankapprox.jl_b200"))
import numpy as np
import torch
import brapprox
from brapprox import _binding as B
from brapprox._binding import DeviceMatrix
from brapprox._frontend import idfact_batched_device, psvdfact_device

what = set(sys.argv[1:] or ["c2", "c3", "c5"])
dev = torch.device("cuda", 0)
ctx = brapprox.Context(0)
torch.manual_seed(0)
if "c2" in what:
    n, r = 8192, 640
    U, _ = torch.linalg.qr(torch.randn(n, r, dtype=torch.float64, device=dev))
    V, _ = torch.linalg.qr(torch.randn(n, r, dtype=torch.float64, device=dev))
    s = 10.0 ** (-12.0 * torch.arange(r, dtype=torch.float64, device=dev) / 500.0)
    At = ((V * s) @ U.T).contiguous()
    A = DeviceMatrix(At.data_ptr(), n, n, n, keep=At)
    del U, V
    torch.cuda.synchronize()
    inf = psvdfact_device(A, rtol=1e-12, seed=1, ctx=ctx)
    print("c2 psvdfact k", inf.k, "ksvd", inf.ksvd, flush=True)
    del A, At
if "c3" in what:
    n = 16384
    A3 = torch.randn((n, n), dtype=torch.float64, device=dev)
    rng = np.random.default_rng(0)
    for order in (40, 520):
        d = torch.from_numpy(np.where(rng.random(n) > 0.5, 1.0, -1.0)).to(dev)
        idx = torch.from_numpy(rng.integers(1, n + 1, size=order)).to(dev)
        out = torch.empty((n, order), dtype=torch.float64, device=dev)
        ctx.check(B.lib.bra_sketch_srft_f64(ctx.handle, b"n", n, n, C.c_void_p(A3.data_ptr()), n, order,
                                            C.c_void_p(d.data_ptr()), C.c_void_p(idx.data_ptr()),
                                            C.c_void_p(out.data_ptr()), order))
    print("c3 srft done", flush=True)
    del A3
if "c5" in what:
    nb, m = 2048, 512
    g = torch.Generator(device=dev)
    g.manual_seed(1)
    x = torch.sort(torch.rand((nb, m), dtype=torch.float64, device=dev, generator=g), dim=1).values
    y = torch.sort(torch.rand((nb, m), dtype=torch.float64, device=dev, generator=g), dim=1).values + 1.02
    Ab = (1.0 / (x[:, None, :] - y[:, :, None])).contiguous()
    kd = torch.zeros(nb, dtype=torch.int64, device=dev)
    pd = torch.zeros((nb, m), dtype=torch.int64, device=dev)
    Td = torch.zeros((nb, m, 32), dtype=torch.float64, device=dev)
    torch.cuda.synchronize()
    idfact_batched_device(Ab.data_ptr(), nb, m, m, m, m * m, kd.data_ptr(), pd.data_ptr(), Td.data_ptr(), 32, 32 * m,
                          None, ctx=ctx, rtol=1e-12, sketch="sprn", seed=1)
    print("c5 batched done, k max", int(kd.max()), flush=True)
