"""On-box probe: SRFT sketch kernel time / achieved HBM bandwidth at the C3 shape for each adaptive order."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lowrankapprox.jl_b200"))
import numpy as np
import torch
import ctypes as C
import brapprox
from brapprox import _binding as B

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
dev = torch.device("cuda", 0)
ctx = brapprox.Context(0)
A = torch.randn((n, n), dtype=torch.float64, device=dev)
rng = np.random.default_rng(0)
for order in (40, 72, 136, 264, 520):
    d = torch.from_numpy(np.where(rng.random(n) > 0.5, 1.0, -1.0)).to(dev)
    idx = torch.from_numpy(rng.integers(1, n + 1, size=order)).to(dev)
    out = torch.empty((n, order), dtype=torch.float64, device=dev)
    for rep in range(3):
        ctx.profile_enable(True)
        ctx.check(B.lib.bra_sketch_srft_f64(ctx.handle, b"n", n, n, C.c_void_p(A.data_ptr()), n, order,
                                            C.c_void_p(d.data_ptr()), C.c_void_p(idx.data_ptr()),
                                            C.c_void_p(out.data_ptr()), order))
        prof = ctx.profile_read()
    ms = prof["sketch_other"][0]
    print(json.dumps({"order": order, "srft_ms": round(ms, 3), "splitk_ms": round(prof["splitk"][0], 3),
                      "GBps": round((8.0 * n * n + 8.0 * order * n) / ms / 1e6, 1)}), flush=True)
