# parity of the panel path vs direct on a tall matrix (stage-wise sketch), on-box
import os, sys, ctypes as C
sys.path.insert(0, "lowrankapprox.jl_b200")
import numpy as np, torch, brapprox
from brapprox import _binding as B
ctx = brapprox.Context(0)
dev = torch.device("cuda", 0)
m, n, l = 140000, 512, 72
A = torch.randn((n, m), dtype=torch.float64, device=dev)       # column-major m x n
Om = torch.randn((m, l), dtype=torch.float64, device=dev)      # column-major l x m
out = torch.empty((n, l), dtype=torch.float64, device=dev)
ctx.check(B.lib.bra_sketch_randn_f64(ctx.handle, b"n", m, n, C.c_void_p(A.data_ptr()), m, l, C.c_void_p(Om.data_ptr()), l,
                                     C.c_void_p(out.data_ptr()), l))
ref = (Om.T @ A.T)          # l x n
err = (out.T - ref).abs().max().item() / ref.abs().max().item()
print("panel path rel err", err)
assert err < 1e-12
