"""Single QRCP launch for ncu (diagnostic): python tools/gpu_qrcp_one.py l n rank"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lowrankapprox.jl_b200"))
import numpy as np
import brapprox
l, n, rank = (int(a) for a in sys.argv[1:4])
ctx = brapprox.Context(0)
B = np.asfortranarray(np.random.default_rng(0).standard_normal((l, n)))
for rep in range(2):
    _, _, _, k, tr = brapprox.geqp3_adap(B, rank=rank, rtol=0.0, ctx=ctx)
print(k, tr["steps"], ctx.qrcp_phases())

import ctypes as C
out = (C.c_int32 * (148 * 8))()
brapprox.lib.bra_debug_qrcp_phases_all.argtypes = [C.c_void_p, C.POINTER(C.c_int32), C.c_int]
brapprox.lib.bra_debug_qrcp_phases_all(ctx.handle, out, 148)
a = np.array(out[:]).reshape(148, 8)[:, :5]
print("phase kcycles per CTA: min", a.min(0), "max", a.max(0), "argmin gather", a[:, 2].argmin(), "argmax gather", a[:, 2].argmax())
print(a[::8])
