"""On-box probe of the fused batched idfact kernel: time per batch and CTA-0 phase cycles.  Not a benchmark."""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lowrankapprox.jl_b200"))
import torch
import brapprox
from brapprox import _binding as B
from brapprox._frontend import idfact_batched_device

nb = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
m = n = 512
dev = torch.device("cuda", 0)
ctx = brapprox.Context(0)
g = torch.Generator(device=dev)
g.manual_seed(1)
At = torch.empty((nb, n, m), dtype=torch.float64, device=dev)
for c0 in range(0, nb, 1024):
    c1 = min(nb, c0 + 1024)
    x = torch.sort(torch.rand((c1 - c0, m), dtype=torch.float64, device=dev, generator=g), dim=1).values
    y = torch.sort(torch.rand((c1 - c0, n), dtype=torch.float64, device=dev, generator=g), dim=1).values + 1.02
    At[c0:c1] = 1.0 / (x[:, None, :] - y[:, :, None])
kd = torch.zeros(nb, dtype=torch.int64, device=dev)
pd = torch.zeros((nb, n), dtype=torch.int64, device=dev)
Td = torch.zeros((nb, n, 32), dtype=torch.float64, device=dev)
torch.cuda.synchronize()
ext = torch.cuda.ExternalStream(int(B.lib.bra_stream(ctx.handle)), device=dev)
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(ext)
    idfact_batched_device(At.data_ptr(), nb, m, n, m, m * n, kd.data_ptr(), pd.data_ptr(), Td.data_ptr(), 32, 32 * n,
                          None, ctx=ctx, rtol=1e-12, sketch="sprn", seed=rep)
    e1.record(ext)
    e1.synchronize()
    ms = e0.elapsed_time(e1)
ph = (ctypes.c_int64 * 8)()
B.lib.bra_debug_batched_phases(ctx.handle, ph)
per = -(-nb // 148)
print(json.dumps({"blocks": nb, "ms": ms, "us_per_block_per_sm": ms * 1e3 / per, "GBps": nb * m * n * 8 / ms / 1e6,
                  "cta0_cycles_per_block": {k: int(v) // per for k, v in zip(["tables", "tile_wait", "qrcp", "out_T", "issue", "gather", "handover", "-"], ph)}}))
