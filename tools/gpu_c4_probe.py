"""On-box probe: sketch-GEMM rate of the tall config as a function of the row count (column stride of A)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "lowrankapprox.jl_b200"))
import torch
import brapprox
from brapprox import _binding as B
import bench
dev = torch.device("cuda", 0)
ctx = brapprox.Context(0)
ext = torch.cuda.ExternalStream(int(B.lib.bra_stream(ctx.handle)), device=dev)
for rows in [int(a) for a in (sys.argv[1:] or ["32768", "131072", "524288"])]:
    r = bench.run_c4(ctx, ext, dev, 0, 1, rows, 4096, steps=2)
    rounds = r["rounds_order_k"]
    f_sk = 2.0 * rows * 4096 * sum(l for l, _ in rounds)
    g = r["stage_ms"]["gemm"]
    print(json.dumps({"rows": rows, "rounds": rounds, "gemm_ms": g, "gemm_tflops": f_sk / g / 1e9, "t_ms": r["t"] * 1e3,
                      "stage_ms": {k: round(v, 2) for k, v in r["stage_ms"].items()}}), flush=True)
    torch.cuda.empty_cache()
