"""On-box: QRCP time per sketch order on random short-wide sketches (CUDA events inside the library)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lowrankapprox.jl_b200"))
import numpy as np
import brapprox
ctx = brapprox.Context(0)
rng = np.random.default_rng(0)
shapes = [(40, 8192, 40), (72, 8192, 72), (136, 8192, 136), (264, 8192, 264), (520, 8192, 498), (520, 16384, 498), (264, 4096, 256), (40, 1024, 27)]
if len(sys.argv) > 1: shapes = [tuple(int(x) for x in a.split(",")) for a in sys.argv[1:]]
tot = 0.0
for (l, n, rank) in shapes:
    B = np.asfortranarray(rng.standard_normal((l, n)))
    ms = []
    for rep in range(4):
        ctx.profile_enable(True)
        _, _, _, k, tr = brapprox.geqp3_adap(B, rank=rank, rtol=0.0, ctx=ctx)
        ms.append(ctx.profile_read()["qrcp"][0])
    best = min(ms[1:])
    print(json.dumps({"l": l, "n": n, "steps": tr["steps"], "qrcp_ms": round(best, 4), "us_per_step": round(best * 1e3 / max(tr["steps"], 1), 3)}), flush=True)
