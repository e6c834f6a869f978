"""Per-phase cost of the persistent QRCP kernel on random sketches (diagnostic)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lowrankapprox.jl_b200"))
import numpy as np
import brapprox
ctx = brapprox.Context(0)
rng = np.random.default_rng(0)
for (l, n, rank) in [(40, 8192, 40), (136, 8192, 136), (520, 8192, 500), (520, 2048, 500), (40, 1024, 27)]:
    B = np.asfortranarray(rng.standard_normal((l, n)))
    for rep in range(2):
        ctx.profile_enable(True)
        _, _, _, k, tr = brapprox.geqp3_adap(B, rank=rank, rtol=0.0, ctx=ctx)
        prof = ctx.profile_read()
    ph = ctx.qrcp_phases()
    steps = tr["steps"]
    print(json.dumps({"l": l, "n": n, "steps": steps, "qrcp_ms": prof["qrcp"][0], "us_per_step": prof["qrcp"][0] * 1e3 / max(steps, 1),
                      "kcycles": ph, "cycles_per_step": {k2: v * 1024 / max(steps, 1) for k2, v in ph.items()}}), flush=True)
