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
    import ctypes
    from brapprox import _binding as Bd
    allph = (ctypes.c_int32 * (160 * 8))()
    Bd.lib.bra_debug_qrcp_phases_all(ctx.handle, allph, 148)
    arr = np.array(list(allph)).reshape(160, 8)[:148, :6] * 1024.0 / max(steps, 1)
    print("  per-CTA cycles/step  min", arr.min(0).astype(int).tolist(), "med", np.median(arr, 0).astype(int).tolist(),
          "max", arr.max(0).astype(int).tolist(), "argmax", arr.argmax(0).tolist(), flush=True)
    print(json.dumps({"l": l, "n": n, "steps": steps, "qrcp_ms": prof["qrcp"][0], "us_per_step": prof["qrcp"][0] * 1e3 / max(steps, 1),
                      "kcycles": ph, "cycles_per_step": {k2: v * 1024 / max(steps, 1) for k2, v in ph.items()}}), flush=True)
