"""Timeline of ONE pivot step of the persistent QRCP kernel (diagnostic): clock64 stamps per warp and CTA.
   BRA_QRCP_TS_STEP=<step> python tools/gpu_qrcp_trace.py l n rank"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lowrankapprox.jl_b200"))
import numpy as np
import brapprox
l, n, rank = (int(a) for a in sys.argv[1:4])
ctx = brapprox.Context(0)
B = np.asfortranarray(np.random.default_rng(0).standard_normal((l, n)))
for rep in range(2):
    _, _, _, k, tr = brapprox.geqp3_adap(B, rank=rank, rtol=0.0, ctx=ctx)
G = 148
out = (C.c_int64 * (G * 16 * 16))()
brapprox.lib.bra_debug_qrcp_trace.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.c_int]
rc = brapprox.lib.bra_debug_qrcp_trace(ctx.handle, out, G)
t = np.array(out[:], dtype=np.int64).reshape(G, 16, 16)
names = {0: "comm: step top (after bar2)", 1: "comm: header pushed", 2: "comm: winner known", 3: "comm: fetched (-> bar1)",
         4: "cmp: top (after bar2)", 5: "cmp: owner dlarfg done", 6: "cmp: pass2 done (-> bar1)", 7: "cmp: after bar1", 8: "cmp: pass1+publish done"}
print("rc", rc, "step", os.environ.get("BRA_QRCP_TS_STEP"), "l", l, "n", n, "steps", tr["steps"])
for c in (0, 73, 140):
    base = t[c, 15, 0]
    print("CTA", c, "stamps relative to the comm warp's step top; compute warps 0..14")
    for k_, nm in names.items():
        row = t[c, :, k_]
        if k_ < 4: print(f"  {nm:30s} {row[15] - base:6d}")
        else: print(f"  {nm:30s}", " ".join(f"{(v - base) if v else -1:6d}" for v in row[:15]))
# global-timer view (ns) of the comm warps: when did every CTA push its header / learn the winner / finish the fetch
g = t[:, 15, 9:13].astype(float)
g0 = g[:, 0].min()
for k_, nm in enumerate(["top", "pushed", "winner known", "fetched"]):
    x = g[:, k_] - g0
    print(f"globaltimer ns  {nm:13s} min {x.min():8.0f} med {np.median(x):8.0f} max {x.max():8.0f}  argmax CTA {int(x.argmax())}  argmin CTA {int(x.argmin())}")
late = np.argsort(g[:, 1])[-8:]
print("latest pushers (CTA: top, pushed):", [(int(c), int(g[c, 0] - g0), int(g[c, 1] - g0)) for c in late])
wc = int(t[0, 15, 13])
print("winner CTA of this step:", wc)
c64 = t.astype(float)
dur = {"top->push": c64[:, 15, 1] - c64[:, 15, 0], "push->winner": c64[:, 15, 2] - c64[:, 15, 1], "winner->fetched": c64[:, 15, 3] - c64[:, 15, 2],
       "pass2 (max warp)": (c64[:, :15, 6] - c64[:, :15, 4]).max(axis=1), "owner dlarfg (max)": (c64[:, :15, 5] - c64[:, :15, 4]).max(axis=1),
       "pass1 (max warp)": (c64[:, :15, 8] - c64[:, :15, 7]).max(axis=1), "pass1 (min warp)": (c64[:, :15, 8] - c64[:, :15, 7]).min(axis=1)}
order = np.argsort(g[:, 0])
print("CTAs by step-top time (ns):  first 5", [(int(c), int(g[c, 0] - g0)) for c in order[:5]], " last 8", [(int(c), int(g[c, 0] - g0)) for c in order[-8:]])
for nm, d in dur.items():
    print(f"{nm:20s} median {np.median(d):7.0f}  max {d.max():7.0f} (CTA {int(d.argmax())})   late CTAs:", [int(d[c]) for c in order[-8:]])
# inside the owner's dlarfg: sweep done / reduction done, relative to the warp's own step top
own = [(c, w) for c in range(G) for w in range(15) if t[c, w, 14] > 0]
if own:
    d1 = np.array([t[c, w, 14] - t[c, w, 4] for c, w in own]); d2 = np.array([t[c, w, 15] - t[c, w, 4] for c, w in own]); d3 = np.array([t[c, w, 5] - t[c, w, 4] for c, w in own])
    print(f"owner warps ({len(own)}): sweep done median {np.median(d1):.0f}, +reduction {np.median(d2):.0f}, dlarfg done {np.median(d3):.0f} cycles after their step top")
