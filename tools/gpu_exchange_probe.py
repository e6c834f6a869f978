import ctypes as C, os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lowrankapprox.jl_b200"))
import brapprox
lib = brapprox.lib
lib.bra_probe_exchange2.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double)]
ctx = brapprox.Context(0)
def run(g, hw, mode, L):
    us = C.c_double(0)
    ctx.check(lib.bra_probe_exchange2(ctx.handle, g, hw, mode, L, 3000, C.byref(us)))
    return round(us.value, 3)
for g in (148, 74):
    for hw in (1, 2, 3, 5):
        print("push  G=%d hw=%d  %.3f us" % (g, hw, run(g, hw, 0, 1)), flush=True)
    for L in (4, 8, 12, 16, 37):
        for hw in (1, 5):
            print("2hop  G=%d L=%d hw=%d  %.3f us" % (g, L, hw, run(g, hw, 1, L)), flush=True)
print("pull1 G=148 %.3f us" % brapprox.probe_exchange_latency(148, 3000, ctx))
