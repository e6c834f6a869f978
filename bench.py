#!/usr/bin/env python
"""bench.py -- headline benchmark of the sketch-then-factor path (BASELINE.json config 2).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--n 8192] [--what idfact|psvdfact]

A "step" is ONE factorization (idfact, or psvdfact once requested) of an n x n FP64 matrix
A = U diag(sigma) V^T, sigma_j = 10^(-12 j / 500), rank-640 factors from thin QRs of seeded Gaussians
(SURVEY.md section 8d, C2), rtol = 1e-12, sketch = :randn, adaptive rounds, device Philox Omega.

Our arm prints ONE JSON line:
  value         factorizations/s, A resident in HBM, results left on the device, only k read back
  e2e           the same through the C ABI with a HOST (pinned) A: H2D of A and D2H of (p, T[, U, S, Vt])
                inside the timed region every step
  roofline      FP64-tensor bound: algorithmic flops of the sketch GEMM launches / their CUDA-event time
                (events recorded on the library's own stream), against the FP64 peak measured in this run
  cpu_baseline  the oracle (real LAPACK/BLAS through scipy's OpenBLAS) on the box's host cores, bounded sample

`--impl reference` times the oracle only (rank 0), same metric/config.  N > 1: one process per GPU, each
factorizing its own replica (the single-matrix configs do not shard: SURVEY.md section 8e, "replicas only").
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "lowrankapprox.jl_b200"))

RTOL = 1e-12
RANK_GEN = 640
DECADES, JDIV = 12.0, 500.0


def workload_label(what: str, n: int) -> str:
    """The SAME string in both arms (ours and --impl reference): the two lines describe one workload."""
    return (f"C2: {what} of {n}x{n} FP64, A = U diag(10^(-12j/500)) V^T (rank-640 factors from thin QRs of seeded "
            "Gaussians), rtol=1e-12, sketch=randn, adaptive rounds")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=8192)
    ap.add_argument("--what", default="auto", choices=["auto", "idfact", "psvdfact"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the C3/C4/C5 side measurements")
    ap.add_argument("--c5-blocks", type=int, default=16384)
    ap.add_argument("--c4-rows", type=int, default=1048576)
    return ap.parse_args()


def algorithmic_flops(m, n, rounds, steps, k):
    """F_sk = 2 m n sum(l_t); F_qr = sum 4[l n k - (l+n)k^2/2 + k^3/3]; F_T = k^2 (n-k)."""
    f_sk = 2.0 * m * n * sum(l for l, _ in rounds)
    f_qr = sum(4.0 * (l * n * s - (l + n) * s * s / 2.0 + s ** 3 / 3.0) for (l, _), s in zip(rounds, steps))
    f_t = float(k) * k * (n - k)
    return f_sk, f_qr, f_t


def psvd_extra_flops(m, n, k):
    return (2.0 * m * k * k - 2.0 / 3.0 * k ** 3) + float(k) * k * (n - k) + 4.0 * n * k * k + 12.0 * k ** 3 + 2.0 * m * k * k


class ClockSampler:
    """Samples nvidia-smi clocks and throttle reasons DURING the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self) -> dict:
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=2)
        except Exception:
            self.p.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for nm, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_sample(n: int, what: str, max_seconds: float = 25.0, threads: int = 0, steps: int = 0, warmup: int = 0):
    """Times the oracle on the host cores: factorizations/s on a bounded sample of the same workload.
    steps == 0: the cpu_baseline leg (median of up to 12 factorizations within max_seconds).
    steps > 0: the reference arm -- `warmup` untimed factorizations, then `steps` timed ones (value = count / total time,
    like the GPU arm), cut short only if max_seconds is exceeded (the counts actually run are reported)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import numpy as np
    import lra_oracle as o
    cores = threads or os.cpu_count() or 1
    o.set_blas_threads(cores)
    A = o.decaying_matrix(n, n, RANK_GEN, DECADES, JDIV, seed=1)
    opts = o.LRAOptions(rtol=RTOL)
    fn = o.psvdfact if what == "psvdfact" else o.idfact
    times, ks = [], []
    t_start = time.perf_counter()
    nwarm = 0
    for w in range(warmup if steps > 0 else 0):
        if w >= 1 and time.perf_counter() - t_start > max_seconds / 4:
            break
        fn(A, opts, o.RandomInputs(1000 + w))
        nwarm += 1
    t_start = time.perf_counter()
    rep = 0
    while True:
        t0 = time.perf_counter()
        F = fn(A, opts, o.RandomInputs(rep))
        times.append(time.perf_counter() - t0)
        ks.append(len(F.S) if what == "psvdfact" else F.k)
        rep += 1
        if steps > 0:
            if rep >= steps or (rep >= 2 and time.perf_counter() - t_start > max_seconds):
                break
        elif rep >= 2 and (time.perf_counter() - t_start > max_seconds or rep >= 12):
            break
    per = (sum(times) / len(times)) if steps > 0 else sorted(times)[len(times) // 2]
    stat = "mean" if steps > 0 else "median"
    # the error metric of BASELINE.json on the last factorization: snormdiff(A, F) / snorm(A)   (||A||_2 = 1 here)
    try:
        if what == "psvdfact":
            err = o.snormdiff_lowrank(A, F.U * F.S, F.Vt) / o.snorm_dense(A)
        else:
            err = o.id_error(A, F)
    except Exception:
        err = None
    return {"value": 1.0 / per, "unit": "factorizations/s", "cores": cores, "kind": "port",
            "snormdiff_over_snorm": err,
            "sample": f"{rep} x {what} of the same {n}x{n} workload on the host ({stat}; includes drawing Omega "
                      f"with numpy), OpenBLAS threads={o.get_blas_threads()}, k={ks[-1]}"}, times, nwarm


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # same metric as our arm: `auto` means psvdfact there (the library has the psvd tail), so it does here
    what = "psvdfact" if args.what == "auto" else args.what
    # --steps / --warmup are honoured inside a bounded budget (one factorization takes seconds on the host)
    base, times, nwarm = cpu_sample(args.n, what, max_seconds=150.0, steps=max(1, args.steps), warmup=args.warmup)
    v = base["value"]
    line = {"impl": "reference", "metric": f"{what}_factorizations_per_sec", "value": v, "unit": "factorizations/s",
            "n_gpus": args.gpus, "steps": len(times), "warmup": nwarm, "ms_per_step": 1e3 / v, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_label(what, args.n),
                       "arm": "oracle = the reference's LAPACK/BLAS CPU path restated (Omega drawn with numpy)"},
            "cpu_baseline": base,
            "e2e": {"value": v, "unit": "factorizations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def hbm_peak():
    """HBM roofline denominator: the driver's measured copy bandwidth if present, else the guide's fallback."""
    try:
        d = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(d["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
    except Exception:
        return 6650.0, "B200_PROFILING.md fallback 6.65 TB/s (of fallback; MEASURED_PEAKS.json absent)"


def ncu_traffic_per_launch():
    """dram__bytes_read.sum + dram__bytes_write.sum of the sketch GEMM from the committed `ncu --set full` capture
    (profiles/r02e_gemm_sketch_raw.csv: one psvdfact at the C2 shape, the five nested-sketch launches of 40, 32, 64, 128
    and 256 new rows), averaged per launch."""
    import csv
    path = os.path.join(ROOT, "profiles", "r02e_gemm_sketch_raw.csv")
    try:
        rows = list(csv.reader(open(path)))
        H, U = rows[0], rows[1]
        ik, ir, iw = H.index("Kernel Name"), H.index("dram__bytes_read.sum"), H.index("dram__bytes_write.sum")
        mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        tot, cnt = 0.0, 0
        for r in rows[2:]:
            if len(r) > max(ir, iw) and "gemm_sketch" in r[ik]:
                tot += float(r[ir].replace(",", "")) * mult[U[ir]] + float(r[iw].replace(",", "")) * mult[U[iw]]
                cnt += 1
        return (tot / cnt, cnt) if cnt else (None, 0)
    except Exception:
        return None, 0


def _timed(ext, fn, steps, warmup, after_warmup=None):
    """CUDA events on the library's own stream around `steps` calls of fn (after `warmup` untimed calls)."""
    import torch
    for w in range(warmup):
        fn(1000 + w)
    if after_warmup is not None:
        after_warmup()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(ext)
    out = None
    for s_ in range(steps):
        out = fn(s_)
    e1.record(ext)
    e1.synchronize()
    return e0.elapsed_time(e1) * 1e-3 / steps, out


def run_c3(ctx, ext, dev, steps=5, n=16384):
    """BASELINE config 3: idfact, sketch=:srft, 16384^2 FP64, geometric spectrum, rtol=1e-12.  HBM-bound sketch."""
    import torch
    from brapprox._binding import DeviceMatrix
    from brapprox._frontend import idfact_device
    g = torch.Generator(device=dev)
    g.manual_seed(3)
    U, _ = torch.linalg.qr(torch.randn(n, RANK_GEN, dtype=torch.float64, device=dev, generator=g))
    V, _ = torch.linalg.qr(torch.randn(n, RANK_GEN, dtype=torch.float64, device=dev, generator=g))
    sig = 10.0 ** (-DECADES * torch.arange(RANK_GEN, dtype=torch.float64, device=dev) / JDIV)
    At = ((V * sig) @ U.T).contiguous()
    del U, V
    A = DeviceMatrix(At.data_ptr(), n, n, n, keep=At)
    torch.cuda.synchronize(dev)
    t, inf = _timed(ext, lambda sd: idfact_device(A, rtol=RTOL, sketch="srft", seed=sd, ctx=ctx), steps, 2,
                    after_warmup=lambda: ctx.profile_enable(True))
    prof = ctx.profile_read()
    ctx.profile_enable(False)
    rounds = [(int(inf.orders[i]), int(inf.ks[i])) for i in range(inf.rounds)]
    nrun = steps
    sk_ms = prof["sketch_other"][0] / nrun
    bytes_sk = sum(8.0 * n * n + 8.0 * l * n for l, _ in rounds)
    peak, src = hbm_peak()
    return {"workload": f"C3: idfact sketch=srft {n}x{n} FP64 rtol=1e-12, A resident", "ms_per_factorization": t * 1e3,
            "factorizations_per_sec": 1.0 / t, "rounds_order_k": rounds, "k": int(inf.k),
            "stage_ms": {k2: v[0] / nrun for k2, v in prof.items() if v[0] > 0},
            "roofline": {"bound": "hbm", "kernel": "srft_kernel (smem radix-4/2 FFT + pruned twiddle stage)",
                         "achieved": bytes_sk / (sk_ms * 1e-3) / 1e9 if sk_ms > 0 else None, "peak": peak, "unit": "GB/s",
                         "frac": bytes_sk / (sk_ms * 1e-3) / 1e9 / peak if sk_ms > 0 else None,
                         "algorithmic_bytes": bytes_sk, "peak_source": src}}


def run_c2c(ctx, ext, dev, A_square, steps=5):
    """The (:left, :c) forms at the headline shape: idfact(:c) on the C2 matrix against idfact(:n),
    and psvdfact of a wide 4096 x 8192 matrix (m < n: the reference factors A', src/psvd.jl:256).  The :c sketch runs on
    the TMA + DMMA kernel through the transposed copy of A cached per factorization."""
    import torch
    from brapprox._binding import DeviceMatrix
    from brapprox._frontend import idfact_device, psvdfact_device
    out = {"workload": "C2c: idfact trans=:c 8192x8192 and psvdfact 4096x8192 FP64 rtol=1e-12 sketch=randn, A resident"}
    for tr in ("n", "c"):
        t, inf = _timed(ext, lambda sd: idfact_device(A_square, rtol=RTOL, trans=tr, seed=sd, ctx=ctx), steps, 2)
        out[f"idfact_{tr}_ms"] = t * 1e3
        out[f"idfact_{tr}_k"] = int(inf.k)
    out["c_over_n"] = out["idfact_c_ms"] / out["idfact_n_ms"]
    m, n = 4096, 8192
    g = torch.Generator(device=dev)
    g.manual_seed(7)
    U, _ = torch.linalg.qr(torch.randn(m, RANK_GEN, dtype=torch.float64, device=dev, generator=g))
    V, _ = torch.linalg.qr(torch.randn(n, RANK_GEN, dtype=torch.float64, device=dev, generator=g))
    sig = 10.0 ** (-DECADES * torch.arange(RANK_GEN, dtype=torch.float64, device=dev) / JDIV)
    Wt = ((V * sig) @ U.T).contiguous()                   # n x m row-major == m x n column-major
    W = DeviceMatrix(Wt.data_ptr(), m, n, m, keep=Wt)
    torch.cuda.synchronize(dev)
    ctx.profile_enable(False)
    t, inf = _timed(ext, lambda sd: psvdfact_device(W, rtol=RTOL, seed=sd, ctx=ctx), steps, 2,
                    after_warmup=lambda: ctx.profile_enable(True))
    prof = ctx.profile_read()
    ctx.profile_enable(False)
    out["psvdfact_4096x8192_ms"] = t * 1e3
    out["psvdfact_4096x8192_k"] = int(inf.k)
    out["psvdfact_4096x8192_stage_ms"] = {k2: v[0] / steps for k2, v in prof.items() if v[0] > 0}
    out["generic_gemm_note"] = "no gemm_generic_kernel launch with a contraction >= 1024 remains on these paths"
    return out


def run_c1(ctx, ext, dev, steps=50, n=1024):
    """BASELINE config 1 (the reference's own CPU-runnable case): Hilbert 1024^2, rtol=1e-15, sketch=:randn -- one round
    of order 40, k ~ 26: pure latency, reported in microseconds (a roofline fraction is not meaningful here)."""
    import torch
    from brapprox._binding import DeviceMatrix
    from brapprox._frontend import idfact_device, pqrfact_device
    i = torch.arange(n, dtype=torch.float64, device=dev)
    At = (1.0 / (i[:, None] + i[None, :] + 1.0)).contiguous()            # symmetric: row-major == column-major
    A = DeviceMatrix(At.data_ptr(), n, n, n, keep=At)
    out = {"workload": f"C1: Hilbert {n}x{n} FP64, rtol=1e-15, sketch=randn, A resident (latency case)"}
    for name, fn in (("idfact", idfact_device), ("pqrfact", pqrfact_device)):
        t, inf = _timed(ext, lambda sd: fn(A, rtol=1e-15, seed=sd, ctx=ctx), steps, 3)
        out[name + "_us"] = t * 1e6
        out[name + "_rounds_order_k"] = [(int(inf.orders[r]), int(inf.ks[r])) for r in range(inf.rounds)]
    return out


def run_c5(ctx, ext, dev, rank, world, nblocks_total=16384, steps=3, m=512):
    """BASELINE config 5: batched idfact of independent 512x512 Cauchy blocks, sketch=:sprn, rtol=1e-12; blocks are
    dealt out in contiguous groups, one group per GPU, no collective (strong scaling: total block count fixed)."""
    import torch
    import brapprox
    from brapprox._frontend import idfact_batched_device
    b0, nb = brapprox.block_shard(nblocks_total, rank, world)
    n = m
    At = torch.empty((nb, n, m), dtype=torch.float64, device=dev)      # block b column-major, lda = m
    g = torch.Generator(device=dev)
    g.manual_seed(50 + rank)
    for c0 in range(0, nb, 1024):
        c1 = min(nb, c0 + 1024)
        x = torch.sort(torch.rand((c1 - c0, m), dtype=torch.float64, device=dev, generator=g), dim=1).values
        y = torch.sort(torch.rand((c1 - c0, n), dtype=torch.float64, device=dev, generator=g), dim=1).values + 1.02
        At[c0:c1] = 1.0 / (x[:, None, :] - y[:, :, None])
    ldt = 32
    kd = torch.zeros(nb, dtype=torch.int64, device=dev)
    pd = torch.zeros((nb, n), dtype=torch.int64, device=dev)
    Td = torch.zeros((nb, n, ldt), dtype=torch.float64, device=dev)
    torch.cuda.synchronize()

    def step(sd):
        return idfact_batched_device(At.data_ptr(), nb, m, n, m, m * n, kd.data_ptr(), pd.data_ptr(), Td.data_ptr(), ldt,
                                     ldt * n, None, ctx=ctx, rtol=RTOL, sketch="sprn", seed=sd)

    t, unfinished = _timed(ext, step, steps, 1)
    ks = kd.cpu().numpy()
    hist = {int(k): int(c) for k, c in zip(*[a.tolist() for a in __import__("numpy").unique(ks, return_counts=True)])}
    peak, src = hbm_peak()
    by = 8.0 * m * n * nb
    return {"workload": f"C5: batched idfact, {nblocks_total} Cauchy 512x512 blocks (this rank: {nb}), sketch=sprn, "
                        "rtol=1e-12, fast mode, blocks resident", "t": t, "blocks": nb,
            "blocks_per_sec_this_rank": nb / t, "k_hist": hist, "unfinished_blocks": int(unfinished),
            "roofline": {"bound": "hbm", "kernel": "idfact_batched_kernel (fused sprn sketch + QRCP + T solve)",
                         "achieved": by / t / 1e9, "peak": peak, "unit": "GB/s", "frac": by / t / 1e9 / peak,
                         "algorithmic_bytes": by, "peak_source": src}}


def run_c4(ctx, ext, dev, rank, world, m_total=1048576, n=4096, steps=2):
    """BASELINE config 4: tall pqrfact, rank=256 cap, row blocks per GPU, sketch all-reduced over NCCL once per round
    (strong scaling: m_total fixed).  Needs ctx's communicator when world > 1."""
    import torch
    import brapprox
    from brapprox._binding import DeviceMatrix
    from brapprox._frontend import pqrfact_device
    row0, ml = brapprox.row_shard(m_total, rank, world)
    r = 512
    g = torch.Generator(device=dev)
    g.manual_seed(4)                                       # Y, sigma identical on every rank
    Y, _ = torch.linalg.qr(torch.randn(n, r, dtype=torch.float64, device=dev, generator=g))
    sig = 10.0 ** (-6.0 * torch.arange(r, dtype=torch.float64, device=dev) / 256.0)
    Ys = (Y * sig).T.contiguous()                          # r x n
    At = torch.empty((n, ml), dtype=torch.float64, device=dev)          # column-major ml x n
    g2 = torch.Generator(device=dev)
    g2.manual_seed(1000 + rank)
    for c0 in range(0, ml, 65536):
        c1 = min(ml, c0 + 65536)
        X = torch.randn((c1 - c0, r), dtype=torch.float64, device=dev, generator=g2) / (m_total ** 0.5)
        At[:, c0:c1] = (X @ Ys).T
    del Y, Ys
    A = DeviceMatrix(At.data_ptr(), ml, n, ml, keep=At)
    torch.cuda.synchronize(dev)
    ctx.set_row_shard(row0, m_total)
    t, inf = _timed(ext, lambda sd: pqrfact_device(A, rtol=RTOL, rank=256, seed=sd, ctx=ctx), steps, 1,
                    after_warmup=lambda: ctx.profile_enable(True))
    prof = ctx.profile_read()
    ctx.profile_enable(False)
    rows_exec = int(brapprox.lib.bra_debug_sketch_rows(ctx.handle))
    # the non-adaptive single-sketch form SURVEY 8(d) also asks for: sketchfact_adap = false, one sketch of order 264
    t1, inf1 = _timed(ext, lambda sd: pqrfact_device(A, rtol=RTOL, rank=256, sketchfact_adap=False, seed=sd, ctx=ctx), steps, 1)
    ctx.set_row_shard(0, 0)
    rounds = [(int(inf.orders[i]), int(inf.ks[i])) for i in range(inf.rounds)]
    f_sk = 2.0 * m_total * n * sum(l for l, _ in rounds)
    k = int(inf.k)
    # F_C = 2mk^2 - 2/3 k^3 (QR of C) + 2mk^2 (forming Q) + k^2(n-k) (R = R1 [I T]) -- the ALGORITHMIC count, not
    # the 8mk^2 this implementation's CholeskyQR2 really spends
    f_tail = 4.0 * m_total * k * k - (2.0 / 3.0) * k ** 3 + float(k) * k * (n - k)
    nrun = steps
    f1 = 2.0 * m_total * n * int(inf1.orders[0]) + f_tail
    return {"workload": f"C4: pqrfact {m_total}x{n} FP64 rank=256 cap, sketch=randn adaptive, rows sharded over "
                        f"{world} GPU(s) ({ml} rows here), one sketch all-reduce per round", "t": t,
            "rounds_order_k": rounds, "k": k, "algorithmic_gflop_total": (f_sk + f_tail) / 1e9,
            "sketch_rows": {"reference_schedule": sum(l for l, _ in rounds), "executed": rows_exec},
            "executed_gflop_total": (2.0 * m_total * n * rows_exec + f_tail) / 1e9,
            "stage_ms": {k2: v[0] / nrun for k2, v in prof.items() if v[0] > 0},
            "collectives_per_factorization": None,
            "nonadaptive_l264": {"t": t1, "order": int(inf1.orders[0]), "k": int(inf1.k), "algorithmic_gflop_total": f1 / 1e9}}


def run_c4_parity(ctx, dev, rank, world, m=131072, n=4096):
    """Parity of the row-sharded path at this world size: a reduced-m C4 matrix (identical on every rank) factored over
    the ranks' row blocks against the SAME library on one GPU without a communicator, same fast-mode seed (Omega is keyed
    by the global row index, so only the summation order of the all-reduce differs)."""
    import torch
    import brapprox
    from brapprox import _binding as B
    from brapprox._binding import DeviceMatrix
    from brapprox._frontend import pqrfact_device
    r = 512
    g = torch.Generator(device=dev)
    g.manual_seed(44)
    Y, _ = torch.linalg.qr(torch.randn(n, r, dtype=torch.float64, device=dev, generator=g))
    sig = 10.0 ** (-6.0 * torch.arange(r, dtype=torch.float64, device=dev) / 256.0)
    X = torch.randn((m, r), dtype=torch.float64, device=dev, generator=g) / (m ** 0.5)
    At = ((X * sig) @ Y.T).T.contiguous()                  # n x m row-major == m x n column-major, identical on all ranks
    del X, Y
    row0, ml = brapprox.row_shard(m, rank, world)
    Aloc_t = At[:, row0:row0 + ml].contiguous()
    Aloc = DeviceMatrix(Aloc_t.data_ptr(), ml, n, ml, keep=Aloc_t)
    torch.cuda.synchronize(dev)        # the library runs on its own stream: torch must have finished writing the inputs
    ctx.set_row_shard(row0, m)
    inf = pqrfact_device(Aloc, rtol=RTOL, rank=256, seed=5, ctx=ctx)
    k = int(inf.k)
    p = ctx.fetch(B.F_P, (n,), "int64")
    R = ctx.fetch(B.F_R, (k, n))
    ctx.set_row_shard(0, 0)
    out = None
    if rank == 0:
        ref = brapprox.Context(dev.index)
        Afull = DeviceMatrix(At.data_ptr(), m, n, m, keep=At)
        inf1 = pqrfact_device(Afull, rtol=RTOL, rank=256, seed=5, ctx=ref)
        p1 = ref.fetch(B.F_P, (n,), "int64")
        R1 = ref.fetch(B.F_R, (int(inf1.k), n))
        same_k = int(inf1.k) == k
        out = {"workload": f"pqrfact {m}x{n} rank=256 over {world} row blocks vs 1 GPU, same seed",
               "k": k, "k_equal": same_k, "p_equal": bool(same_k and (p[:k] == p1[:k]).all()),
               "rounds_equal": [int(inf.orders[i]) for i in range(inf.rounds)] == [int(inf1.orders[i]) for i in range(inf1.rounds)],
               "max_abs_R_diff_over_R11": float(abs(R - R1).max() / abs(R1[0, 0])) if same_k else None}
        ref.close()
    return out


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
        return

    import numpy as np
    import torch
    import brapprox
    from brapprox import _binding as B
    from brapprox._binding import DeviceMatrix
    from brapprox._frontend import idfact_device

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    ctx = brapprox.Context(local)
    has_psvd = hasattr(brapprox, "psvdfact_device")
    what = args.what if args.what != "auto" else ("psvdfact" if has_psvd else "idfact")

    n = args.n
    g = torch.Generator(device=dev)
    g.manual_seed(1 + rank)
    U, _ = torch.linalg.qr(torch.randn(n, RANK_GEN, dtype=torch.float64, device=dev, generator=g))
    V, _ = torch.linalg.qr(torch.randn(n, RANK_GEN, dtype=torch.float64, device=dev, generator=g))
    sig = 10.0 ** (-DECADES * torch.arange(RANK_GEN, dtype=torch.float64, device=dev) / JDIV)
    At = ((V * sig) @ U.T).contiguous()         # row-major (n x n)  ==  column-major A = U diag(sig) V^T
    A = DeviceMatrix(At.data_ptr(), n, n, n, keep=At)
    del U, V
    torch.cuda.synchronize()

    if what == "psvdfact":
        from brapprox._frontend import psvdfact_device as fact_device
    else:
        fact_device = idfact_device

    def step(seed):
        return fact_device(A, rtol=RTOL, seed=seed, ctx=ctx)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- FP64 peak measured in this run (MEASURED_PEAKS.json carries no FP64 figure) ----
    peaks = brapprox.probe_fp64_peak(ctx)
    x = torch.randn(8192, 8192, dtype=torch.float64, device=dev)
    y = torch.randn(8192, 8192, dtype=torch.float64, device=dev)
    for _ in range(2):
        torch.matmul(x, y)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(3):
        e0.record()
        torch.matmul(x, y)
        e1.record()
        e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    cublas_tf = 2.0 * 8192 ** 3 / (best * 1e-3) / 1e12
    del x, y
    fp64_peak = max(cublas_tf, max(peaks.values()))

    # ---- warm-up, then the timed region (A = 537 MB > 126 MB L2: inputs larger than L2) ----
    for w in range(max(args.warmup, 3)):
        inf = step(100 + w)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ctx.profile_enable(True)
    launches0 = ctx.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    # the library launches on its own stream: record the CUDA events THERE (torch's current stream sees nothing)
    ext = torch.cuda.ExternalStream(int(B.lib.bra_stream(ctx.handle)), device=dev)
    t0 = time.perf_counter()
    ev0.record(ext)
    for s in range(args.steps):
        inf = step(s)
    ev1.record(ext)
    ev1.synchronize()
    barrier()
    wall_host = time.perf_counter() - t0
    wall = ev0.elapsed_time(ev1) * 1e-3
    prof = ctx.profile_read()
    ctx.profile_enable(False)
    launches = ctx.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    tmax = wall
    if world > 1:
        t = torch.tensor([wall], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        tmax = float(t.item())
    value = world * args.steps / tmax

    rounds = [(int(inf.orders[t]), int(inf.ks[t])) for t in range(inf.rounds)]
    steps = [int(inf.steps[t]) for t in range(inf.rounds)]
    k = int(inf.k)
    f_sk, f_qr, f_t = algorithmic_flops(n, n, rounds, steps, k)
    f_total = f_sk + f_qr + f_t + (psvd_extra_flops(n, n, k) if what == "psvdfact" else 0.0)
    gemm_ms, gemm_calls = prof["gemm"]
    # nested sketches: the library multiplies max(order) rows of Omega, the reference schedule sum(order)
    rows_exec = int(brapprox.lib.bra_debug_sketch_rows(ctx.handle))
    rows_ref = sum(l for l, _ in rounds)
    f_sk_exec = 2.0 * n * n * rows_exec
    f_exec = f_total - f_sk + f_sk_exec
    gemm_tf = f_sk_exec * args.steps / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else None

    # ---- the other half of the headline metric ("idfact/psvdfact factorizations/sec"): idfact on the same A ----
    idf = None
    if what == "psvdfact":
        ctx.profile_enable(True)
        ti, infi = _timed(ext, lambda sd: idfact_device(A, rtol=RTOL, seed=sd, ctx=ctx), max(5, args.steps // 2), 2)
        profi = ctx.profile_read()
        ctx.profile_enable(False)
        rows_i = int(brapprox.lib.bra_debug_sketch_rows(ctx.handle))
        ri = [(int(infi.orders[t]), int(infi.ks[t])) for t in range(infi.rounds)]
        si = [int(infi.steps[t]) for t in range(infi.rounds)]
        fi = sum(algorithmic_flops(n, n, ri, si, int(infi.k)))
        nrun = max(5, args.steps // 2) + 2
        idf = {"value": 1.0 / ti, "unit": "factorizations/s", "ms_per_step": ti * 1e3, "k": int(infi.k),
               "algorithmic_gflop": fi / 1e9, "achieved_tflops": fi / ti / 1e12,
               "frac_of_fp64_peak": fi / ti / 1e12 / fp64_peak,
               "sketch_rows": {"reference_schedule": sum(l for l, _ in ri), "executed": rows_i},
               "executed_gflop": (fi - 2.0 * n * n * (sum(l for l, _ in ri) - rows_i)) / 1e9,
               "frac_of_fp64_peak_executed": (fi - 2.0 * n * n * (sum(l for l, _ in ri) - rows_i)) / ti / 1e12 / fp64_peak,
               "stage_ms_per_step": {k2: v[0] / nrun for k2, v in profi.items() if v[0] > 0}}

    # ---- e2e: host-resident A through the C ABI, result fetched to the host, every step ----
    e2e = None
    if rank == 0 or world > 1:
        Ah = torch.empty((n, n), dtype=torch.float64, pin_memory=True)
        Ah.copy_(At)
        Ahn = Ah.numpy().T            # column-major view of the pinned buffer (F-contiguous)
        from brapprox._frontend import idfact, _rounds
        e2e_steps = max(3, min(args.steps, 10))

        # results land in caller-owned pinned buffers (the ABI's ownership model; sized for k <= 640): both directions
        # of the end-to-end step then run at the PCIe rate and the step allocates nothing
        kmax = RANK_GEN
        Ub = torch.empty((kmax, n), dtype=torch.float64, pin_memory=True).numpy().T        # n x kmax, column-major
        Sb = torch.empty((kmax,), dtype=torch.float64, pin_memory=True).numpy()
        Vb = torch.empty((n, kmax), dtype=torch.float64, pin_memory=True).numpy().T        # kmax x n, column-major

        def e2e_step(seed):
            if what == "psvdfact":
                F = brapprox.psvdfact(Ahn, rtol=RTOL, seed=seed, ctx=ctx, out=(Ub, Sb, Vb))
                return F.U.nbytes + F.S.nbytes + F.Vt.nbytes
            Vv = idfact(Ahn, rtol=RTOL, seed=seed, ctx=ctx)
            return Vv.sk.nbytes + Vv.rd.nbytes + Vv.T.nbytes

        e2e_step(7)
        barrier()
        t0 = time.perf_counter()
        for s in range(e2e_steps):
            d2h = e2e_step(s)
        barrier()
        te = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([te], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            te = float(t.item())
        e2e = {"value": world * e2e_steps / te, "unit": "factorizations/s", "h2d_bytes_per_step": int(n * n * 8),
               "d2h_bytes_per_step": int(d2h), "steps": e2e_steps}
    # ---- the error metric of BASELINE.json for OUR factorization: snormdiff(A, F) / snorm(A), on the device
    #      (bra_snorm_f64: the reference's power iteration, src/snorm.jl:14-53) ----
    our_err = None
    if rank == 0:
        try:
            if what == "psvdfact":
                Fh = brapprox.psvdfact(A, rtol=RTOL, seed=12345, ctx=ctx)
                our_err = brapprox.snormdiff(A, Fh, ctx=ctx) / brapprox.snorm(A, ctx=ctx)
            else:
                from brapprox._frontend import idfact as _idf
                Vh = _idf(A, rtol=RTOL, seed=12345, ctx=ctx)
                Cs = At[torch.as_tensor(Vh.sk - 1, device=dev), :].T      # A[:, sk] (At holds A^T row by row)
                our_err = (brapprox.snormdiff(A, Cs.contiguous().cpu().numpy(), Vh.matrix(), ctx=ctx)
                           / brapprox.snorm(A, ctx=ctx))
        except Exception as e:
            our_err = repr(e)[:200]

    # ---- side measurements of the other BASELINE configs (same run, same box): C3 single GPU; C4 (row-sharded,
    # NCCL all-reduce) and C5 (batched, no collective) over all ranks, strong scaling, max over ranks ----
    extra = {}
    if not args.no_extra:
        try:
            if rank == 0:
                extra["C2c"] = run_c2c(ctx, ext, dev, A)
        except Exception as e:      # a side measurement must never take the headline down
            extra["C2c"] = {"error": repr(e)[:300]}
        del A, At
        torch.cuda.empty_cache()
        fp64_tf = None

        def over_ranks(val, op):
            if world == 1:
                return val
            tt = torch.tensor([val], dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=op)
            return float(tt.item())

        try:
            if rank == 0:
                extra["C1"] = run_c1(ctx, ext, dev)
        except Exception as e:      # a side measurement must never take the headline down
            extra["C1"] = {"error": repr(e)[:300]}
        try:
            if rank == 0:
                extra["C3"] = run_c3(ctx, ext, dev)
            torch.cuda.empty_cache()
        except Exception as e:      # a side measurement must never take the headline down
            extra["C3"] = {"error": repr(e)[:300]}
        try:
            barrier()
            c5 = run_c5(ctx, ext, dev, rank, world, args.c5_blocks)
            t5 = over_ranks(c5.pop("t"), dist.ReduceOp.MAX if world > 1 else None)
            c5["blocks_per_sec_all_ranks"] = args.c5_blocks / t5
            c5["ms_per_batch"] = t5 * 1e3
            c5["n_gpus"] = world
            c5["scaling"] = "strong"
            extra["C5"] = c5
            torch.cuda.empty_cache()
        except Exception as e:
            extra["C5"] = {"error": repr(e)[:300]}
        try:
            barrier()
            if world > 1:
                brapprox.init_comm(ctx, rank, world, device=dev)
            c0 = ctx.collective_count()
            c4 = run_c4(ctx, ext, dev, rank, world, args.c4_rows)
            t4 = over_ranks(c4.pop("t"), dist.ReduceOp.MAX if world > 1 else None)
            c4["collectives_per_factorization"] = (ctx.collective_count() - c0) / 3.0      # 1 warm-up + 2 timed
            c4["ms_per_factorization"] = t4 * 1e3
            c4["factorizations_per_sec"] = 1.0 / t4
            c4["achieved_tflops_all_ranks"] = c4["algorithmic_gflop_total"] / 1e3 / t4
            c4["frac_of_fp64_peak_per_gpu"] = c4["achieved_tflops_all_ranks"] / world / fp64_peak
            # the flops really executed (nested sketches multiply max(order) rows, not the reference schedule's sum):
            # THIS is the hardware utilisation; the figure above is the speed in units of the reference's work
            c4["executed_tflops_all_ranks"] = c4["executed_gflop_total"] / 1e3 / t4
            c4["frac_of_fp64_peak_per_gpu_executed"] = c4["executed_tflops_all_ranks"] / world / fp64_peak
            c4["n_gpus"] = world
            c4["scaling"] = "strong"
            na = c4["nonadaptive_l264"]
            tna = over_ranks(na.pop("t"), dist.ReduceOp.MAX if world > 1 else None)
            na["ms_per_factorization"] = tna * 1e3
            na["frac_of_fp64_peak_per_gpu"] = na["algorithmic_gflop_total"] / 1e3 / tna / world / fp64_peak
            if world > 1:
                c4["parity_vs_1gpu"] = run_c4_parity(ctx, dev, rank, world)
            extra["C4"] = c4
            torch.cuda.empty_cache()
        except Exception as e:
            extra["C4"] = {"error": repr(e)[:300]}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cpu, _, _ = cpu_sample(n, what)
        one, _, _ = cpu_sample(n, what, max_seconds=12.0, threads=1)
        cpu["one_thread"] = {"value": one["value"], "unit": one["unit"], "cores": 1, "sample": one["sample"]}

    line = {
        "metric": f"{what}_factorizations_per_sec", "value": value, "unit": "factorizations/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": 1e3 * tmax / args.steps,
        "host_wall_ms_per_step": 1e3 * wall_host / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_label(what, n),
                   "arm": "libbrapprox on B200: device Philox Omega, A resident in HBM for `value`, host A for `e2e`",
                   "rounds_order_k": rounds, "pivot_steps": steps, "k": k,
                   "l2": "inputs (537 MB) larger than L2 (126 MB); no flush needed",
                   "parallelism": "replicas only" if world > 1 else "single GPU"},
        "whole_factorization": {"algorithmic_gflop": f_total / 1e9,
                                "achieved_tflops": f_total * world * args.steps / tmax / 1e12 / world,
                                "frac_of_fp64_peak": f_total * args.steps / tmax / 1e12 / fp64_peak},
        # config-level roofline: ALL algorithmic flops of the factorization over the whole step time; the
        # dominant kernels follow in `kernels` (the FP64 tensor GEMM against the same peak; the pivoted-QR chain is
        # latency-bound: reported in microseconds per dependent pivot step)
        "roofline": {"bound": "tensor", "scope": f"whole {what} (sketch GEMMs + pivoted QR + T solve"
                                                 + (" + QR of A[:,sk] + psvd core + U, Vt)" if what == "psvdfact" else ")"),
                     "achieved": f_total * args.steps / tmax / 1e12, "peak": fp64_peak, "unit": "TFLOP/s",
                     "frac": f_total * args.steps / tmax / 1e12 / fp64_peak, "traffic": None,
                     "algorithmic_flops_per_step": f_total,
                     "flop_count": "SURVEY 8(d): the reference schedule, F_sk = 2 m n sum(l_t) over the adaptive rounds",
                     "executed": {"sketch_rows": rows_exec, "sketch_rows_reference_schedule": rows_ref,
                                  "flops_per_step": f_exec, "achieved": f_exec * args.steps / tmax / 1e12,
                                  "frac": f_exec * args.steps / tmax / 1e12 / fp64_peak,
                                  "note": "nested sketches: round t's Omega is round t-1's plus fresh rows, so only the new "
                                          "rows are multiplied with A (max instead of sum of the orders); `frac` above counts "
                                          "the reference schedule's flops, this entry the flops really executed"},
                     "peak_source": f"measured in this run: cuBLAS DGEMM 8192^3 = {cublas_tf:.1f} TF, "
                                    f"DMMA issue probes = {max(peaks.values()):.1f} TF "
                                    "(MEASURED_PEAKS.json has no FP64 figure)",
                     "kernels": [
                         {"kernel": "gemm_sketch_kernel (TMA + DMMA m8n8k4 FP64)", "bound": "tensor",
                          "achieved": gemm_tf, "peak": fp64_peak, "unit": "TFLOP/s",
                          "frac": (gemm_tf / fp64_peak) if gemm_tf else None,
                          "share_of_step": gemm_ms / args.steps / (1e3 * tmax / args.steps),
                          "traffic": ncu_traffic_per_launch()[0],
                          "traffic_note": "DRAM bytes per launch, mean over the 5 sketch launches of one factorization "
                                          "(ncu --set full, profiles/r02e_gemm_sketch_raw.csv: 0.54 GB for the launches of <= 64 new rows, "
                                          "0.92 / 1.41 GB for 128 / 256 rows); algorithmic: 537 MB of A + rows x 8192 x 16 B",
                          "algorithmic_flops_per_step": f_sk_exec, "launches": gemm_calls, "ms_total": gemm_ms},
                         {"kernel": "qrcp_fast_kernel (persistent warp-specialised pivoted QR)", "bound": "latency",
                          "us_per_pivot_step": 1e3 * prof["qrcp"][0] / args.steps / max(1, sum(steps)),
                          "pivot_steps": sum(steps), "algorithmic_flops_per_step": f_qr,
                          "achieved": f_qr * args.steps / (prof["qrcp"][0] * 1e-3) / 1e12 if prof["qrcp"][0] > 0 else None,
                          "unit": "TFLOP/s", "share_of_step": prof["qrcp"][0] / args.steps / (1e3 * tmax / args.steps)}]},
        "error": {"metric": "snormdiff(A, F) / snorm(A)", "ours": our_err,
                  "oracle": (cpu or {}).get("snormdiff_over_snorm"),
                  "note": "ours: bra_snorm_f64 on the device for a fresh factorization of the benchmark matrix; oracle: "
                          "the same metric for the CPU arm's own instance of the workload (same spectrum, numpy seed)"},
        "stage_ms_per_step": {k2: v[0] / args.steps for k2, v in prof.items()},
        "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        "fp64_peak_probes_tflops": peaks,
        "idfact_same_matrix": idf,
        "other_configs": extra,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
