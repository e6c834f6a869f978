"""Summarise `ncu --page raw --csv` exports (one row per captured launch) into a small table.
    python tools/ncu_summary.py profiles/r01b_*_raw.csv > profiles/r01b_summary.md"""
import csv
import sys

COLS = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("FBSP.TriageCompute.dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram_%"),
        ("sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active", "dmma_pipe_%"),
        ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64_pipe_%"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_%"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block")]


def fmt(v, u):
    try:
        x = float(v.replace(",", ""))
    except Exception:
        return v
    if u in ("byte", "Kbyte", "Mbyte", "Gbyte"):
        x *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
        return f"{x / 1e6:.1f} MB"
    if u in ("ns", "nsecond", "us", "usecond", "ms", "msecond", "s", "second"):
        x *= {"n": 1e-3, "u": 1.0, "m": 1e3, "s": 1e6}[u[0]]
        return f"{x:.1f} us"
    return f"{x:.1f}" if x != int(x) else str(int(x))


print("| file | kernel | " + " | ".join(c[1] for c in COLS) + " | HBM GB/s |")
print("|---|---|" + "---|" * (len(COLS) + 1))
for path in sys.argv[1:]:
    rows = list(csv.reader(open(path)))
    if len(rows) < 3:
        continue
    H, U = rows[0], rows[1]
    ik = H.index("Kernel Name")
    for r in rows[2:]:
        if len(r) < len(H):
            continue
        cells = []
        t_us = rd = wr = None
        for name, short in COLS:
            if name not in H:
                cells.append("-")
                continue
            i = H.index(name)
            cells.append(fmt(r[i], U[i]))
            try:
                x = float(r[i].replace(",", ""))
                if short == "time":
                    t_us = x * {"n": 1e-3, "u": 1.0, "m": 1e3, "s": 1e6}[U[i][0]]
                if short in ("dram_rd", "dram_wr"):
                    b = x * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[U[i]]
                    if short == "dram_rd":
                        rd = b
                    else:
                        wr = b
            except Exception:
                pass
        bw = f"{(rd + wr) / t_us / 1e3:.0f}" if t_us and rd is not None and wr is not None else "-"
        kname = r[ik].replace("(anonymous namespace)::", "").replace("<unnamed>::", "").replace("void ", "").split("(")[0]
        print(f"| {path.split('/')[-1].replace('_raw.csv', '')} | {kname} | " + " | ".join(cells) + f" | {bw} |")
