"""Per-CUDA-line summary of an `ncu --page source --csv --print-source cuda,sass` export:
   python tools/ncu_lines.py src.csv steps ctas [lo hi]   -> warp-instructions per CTA-step and stall samples per line."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
steps, ctas = int(sys.argv[2]), int(sys.argv[3])
lo, hi = (int(sys.argv[4]), int(sys.argv[5])) if len(sys.argv) > 5 else (0, 10 ** 9)
hdr = rows[2]
ix = {}
for i, h in enumerate(hdr): ix.setdefault(h, i)
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
tot = 0; out = []
for r in rows[3:]:
    if len(r) < len(hdr) or r[0] == '': continue
    try: s = int(r[ix['# Samples']]); ie = int(r[ix['Instructions Executed']]); line = int(r[0])
    except ValueError: continue
    tot += s
    st = sorted(((int(r[ix[h]]), h[6:]) for h in stalls if r[ix[h]] not in ('', '0', '-')), reverse=True)[:3]
    out.append((line, s, ie, r[1].strip(), st))
print("total samples", tot)
for line, s, ie, txt, st in out:
    if lo <= line <= hi and (ie or s):
        print(f"L{line:5d} {100*s/tot:5.1f}% ie/cta-step {ie/ctas/steps:7.1f}  {txt[:70]:70s} {st}")
