"""Per-kernel counts of the SASS mnemonics that prove the tensor / TMA / TMEM paths, from the in-tree library:
    python tools/sass_summary.py > profiles/r02_sass_summary.txt
(cuobjdump -sass on lowrankapprox.jl_b200/brapprox/libbrapprox.so; sm_100a only.)"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "lowrankapprox.jl_b200", "brapprox", "libbrapprox.so")
PAT = {"DMMA": r"\bDMMA\.", "DFMA": r"\bDFMA\b", "UTMALDG": r"\bUTMALDG", "UBLKCP": r"\bUBLKCP", "SYNCS": r"\bSYNCS\.",
       "LDTM": r"\bLDTM", "STTM": r"\bSTTM", "UTCATOMSWS/TMEM alloc": r"\bUTCATOMSWS|\bUTCBAR", "REDUX": r"\bREDUX",
       "LDGSTS": r"\bLDGSTS", "BAR": r"\bBAR\.", "MUFU.RSQ64H": r"MUFU\.RSQ64H"}
out = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout
arch = sorted(set(re.findall(r"arch = (sm_\w+)", out)))
counts = collections.OrderedDict()
cur = None
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(anonymous namespace\)::|braq::", "", name).split("(")[0].replace("void ", "")
        cur = counts.setdefault(name, collections.Counter())
        continue
    if cur is None:
        continue
    for k, p in PAT.items():
        if re.search(p, line):
            cur[k] += 1
print(f"# SASS mnemonic counts per kernel of {os.path.relpath(SO, ROOT)} (architectures in the fatbin: {', '.join(arch)})")
print("# DMMA = FP64 tensor-core MMA (FP64 has no tcgen05 kind); UTMALDG = TMA tiled load; UBLKCP = bulk (1-D TMA) copy;")
print("# SYNCS = mbarrier ops; LDTM / STTM = tcgen05.ld / tcgen05.st (tensor memory); LDGSTS = cp.async")
keys = list(PAT)
print("| kernel | " + " | ".join(keys) + " |")
print("|---|" + "---|" * len(keys))
for name, c in counts.items():
    if any(c[k] for k in keys if k not in ("BAR", "DFMA")) or "kernel" in name:
        print(f"| {name} | " + " | ".join(str(c[k]) if c[k] else "" for k in keys) + " |")
