"""Hottest SASS instructions of a kernel in an .ncu-rep (source page): python tools/ncu_hot.py report.ncu-rep kernel_regex [N]"""
import csv
import subprocess
import sys

rep, rx = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{rx}"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
lo = [k for k, r in enumerate(rows) if r and r[0] == "Kernel Name"][0]
hdr = rows[lo + 1]
isrc, isamp = hdr.index("Source"), hdr.index("# Samples")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
items = []
for idx, r in enumerate(rows[lo + 2:]):
    if len(r) < len(hdr):
        continue
    try:
        s = int(r[isamp])
    except ValueError:
        continue
    st = sorted(((int(r[i]) if r[i].isdigit() else 0, hdr[i][6:]) for i in stall_cols), reverse=True)[:2]
    items.append((s, idx, r[isrc][:70], st))
tot = sum(i[0] for i in items)
for s, idx, src, st in sorted(items, reverse=True)[:n]:
    print(f"{100 * s / tot:5.2f}% #{idx:5d} {src:70s} {st}")
