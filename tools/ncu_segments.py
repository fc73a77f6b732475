"""Per-segment stall attribution: splits a kernel's SASS at BAR.SYNC and sums the ncu source-page samples.
usage: python tools/ncu_segments.py report.ncu-rep kernel_regex [instance]"""
import collections
import csv
import subprocess
import sys

rep, rx = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{rx}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
sections = [k for k, r in enumerate(rows) if r and r[0] == "Kernel Name"]
sec = int(sys.argv[3]) if len(sys.argv) > 3 else 0
lo = sections[sec]
hi = sections[sec + 1] if sec + 1 < len(sections) else len(rows)
print(rows[lo][1])
hdr = rows[lo + 1]
isrc, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]


def new():
    return {"samples": 0, "sass": 0, "ex": 0, "ops": collections.Counter(), "st": collections.Counter()}


segs, cur, tot = [], new(), 0
for r in rows[lo + 2:hi]:
    if len(r) < len(hdr):
        continue
    try:
        s, ex = int(r[isamp]), int(r[iex])
    except ValueError:
        continue
    parts = r[isrc].split()
    op = parts[0] if parts else ""
    if op.startswith("@") and len(parts) > 1:
        op = parts[1]
    cur["samples"] += s
    cur["sass"] += 1
    cur["ex"] += ex
    cur["ops"][op.split(".")[0]] += ex
    for i in stall_cols:
        try:
            cur["st"][hdr[i]] += int(r[i])
        except ValueError:
            pass
    tot += s
    if op.startswith("BAR"):
        segs.append(cur)
        cur = new()
segs.append(cur)
print("total samples", tot)
for k, v in enumerate(segs):
    top = ", ".join(f"{o[6:]}:{c}" for o, c in v["st"].most_common(4))
    mem = ", ".join(f"{o}:{v['ops'][o]}" for o in ("LDG", "STG", "LDL", "STL", "LDS", "STS") if v["ops"][o])
    print(f"seg {k:3d} sass={v['sass']:5d} warp_ex={v['ex']:9d} samples={v['samples']:6d} ({100 * v['samples'] / max(tot, 1):4.1f}%) | {top} | {mem}")
