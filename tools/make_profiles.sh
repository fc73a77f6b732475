#!/bin/bash
# turn gpurun_out/{prof_G.ncu-rep,launches_G.csv} into the tracked summaries under profiles/
R=${2:-r01}
G=${1:-8192}
python tools/ncu_summary.py gpurun_out/prof_$G.ncu-rep > profiles/${R}_${G}_ncu_full_summary.txt
{ python tools/ncu_segments.py gpurun_out/prof_$G.ncu-rep col_kernel; python tools/ncu_segments.py gpurun_out/prof_$G.ncu-rep row_kernel; } > profiles/${R}_${G}_stall_segments.txt
grep -v "^==" gpurun_out/launches_$G.csv > profiles/${R}_${G}_launches.csv
python - <<PY
import csv, collections
rows = [r for r in csv.reader(open("profiles/${R}_${G}_launches.csv")) if len(r) > 5]
hdr = rows[0]; k = hdr.index("Kernel Name"); v = hdr.index("Metric Value")
tot = collections.Counter(); cnt = collections.Counter()
for r in rows[1:]:
    try: t = float(r[v].replace(",", ""))
    except ValueError: continue
    name = r[k].split("(")[0]; tot[name] += t; cnt[name] += 1
s = sum(tot.values())
with open("profiles/${R}_${G}_launch_shares.txt", "w") as f:
    f.write("kernel share of summed gpu__time_duration (ncu, cold cache, serialised): shares only\n")
    for n, t in tot.most_common():
        f.write(f"{100*t/s:6.2f}%  launches={cnt[n]:4d}  avg={t/cnt[n]/1e3:9.1f} us  {n}\n")
print(open("profiles/${R}_${G}_launch_shares.txt").read())
PY
