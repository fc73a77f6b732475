"""Summarise bench.py JSON lines read from stdin: one short line each (used in gpurun A/B scripts)."""
import json
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else ""
for line in sys.stdin:
    line = line.strip()
    if not line.startswith("{"):
        continue
    d = json.loads(line)
    if "roofline" not in d:
        print(tag, json.dumps(d)[:300])
        continue
    k = d["roofline"]["kernels"]
    print(f"{tag} grid={d['config'].get('grid')} ms/step={d['ms_per_step']:.3f} value={d['value']:.3e} "
          f"step_frac={d['roofline']['whole_step']['frac']:.3f} "
          f"col={k['col_step']['avg_ms']:.3f}ms({k['col_step']['frac']:.3f}) "
          f"row={k['row_jac']['avg_ms']:.3f}ms({k['row_jac']['frac']:.3f}) e2e={d['e2e']['value']:.3e} "
          f"finite={d['config'].get('state_finite')} clk={d['clocks']['sm_mhz']} {d['clocks']['reasons']}")
