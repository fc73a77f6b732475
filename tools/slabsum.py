"""One-line summary of a slab bench JSON line (bench.py --gpus N>1), for gpurun A/B scripts."""
import json
import sys

for path in sys.argv[1:]:
    d = None
    for line in open(path):
        if line.startswith("{"):
            d = json.loads(line)
    if d is None:
        print(path, "no JSON line")
        continue
    ck = d["config"].get("check_vs_single_gpu") or {}
    print(f"{path}: N={d['n_gpus']} ms/step={d['ms_per_step']:.3f} value={d['value']:.3e} eff={d.get('parallel_efficiency')} "
          f"t1={(d.get('t1') or {}).get('ms_per_step')} check={ck.get('rel_l2_slab_vs_single_gpu')} ok={ck.get('ok')} "
          f"row={d['roofline']['kernels']['row_ms_per_step']:.2f} col={d['roofline']['kernels']['col_ms_per_step']:.2f} "
          f"a2a={d['nvlink']['a2a_ms_per_step_on_comm_stream']:.2f} e2e_ms={d['e2e']['ms_per_step']:.1f} finite={d['config']['state_finite']}")
    print("   ", d["config"]["workload"])
