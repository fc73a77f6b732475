"""BASELINE.json configs[2]: elliptic vortex at 4096^2 with filamentation time, deformation factor and the
effective-diffusivity histograms taken EVERY step.  Prints one JSON line (CUDA-event times on the library's stream).
  python tools/diag_bench.py [--grid 4096] [--steps 20]"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", type=int, default=4096)
    ap.add_argument("--steps", type=int, default=20)
    a = ap.parse_args()
    import torch
    import fields
    import xlab_fftbarotropic_b200 as xfb
    import ctypes as C
    n = a.grid
    dt = 1.0 if n >= 4096 else 3.0
    b = xfb.Backend(n)
    v0 = fields.elliptic(n)
    b.set_vorticity(v0)
    stream = torch.cuda.ExternalStream(b.stream)
    tfil = torch.empty((n, n), dtype=torch.float32, device="cuda")
    deform = torch.empty_like(tfil)
    cmin, cmax = float(v0.min()) - 1e-6, float(v0.max()) * 1.01

    def diag_step():
        b.step(1, dt)
        b._ck(b._L.xfb_get_diagnostics(b._h, 0, C.c_void_p(tfil.data_ptr()), C.c_void_p(deform.data_ptr())))
        return b.keff_hist(64, cmin, cmax)

    for _ in range(3):
        diag_step()
    b.sync()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    with torch.cuda.stream(stream):
        e[0].record(stream)
        for _ in range(a.steps):
            area, g2 = diag_step()
        e[1].record(stream)
    torch.cuda.synchronize()
    ms_diag = e[0].elapsed_time(e[1]) / a.steps
    with torch.cuda.stream(stream):
        b.step(2, dt)
        e[0].record(stream)
        b.step(a.steps, dt)
        e[1].record(stream)
    torch.cuda.synchronize()
    ms_plain = e[0].elapsed_time(e[1]) / a.steps
    print(json.dumps({
        "workload": f"elliptic {n}^2, dt={dt}, diagnostics every step (tfil + deform on device, 64-bin k_eff histograms to host)",
        "ms_per_step_with_diagnostics": ms_diag, "ms_per_step_plain": ms_plain,
        "grid_pt_steps_per_s_with_diagnostics": n * n / (ms_diag * 1e-3),
        "area_sum_over_domain": float(area.sum()) / 600000.0 ** 2, "finite": bool(np.isfinite(g2).all() and torch.isfinite(deform).all().item()),
    }))
    b.close()


if __name__ == "__main__":
    main()
