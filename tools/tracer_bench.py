"""Cost of the passive tracer: ms per RK4 step with and without it (CUDA events on the handle's stream).

    python tools/tracer_bench.py [--grids 4096 8192] [--steps 20]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))

import numpy as np
import torch

import fields
import xlab_fftbarotropic_b200 as xfb


def timed(b, steps, dt):
    st = torch.cuda.ExternalStream(b.stream)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    b.step(3, dt)
    b.sync()
    e0.record(st)
    b.step(steps, dt)
    e1.record(st)
    b.sync()
    return e0.elapsed_time(e1) / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grids", type=int, nargs="+", default=[1024, 4096, 8192])
    ap.add_argument("--steps", type=int, default=20)
    a = ap.parse_args()
    for n in a.grids:
        v0 = fields.const_vortex(n)
        b = xfb.Backend(n)
        b.set_vorticity(v0)
        base = timed(b, a.steps, 3.0)
        b.set_tracer(b.get_field(xfb.capi.VORT), 6.5)      # tracer := current vorticity, kappa = nu: must stay equal to it
        with_t = timed(b, a.steps, 3.0)
        z = b.get_field(xfb.capi.VORT)
        c = b.get_field(xfb.capi.TRACER)
        err = float(np.linalg.norm((c - z).ravel()) / np.linalg.norm(z.ravel()))
        print(json.dumps({"grid": n, "ms_per_step": round(base, 4), "ms_per_step_with_tracer": round(with_t, 4),
                          "tracer_cost_ms": round(with_t - base, 4), "tracer_vs_vorticity_rel_l2": err}))
        b.close()


if __name__ == "__main__":
    main()
