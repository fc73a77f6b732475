"""Shortest program that runs the stepper (for ncu): python tools/prof_step.py GRID [STEPS]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import fields  # noqa: E402
import xlab_fftbarotropic_b200 as xfb  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
b = xfb.Backend(n)
b.set_vorticity(fields.elliptic(n))
b.step(steps, 0.5 if n >= 8192 else 1.0)
b.sync()
print("ok", b.launch_count)
b.close()
