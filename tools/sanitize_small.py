"""Small end-to-end exercise of every kernel family for compute-sanitizer (memcheck):
  compute-sanitizer --tool memcheck python tools/sanitize_small.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import fields  # noqa: E402
import xlab_fftbarotropic_b200 as xfb  # noqa: E402

for n in (256, 1024, 96):
    b = xfb.Backend(n, batch=2 if n == 256 else 1)
    v0 = fields.elliptic(n)
    for m in range(b.batch):
        b.set_vorticity(v0, member=m)
    b.set_source((1e-9 * v0).astype(np.float32))
    b.step(3, 3.0)
    for w in (xfb.capi.VORT, xfb.capi.PSI, xfb.capi.U, xfb.capi.TFIL):
        assert np.isfinite(b.get_field(w)).all()
    b.diagnostics()
    b.keff_hist(32, float(v0.min()) - 1e-6, float(v0.max()) * 1.01)
    b.invert_pres(b.get_field(xfb.capi.PSI), 1, 2)
    z = b.get_spectrum()
    b.set_spectrum(z)
    b.step(1, 3.0)
    b.r2c(v0)
    b.close()
t = xfb.LoopbackTeam(256, 2, 2)
t.set_vorticity(fields.kuo2004(256))
t.step(2, 3.0)
assert np.isfinite(t.get_field(xfb.capi.VORT)).all()
t.close()
print("sanitize_small ok")
