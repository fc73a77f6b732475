import os, sys, ctypes as C
import numpy as np, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import fields, xlab_fftbarotropic_b200 as xfb
n = 4096; dt = 1.0
b = xfb.Backend(n); v0 = fields.elliptic(n); b.set_vorticity(v0)
stream = torch.cuda.ExternalStream(b.stream)
tfil = torch.empty((n, n), dtype=torch.float32, device="cuda"); deform = torch.empty_like(tfil)
cmin, cmax = float(v0.min()) - 1e-6, float(v0.max()) * 1.01
def timeit(fn, reps=20):
    for _ in range(3): fn()
    b.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record(stream)
        for _ in range(reps): fn()
        e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
print("step(1)           ", timeit(lambda: b.step(1, dt)))
print("diagnostics only  ", timeit(lambda: b._ck(b._L.xfb_get_diagnostics(b._h, 0, C.c_void_p(tfil.data_ptr()), C.c_void_p(deform.data_ptr())))))
print("keff only         ", timeit(lambda: b.keff_hist(64, cmin, cmax)))
print("step+diag         ", timeit(lambda: (b.step(1, dt), b._ck(b._L.xfb_get_diagnostics(b._h, 0, C.c_void_p(tfil.data_ptr()), C.c_void_p(deform.data_ptr()))))))
