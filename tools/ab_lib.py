"""A/B of two builds of libxfb.so on one GPU: per grid, time K steps (host clock around xfb_step + xfb_sync and the
per-kernel event times of xfb_profile) and hash the spectral state, so that a variant that must be bit-identical to
the baseline (same arithmetic, different synchronisation) can be checked and timed in one short run.

    XFB_LIB=/path/to/libxfb_variant.so python tools/ab_lib.py 8192 4096 [--steps 20] [--forcing] [--tracer]

Prints one JSON line per grid.  No torch import (a fresh box pays a minute for it)."""
import hashlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import fields  # noqa: E402
import xlab_fftbarotropic_b200 as xfb  # noqa: E402


def dt_for(n):
    return {256: 8.0, 512: 6.0, 1024: 4.0, 2048: 2.0, 4096: 1.0, 8192: 0.5, 16384: 0.25}.get(n, 1.0)


def main():
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("grids", type=int, nargs="*", default=[8192])
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--forcing", action="store_true")
    ap.add_argument("--tracer", action="store_true")
    a = ap.parse_args()
    steps, grids = a.steps, a.grids
    for n in grids:
        b = xfb.Backend(n)
        v0 = fields.elliptic(n)
        b.set_vorticity(v0)
        if a.forcing:
            b.set_source((1e-9 * np.roll(v0, n // 7, axis=0)).astype(np.float32))
        if a.tracer:
            b.set_tracer(np.roll(v0, n // 5, axis=1).astype(np.float32), 6.5)
        dt = dt_for(n)
        b.step(3, dt)
        b.sync()
        b.profile(True)
        b.step(steps, dt)
        b.sync()
        prof = b.profile_read()
        b.profile(False)
        b.step(2, dt)
        b.sync()
        t0 = time.perf_counter()
        b.step(steps, dt)
        b.sync()
        ms = (time.perf_counter() - t0) * 1e3 / steps
        z = b.get_spectrum(0)
        h = hashlib.sha256(np.ascontiguousarray(z).view(np.uint8)).hexdigest()[:16]
        out = {"lib": os.path.basename(xfb.capi.LIB_PATH), "grid": n, "ms_per_step": round(ms, 4),
               "row_ms": round(prof["row_ms"] / max(1, prof["row_launches"]), 4),
               "col_ms": round(prof["col_ms"] / max(1, prof["col_launches"]), 4),
               "finite": bool(np.isfinite(z.view(np.float32)).all()), "state_sha": h}
        if a.tracer:
            c = b.get_field(xfb.capi.TRACER)
            out["tracer_sha"] = hashlib.sha256(np.ascontiguousarray(c).view(np.uint8)).hexdigest()[:16]
        print(json.dumps(out), flush=True)
        b.close()


if __name__ == "__main__":
    main()
