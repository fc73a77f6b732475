"""Golden vectors for BASELINE.json's headline configurations, from the UNMODIFIED reference main.cpp
(oracle/_ref/main_n<N>.out = /root/reference/src/main.cpp compiled where it lies against the FFT shim).

Run in the build container (needs /root/reference):   python tests/golden/make_golden_headline.py

  ref_kuo1024.npz   configs[1] "Kuo et al. 2004 vortex at 1024^2": the reference generator's field, advanced by the
                    reference binary with its own dt = 3 s; `init` = the generator's field (full), vort/psi/u/v after
                    1 step, vort after 1000 steps.
                    A 1024^2 field is 4 MB, so each field is stored as
                      <name>_sub   : every 4th point in both directions (256 x 256 float32), and
                      <name>_blk   : float64 sums and sums of squares over the 16 x 16 blocks of 64 x 64 points
                                     (the full field enters the check through them)
  ref_elliptic4096.npz   configs[2] "elliptic vortex at 4096^2", dt = 1 s: vort after 1 step, same storage (stride 16)

The tests recompute the same reductions from the GPU result (tests/test_headline_parity.py).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import oracle as orc  # noqa: E402


def reduce_field(f, stride, nblk=16):
    n = f.shape[0]
    b = n // nblk
    f64 = f.astype(np.float64).reshape(nblk, b, nblk, b)
    return {"sub": np.ascontiguousarray(f[::stride, ::stride]),
            "blk": np.stack([f64.sum(axis=(1, 3)), (f64 * f64).sum(axis=(1, 3))])}


def main():
    threads = os.cpu_count() or 1
    out = {}
    n = 1024
    k = orc.run_reference_generator("makefield-Kuo2004", n)
    out["init"] = k          # full field: almost all zeros, compresses to a few KB (tests/fields.py differs from the
                             # reference generator by 1 ulp of exp() at ~700 skirt points, so the tests start from THIS)
    r1 = orc.run_reference_main(k, n, 3.0, 2, 1, env_threads=threads)
    for kind in ("vort", "psi", "u", "v"):
        for key, val in reduce_field(r1[(kind, 1)], 4).items():
            out[f"{kind}_1_{key}"] = val
    r1000 = orc.run_reference_main(k, n, 3.0, 1001, 1000, env_threads=threads)
    for key, val in reduce_field(r1000[("vort", 1000)], 4).items():
        out[f"vort_1000_{key}"] = val
    np.savez_compressed(os.path.join(HERE, "ref_kuo1024.npz"), **out)

    n = 4096
    # the 4096 generator binary is not prebuilt; tests/fields.py is bit-identical to it at 768 (generators.json)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import fields
    e = fields.elliptic(n)
    r = orc.run_reference_main(e, n, 1.0, 2, 1, env_threads=threads)
    out4 = {}
    for kind in ("vort", "u"):
        for key, val in reduce_field(r[(kind, 1)], 16).items():
            out4[f"{kind}_1_{key}"] = val
    np.savez_compressed(os.path.join(HERE, "ref_elliptic4096.npz"), **out4)
    print("headline golden vectors written to", HERE)


if __name__ == "__main__":
    main()
