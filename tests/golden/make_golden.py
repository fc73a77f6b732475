"""Generate the golden vectors in this directory from the UNMODIFIED reference programs.

Run in the build container (needs /root/reference):   python tests/golden/make_golden.py
The reference binaries are built by oracle/build_oracle.py (reference sources compiled where they lie,
FFTW replaced by the shim in oracle/fftw3_shim).  Outputs:

  ref_n64.npz    elliptic vortex (reference generator binary) at 64^2, dt=3: vort/psi/u/v at record steps 0 and 3
  ref_n256.npz   elliptic 256^2: vort/psi/u/v at step 1; Kuo2004 256^2: vort at step 10;
                 invert_pres of the elliptic psi at step 1 (ref point x=3, y=5)
  generators.json  md5 of the four reference generators' 768^2 output
"""
import hashlib
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import build_oracle, oracle as orc  # noqa: E402


def main():
    out = {}
    v64 = orc.run_reference_generator("makefield-elliptic-vortex", 64)
    r = orc.run_reference_main(v64, 64, 3.0, 4, 3)
    out64 = {"init": v64}
    for (kind, step), f in r.items():
        if kind != "vort_src_input":
            out64[f"{kind}_{step}"] = f
    np.savez_compressed(os.path.join(HERE, "ref_n64.npz"), **out64)

    v256 = orc.run_reference_generator("makefield-elliptic-vortex", 256)
    r = orc.run_reference_main(v256, 256, 3.0, 2, 1)
    out256 = {"elliptic_init": v256}
    for kind in ("vort", "psi", "u", "v"):
        out256[f"elliptic_{kind}_1"] = r[(kind, 1)]
    k256 = orc.run_reference_generator("makefield-Kuo2004", 256)
    r2 = orc.run_reference_main(k256, 256, 3.0, 11, 10)
    out256["kuo_init"] = k256
    out256["kuo_vort_10"] = r2[("vort", 10)]
    exe = build_oracle.build_reference(256, programs=("invert_pres",))["invert_pres"]
    with tempfile.TemporaryDirectory() as d:
        r[("psi", 1)].tofile(os.path.join(d, "psi.bin"))
        subprocess.run([exe, "-x", "3", "-y", "5"], input=f"{d}/psi.bin=>{d}/pres.bin\n", text=True, check=True,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        out256["elliptic_pres_1"] = np.fromfile(os.path.join(d, "pres.bin"), dtype="<f4").reshape(256, 256)
    np.savez_compressed(os.path.join(HERE, "ref_n256.npz"), **out256)

    md5 = {}
    for g in build_oracle.GENERATORS:
        md5[g] = hashlib.md5(orc.run_reference_generator(g, 768).tobytes()).hexdigest()
    with open(os.path.join(HERE, "generators.json"), "w") as fh:
        json.dump(md5, fh, indent=1)
    print("golden vectors written to", HERE)


if __name__ == "__main__":
    main()
