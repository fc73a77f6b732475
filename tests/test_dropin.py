"""The drop-in boundary, proven with the reference's own sources (INTEGRATION.md sections 1-2):

  * oracle/_ref/{invert_pres,main}_xfb_n256.out are the UNMODIFIED /root/reference/src/invert_pres.cpp and main.cpp
    compiled (by oracle/build_oracle.py:build_dropin, in the build container) against include/compat/fftw3.h,
    include/compat/fftwfop.cpp -> include/fftwfop.hpp and linked with libxfb.so instead of FFTW: every fftwf_execute and
    every fftwf_operation method of those programs runs on the GPU.  They are run here next to the reference binaries
    built against the CPU FFT shim, on the same input files.
  * tests/cpp/fftwfop_wrapper.cpp instantiates the class like the reference does (global object, five methods, index
    helpers) and is compiled here with g++.
"""
import os
import subprocess
import tempfile

import numpy as np
import pytest

from conftest import rel_l2

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFDIR = os.path.join(ROOT, "oracle", "_ref")
GOLD = os.path.join(ROOT, "tests", "golden")


def _need(*names):
    paths = [os.path.join(REFDIR, n) for n in names]
    if not all(os.path.exists(p) for p in paths):
        pytest.skip("drop-in binaries are built where /root/reference exists (python __graft_entry__.py)")
    return paths


def test_compat_headers_declare_the_fftw_names_the_reference_uses():
    """CPU: the eight FFTW names of main.cpp:103-135 / invert_pres.cpp:84-107 and the class surface of fftwfop.hpp:9-29"""
    h = open(os.path.join(ROOT, "include", "compat", "fftw3.h")).read()
    for name in ("fftwf_complex", "fftwf_plan", "FFTW_ESTIMATE", "fftwf_malloc", "fftwf_free", "fftwf_plan_dft_r2c_2d",
                 "fftwf_plan_dft_c2r_2d", "fftwf_execute"):
        assert name in h, name
    # the wrapper and the compat header compile as C++11 without a GPU (syntax + template instantiation)
    src = os.path.join(ROOT, "tests", "cpp", "fftwfop_wrapper.cpp")
    r = subprocess.run(["g++", "-std=c++11", "-fsyntax-only", "-I", os.path.join(ROOT, "include", "compat"), "-I",
                        os.path.join(ROOT, "include"), src], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


@pytest.mark.gpu
def test_fftwfop_wrapper_class_matches_oracle(tmp_path):
    from oracle import oracle as orc
    n = 256
    exe = str(tmp_path / "wrapper.out")
    lib = os.path.join(ROOT, "xlab_fftbarotropic_b200")
    subprocess.run(["g++", "-std=c++11", "-O2", "-I", os.path.join(ROOT, "include", "compat"), "-I", os.path.join(ROOT, "include"),
                    "-o", exe, os.path.join(ROOT, "tests", "cpp", "fftwfop_wrapper.cpp"), "-L", lib, "-lxfb", f"-Wl,-rpath,{lib}"],
                   check=True)
    rng = np.random.default_rng(11)
    h = n // 2 + 1
    z = (rng.standard_normal((n, h)) + 1j * rng.standard_normal((n, h))).astype(np.complex64)
    z.tofile(str(tmp_path / "in.bin"))
    r = subprocess.run([exe, str(tmp_path / "in.bin"), str(tmp_path / "out")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    # index helpers, fftwfop.hpp:26-28
    assert f"HIDX(3,5)={h * 3 + 5} R_HIDX(3,5)={h * (n - 3) + 5} reflected(1)={n - 1} reflected({n // 2})={n // 2}" in r.stdout
    o = orc.Oracle(n)
    for name, ref in (("gradx", o.gradx(z)), ("grady", o.grady(z)), ("laplacian", o.laplacian(z)),
                      ("invertLaplacian", o.invert_laplacian(z)), ("dealiase", o.dealias(z))):
        got = np.fromfile(str(tmp_path / f"out.{name}"), dtype=np.complex64).reshape(n, h)
        assert np.array_equal(got.view(np.float32), ref.view(np.float32)), name       # bit-exact like the operator tier


@pytest.mark.gpu
def test_unmodified_invert_pres_cpp_on_the_gpu_backend():
    exe_xfb, exe_ref = _need("invert_pres_xfb_n256.out", "invert_pres_n256.out")
    n = 256
    psi = np.load(os.path.join(GOLD, "ref_n256.npz"))["elliptic_psi_1"]
    with tempfile.TemporaryDirectory() as d:
        psi.tofile(os.path.join(d, "psi.bin"))
        out = {}
        for tag, exe in (("xfb", exe_xfb), ("ref", exe_ref)):
            r = subprocess.run([exe, "-x", "3", "-y", "5"], input=f"{d}/psi.bin=>{d}/pres_{tag}.bin\nnot a pair\n", text=True,
                               capture_output=True, cwd=d)
            assert r.returncode == 0, r.stderr
            assert "Error reading input: not a pair" in r.stdout
            out[tag] = np.fromfile(os.path.join(d, f"pres_{tag}.bin"), dtype="<f4").reshape(n, n)
    assert out["xfb"][5, 3] == 0.0
    assert rel_l2(out["xfb"], out["ref"]) < 1e-5, rel_l2(out["xfb"], out["ref"])


@pytest.mark.gpu
def test_unmodified_main_cpp_on_the_gpu_backend():
    exe_xfb, exe_ref = _need("main_xfb_n256.out", "main_n256.out")
    n = 256
    v0 = np.load(os.path.join(GOLD, "ref_n256.npz"))["elliptic_init"]
    env = dict(os.environ, XFB_DT="3.0", XFB_TOTAL_STEPS="3", XFB_RECORD_STEP="2", XFB_SHIM_THREADS="4")
    res = {}
    for tag, exe in (("xfb", exe_xfb), ("ref", exe_ref)):
        with tempfile.TemporaryDirectory() as d:
            os.makedirs(os.path.join(d, "input"))
            os.makedirs(os.path.join(d, "output"))
            v0.tofile(os.path.join(d, "input", "initial_vorticity.bin"))
            r = subprocess.run([exe], cwd=d, capture_output=True, text=True, env=env)
            assert r.returncode == 0, r.stderr[-2000:]
            assert "Program ends. Congrats!" in r.stdout
            log = [ln.strip() for ln in open(os.path.join(d, "log")) if ln.strip()]
            res[tag] = (log, {p: np.fromfile(os.path.join(d, p), dtype="<f4").reshape(n, n) for p in log})
    assert res["xfb"][0] == res["ref"][0]                  # same files in the same order
    for p in res["ref"][0]:
        if "vort_src_input" in p:
            continue                                       # main.cpp never initialises vort_src (main.cpp:110)
        assert rel_l2(res["xfb"][1][p], res["ref"][1][p]) < 1e-5, p
