"""CPU tests: the C-ABI shared library loads and exports every symbol include/xfb.h declares (no compute
calls -- there is no GPU here), the product has no route into oracle/, and without a device it fails loudly."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "xfb.h")
LIB = os.path.join(ROOT, "xlab_fftbarotropic_b200", "libxfb.so")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(xfb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    assert os.path.exists(LIB), "build with `python -m xlab_fftbarotropic_b200.build`"
    lib = ctypes.CDLL(LIB)
    syms = declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/xfb.h but not exported"


def test_python_binding_covers_the_header():
    from xlab_fftbarotropic_b200 import capi
    assert sorted(capi.SYMBOLS) == declared_symbols()


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import xlab_fftbarotropic_b200 as xfb
    with pytest.raises(xfb.XfbError) as e:
        xfb.Backend(256)
    assert "no CUDA device" in str(e.value) or "CUDA" in str(e.value)
    assert xfb.load().xfb_size_supported(256, 256) == 1
    assert xfb.load().xfb_size_supported(768, 768) == 2          # the reference's default NPTS: generic mixed-radix path
    assert xfb.load().xfb_size_supported(254, 254) == 0          # 2 * 127


def test_product_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "xlab_fftbarotropic_b200")
    for base, _, files in os.walk(pkg):
        if "_build" in base:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                text = open(os.path.join(base, f), errors="ignore").read()
                assert not re.search(r"(import\s+oracle|from\s+oracle|#include\s*[<\"][^>\"]*oracle|liboracle|shim_fft|fftw3_shim)", text), \
                    f"{f} reaches into oracle/"
    out = subprocess.run(["ldd", LIB], capture_output=True, text=True).stdout
    assert "liboracle" not in out and "fftw" not in out and "cufft" not in out


def test_host_programs_are_built():
    for name in ("main.out", "invert_pres.out", "makefield-elliptic-vortex.out", "makefield-Kuo2004.out"):
        assert os.path.exists(os.path.join(ROOT, "xlab_fftbarotropic_b200", "bin", name)), name
