"""ctypes front end of the CPU oracle (oracle/_ref/liboracle.so).

TEST INFRASTRUCTURE: imported only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference leg.  Nothing under xlab_fftbarotropic_b200/ imports it.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import tempfile

import numpy as np

from . import build_oracle

_LIB = None

VORT, PSI, U, V, SRC = 0, 1, 2, 3, 4


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(build_oracle.OUT, "liboracle.so")
        if not os.path.exists(path) or os.path.exists(os.path.join(build_oracle.HERE, "barotropic_oracle.c")):
            try:
                path = build_oracle.build_restatement()
            except Exception:
                if not os.path.exists(path):
                    raise
        L = C.CDLL(path)
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.c_int, C.c_float, C.c_float, C.c_float]
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_table.restype = C.POINTER(C.c_float)
        L.orc_table.argtypes = [C.c_void_p, C.c_int]
        for name in ("orc_gradx", "orc_grady", "orc_laplacian", "orc_invert_laplacian", "orc_dealias",
                     "orc_r2c", "orc_c2r"):
            getattr(L, name).argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_set_vorticity.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_set_spectrum.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_get_spectrum.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_set_source.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_step.argtypes = [C.c_void_p, C.c_int, C.c_float]
        L.orc_get_field.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.orc_invert_pres.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_float, C.c_float]
        L.orc_diagnostics.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_keff_hist.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_void_p, C.c_void_p]
        L.orc_tracer_keff_hist.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_void_p, C.c_void_p]
        L.orc_set_tracer.argtypes = [C.c_void_p, C.c_void_p, C.c_float]
        L.orc_get_tracer.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_keff_from_hist.argtypes = [C.c_int, C.c_float, C.c_float, C.c_double, C.c_void_p, C.c_void_p,
                                         C.c_void_p, C.c_void_p]
        L.xfb_shim_set_threads.argtypes = [C.c_int]
        _LIB = L
    return _LIB


def set_threads(n: int):
    """threads of the shim FFT (0 = all); the reference itself is single-threaded"""
    lib().xfb_shim_set_threads(int(n))


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class Oracle:
    """CPU restatement of the reference operator class + RK4 driver (float32, reference layout:
    real N x N `IDX = N*i + j`, half spectrum N x (N/2+1) complex64 `HIDX = (N/2+1)*i + j`)."""

    def __init__(self, n: int, lx: float = 600000.0, ly: float = 600000.0, nu: float = 6.5):
        self.n, self.h = n, n // 2 + 1
        self.lx, self.ly, self.nu = lx, ly, nu
        self._o = lib().orc_create(n, lx, ly, nu)

    def __del__(self):
        if getattr(self, "_o", None):
            lib().orc_destroy(self._o)
            self._o = None

    # tables -------------------------------------------------------------------------------
    def table(self, which: int) -> np.ndarray:
        n, h = self.n, self.h
        shape = {0: (n,), 1: (h,), 2: (n, h), 3: (n, h), 4: (n, h)}[which]
        ptr = lib().orc_table(self._o, which)
        return np.ctypeslib.as_array(ptr, shape=(int(np.prod(shape)),)).reshape(shape).copy()

    # operators ----------------------------------------------------------------------------
    def _cop(self, name, a):
        a = np.ascontiguousarray(a, dtype=np.complex64).reshape(self.n, self.h)
        out = np.empty_like(a)
        getattr(lib(), name)(self._o, _p(a), _p(out))
        return out

    def gradx(self, a): return self._cop("orc_gradx", a)
    def grady(self, a): return self._cop("orc_grady", a)
    def laplacian(self, a): return self._cop("orc_laplacian", a)
    def invert_laplacian(self, a): return self._cop("orc_invert_laplacian", a)
    def dealias(self, a): return self._cop("orc_dealias", a)

    def r2c(self, f):
        f = np.ascontiguousarray(f, dtype=np.float32).reshape(self.n, self.n).copy()
        out = np.empty((self.n, self.h), np.complex64)
        lib().orc_r2c(self._o, _p(f), _p(out))
        return out

    def c2r(self, a):
        a = np.ascontiguousarray(a, dtype=np.complex64).reshape(self.n, self.h).copy()
        out = np.empty((self.n, self.n), np.float32)
        lib().orc_c2r(self._o, _p(a), _p(out))
        return out

    # model --------------------------------------------------------------------------------
    def set_vorticity(self, f):
        f = np.ascontiguousarray(f, dtype=np.float32).reshape(self.n, self.n)
        lib().orc_set_vorticity(self._o, _p(f))

    def set_spectrum(self, z):
        z = np.ascontiguousarray(z, dtype=np.complex64).reshape(self.n, self.h)
        lib().orc_set_spectrum(self._o, _p(z))

    def get_spectrum(self):
        z = np.empty((self.n, self.h), np.complex64)
        lib().orc_get_spectrum(self._o, _p(z))
        return z

    def set_source(self, s):
        if s is None:
            lib().orc_set_source(self._o, None)
        else:
            s = np.ascontiguousarray(s, dtype=np.float32).reshape(self.n, self.n)
            lib().orc_set_source(self._o, _p(s))

    def step(self, nsteps: int, dt: float):
        lib().orc_step(self._o, int(nsteps), float(dt))

    def get_field(self, which: int):
        out = np.empty((self.n, self.n), np.float32)
        lib().orc_get_field(self._o, which, _p(out))
        return out

    def invert_pres(self, psi, ref_x=0, ref_y=0, rho=1.0, f=1e-5):
        psi = np.ascontiguousarray(psi, dtype=np.float32).reshape(self.n, self.n)
        out = np.empty((self.n, self.n), np.float32)
        lib().orc_invert_pres(self._o, _p(psi), _p(out), ref_x, ref_y, rho, f)
        return out

    def diagnostics(self):
        tfil = np.empty((self.n, self.n), np.float32)
        deform = np.empty_like(tfil)
        s1 = np.empty_like(tfil)
        s2 = np.empty_like(tfil)
        lib().orc_diagnostics(self._o, _p(tfil), _p(deform), _p(s1), _p(s2))
        return tfil, deform, s1, s2

    def keff_hist(self, nbins: int, cmin: float, cmax: float):
        area = np.zeros(nbins, np.float64)
        g2 = np.zeros(nbins, np.float64)
        lib().orc_keff_hist(self._o, nbins, cmin, cmax, _p(area), _p(g2))
        return area, g2

    # passive tracer (no reference code: a restatement of dc/dt = -u.grad c + kappa lap c in the
    # vorticity tendency's own operation order, see barotropic_oracle.c get_dtrcdt)
    def set_tracer(self, c, kappa: float):
        c = np.ascontiguousarray(c, dtype=np.float32).reshape(self.n, self.n)
        lib().orc_set_tracer(self._o, _p(c), float(kappa))

    def get_tracer(self):
        out = np.empty((self.n, self.n), np.float32)
        lib().orc_get_tracer(self._o, _p(out))
        return out

    def tracer_keff_hist(self, nbins: int, cmin: float, cmax: float):
        area = np.zeros(nbins, np.float64)
        g2 = np.zeros(nbins, np.float64)
        lib().orc_tracer_keff_hist(self._o, nbins, cmin, cmax, _p(area), _p(g2))
        return area, g2


def keff_from_hist(nbins, cmin, cmax, kappa, area, g2):
    a = np.zeros(nbins + 1, np.float64)
    k = np.zeros(nbins + 1, np.float64)
    area = np.ascontiguousarray(area, np.float64)
    g2 = np.ascontiguousarray(g2, np.float64)
    lib().orc_keff_from_hist(nbins, cmin, cmax, kappa, _p(area), _p(g2), _p(a), _p(k))
    return a, k


# ---------------------------------------------------------------------------------------------
# initial fields: numpy restatement of the reference generators (float32 where the reference is)
# ---------------------------------------------------------------------------------------------

def run_reference_generator(name: str, n: int, workdir: str | None = None) -> np.ndarray:
    """Run the UNMODIFIED reference generator binary (built here; prebuilt on the GPU box)."""
    exe = build_oracle.build_reference(n, programs=(name,)).get(name)
    if exe is None:
        raise FileNotFoundError(f"no reference binary for {name} at n={n}")
    with tempfile.TemporaryDirectory(dir=workdir) as d:
        os.makedirs(os.path.join(d, "input"))
        subprocess.run([exe], cwd=d, check=True, stderr=subprocess.DEVNULL)
        return np.fromfile(os.path.join(d, "input", "initial_vorticity.bin"), dtype="<f4").reshape(n, n)


def run_reference_main(vort0: np.ndarray, n: int, dt: float, steps: int, record: int, env_threads: int = 1,
                       program: str = "main") -> dict:
    """Run the UNMODIFIED reference main.cpp for `steps` steps; returns {(kind, step): field}."""
    exe = build_oracle.build_reference(n, programs=(program,)).get(program)
    if exe is None:
        raise FileNotFoundError(f"no reference binary for n={n}")
    out = {}
    with tempfile.TemporaryDirectory() as d:
        os.makedirs(os.path.join(d, "input"))
        os.makedirs(os.path.join(d, "output"))
        np.ascontiguousarray(vort0, dtype="<f4").tofile(os.path.join(d, "input", "initial_vorticity.bin"))
        env = dict(os.environ, XFB_SHIM_THREADS=str(env_threads), XFB_DT=repr(float(dt)),
                   XFB_TOTAL_STEPS=str(int(steps)), XFB_RECORD_STEP=str(int(record)))
        subprocess.run([exe], cwd=d, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, env=env)
        for line in open(os.path.join(d, "log")):
            path = line.strip()
            if not path:
                continue
            base = os.path.basename(path)[:-4]
            kind, _, step = base.rpartition("_step_")
            out[(kind, int(step))] = np.fromfile(os.path.join(d, path), dtype="<f4").reshape(n, n)
    return out
