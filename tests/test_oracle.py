"""CPU tests (-m "not gpu"): the oracle against the golden vectors of the unmodified reference,
against scipy, and the known-answer tests of SURVEY.md section 4."""
import hashlib
import json
import os

import numpy as np
import pytest
import scipy.fft as sf

from conftest import rel_l2
import fields

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def orc():
    from oracle import oracle
    return oracle


def test_restatement_matches_golden_bit_for_bit(orc):
    """The C restatement, linked against the same FFT shim, reproduces the reference binary's record files exactly."""
    g = np.load(os.path.join(GOLD, "ref_n64.npz"))
    o = orc.Oracle(64)
    o.set_vorticity(g["init"])
    for step in (0, 3):
        if step:
            o.step(3, 3.0)
        for kind, which in (("vort", orc.VORT), ("psi", orc.PSI), ("u", orc.U), ("v", orc.V)):
            assert np.array_equal(o.get_field(which), g[f"{kind}_{step}"]), (kind, step)


def test_restatement_matches_golden_256(orc):
    g = np.load(os.path.join(GOLD, "ref_n256.npz"))
    o = orc.Oracle(256)
    o.set_vorticity(g["elliptic_init"])
    o.step(1, 3.0)
    for kind, which in (("vort", orc.VORT), ("psi", orc.PSI), ("u", orc.U), ("v", orc.V)):
        assert np.array_equal(o.get_field(which), g[f"elliptic_{kind}_1"]), kind
    assert np.array_equal(o.invert_pres(g["elliptic_psi_1"], 3, 5), g["elliptic_pres_1"])
    o2 = orc.Oracle(256)
    o2.set_vorticity(g["kuo_init"])
    o2.step(10, 3.0)
    assert np.array_equal(o2.get_field(orc.VORT), g["kuo_vort_10"])


def test_restatement_matches_reference_binary(orc):
    """Live version of the pin above, when the reference binaries are available (build container or prebuilt)."""
    from oracle import build_oracle
    if not build_oracle.build_reference(96, programs=("main",)):
        pytest.skip("no reference binary for n=96")
    v0 = fields.kuo2004(96)
    ref = orc.run_reference_main(v0, 96, 3.0, 4, 1)
    o = orc.Oracle(96)
    o.set_vorticity(v0)
    for s in range(4):
        assert np.array_equal(o.get_field(orc.VORT), ref[("vort", s)])
        assert np.array_equal(o.get_field(orc.U), ref[("u", s)])
        o.step(1, 3.0)


@pytest.mark.parametrize("n", [64, 96, 100, 256, 768])
def test_shim_fft_against_scipy(orc, n):
    """the FFTW stand-in computes FFTW's published transform: rfft2 / unnormalised irfft2"""
    rng = np.random.default_rng(n)
    f = rng.standard_normal((n, n)).astype(np.float32)
    o = orc.Oracle(n)
    F = o.r2c(f)
    assert rel_l2(F, sf.rfft2(f.astype(np.float64))) < 5e-7
    back = o.c2r(F) / np.float32(n * n)
    assert rel_l2(back, f) < 1e-6
    # c2r of a non-Hermitian-consistent spectrum: Im of the j = 0 and j = n/2 bins is dropped after the x pass
    z = (rng.standard_normal((n, n // 2 + 1)) + 1j * rng.standard_normal((n, n // 2 + 1))).astype(np.complex64)
    zx = sf.ifft(z.astype(np.complex128), axis=0) * n
    zx[:, 0] = zx[:, 0].real
    if n % 2 == 0:
        zx[:, -1] = zx[:, -1].real
    expect = sf.irfft(zx, n=n, axis=1) * n
    assert rel_l2(o.c2r(z), expect) < 5e-7


def test_tables_follow_the_reference_formulas(orc):
    n, L = 96, 600000.0
    o = orc.Oracle(n, L, L)
    kx, ky, lap, lapinv, mask = (o.table(i) for i in range(5))
    twopi = np.float32(np.arccos(np.float32(-1.0)) * np.float32(2.0))
    assert kx[5] == np.float32(np.float32(twopi * np.float32(5)) / np.float32(L))
    assert kx[n // 2] > 0 and kx[n // 2 + 1] == -kx[n // 2 - 1]          # Nyquist kept with sign +
    assert ky.shape == (n // 2 + 1,)
    assert lapinv[0, 0] == 1.0 and lap[0, 0] == 0.0
    assert np.array_equal(lap[1:, :], lap[:0:-1, :][::1][::-1][::-1]) or np.array_equal(lap[1:], lap[1:][::-1])
    # KAT-6 mask census: circle of radius sqrt(2) ceil(N/3) in index space, NOT the textbook 2/3 square
    i = np.minimum(np.arange(n), n - np.arange(n))[:, None]
    j = np.arange(n // 2 + 1)[None, :]
    assert np.array_equal(mask, (i * i + j * j < 2 * int(np.ceil(n / 3.0)) ** 2).astype(np.float32))


def test_kat_poisson_round_trip(orc):
    # KAT-2: invertLaplacian(laplacian(x)) == x except mode (0,0) -> 0
    n = 64
    rng = np.random.default_rng(0)
    z = (rng.standard_normal((n, n // 2 + 1)) + 1j * rng.standard_normal((n, n // 2 + 1))).astype(np.complex64)
    o = orc.Oracle(n)
    back = o.invert_laplacian(o.laplacian(z))
    assert back[0, 0] == 0
    back[0, 0] = z[0, 0]
    assert rel_l2(back, z) < 2e-7


def test_kat_linear_decay_and_invariants(orc):
    # KAT-3/5 on the oracle itself
    n, dt, nu, L = 64, 3.0, 6.5, 600000.0
    o = orc.Oracle(n)
    z = np.zeros((n, n // 2 + 1), np.complex64)
    z[2, 3] = 1.0
    o.set_spectrum(z)
    o.step(1, dt)
    out = o.get_spectrum()
    zz = -nu * ((2 * np.pi * 2 / L) ** 2 + (2 * np.pi * 3 / L) ** 2) * dt
    assert abs(out[2, 3] - (1 + zz + zz ** 2 / 2 + zz ** 3 / 6 + zz ** 4 / 24)) < 1e-6
    o.set_vorticity(fields.gaussian(n))
    z0 = o.get_spectrum()
    o.step(5, dt)
    assert o.get_spectrum()[0, 0] == z0[0, 0]


def test_numpy_generators_against_reference_md5():
    """tests/fields.py restates the reference generators; const and gaussian are byte-identical at 768^2,
    elliptic and Kuo2004 to within one float32 ulp of the peak (libm exp differences)."""
    md5 = json.load(open(os.path.join(GOLD, "generators.json")))
    assert hashlib.md5(fields.const_vortex(768).tobytes()).hexdigest() == md5["makefield-const-vortex"]
    assert hashlib.md5(fields.gaussian(768).tobytes()).hexdigest() == md5["makefield-gaussian"]
    g = np.load(os.path.join(GOLD, "ref_n256.npz"))
    assert np.abs(fields.elliptic(256) - g["elliptic_init"]).max() < 1e-9
    assert np.abs(fields.kuo2004(256) - g["kuo_init"]).max() < 5e-9


def test_diagnostic_definitions(orc):
    """parity unpinned: properties that follow from the definitions (Rozoff 2006; builder-defined deformation factor)"""
    n = 128
    o = orc.Oracle(n)
    o.set_vorticity(fields.gaussian(n))
    tfil, deform, s1, s2 = o.diagnostics()
    assert np.all(deform <= 1.0 + 1e-6) and np.all(deform >= -1.0 - 1e-6)
    assert np.all((tfil > 0) == (deform > 0))              # filamentation time defined exactly where strain dominates
    centre = deform[n // 2, n // 2]
    assert centre < 0                                        # vortex core: rotation dominates
    area, g2 = o.keff_hist(32, -1e-6, 1.1e-3)
    assert abs(area.sum() - 600000.0 ** 2) < 1e-6 * 600000.0 ** 2
    a, k = orc.keff_from_hist(32, -1e-6, 1.1e-3, 6.5, area, g2)
    assert a[0] >= a[10] >= a[-1] and np.all(k >= 0)
