"""GPU parity: every C-ABI entry point against the CPU oracle on the same inputs.

Tolerances (BASELINE.json north_star): relative L2 <= 1e-5 after one RK4 step, <= 1e-3 after 1000
steps, diagnostics within the same tolerance.  Pointwise operators and tables are required to be
BIT-EXACT (they are float expressions with one rounding each); single transforms must agree to
5e-7 relative L2 (two float32 FFTs with different factorizations).
"""
import numpy as np
import pytest

from conftest import rel_l2
import fields

pytestmark = pytest.mark.gpu

SIZES = [256, 512, 1024]


@pytest.fixture(scope="module")
def xfb():
    import xlab_fftbarotropic_b200 as pkg
    return pkg


@pytest.fixture(scope="module")
def orc():
    from oracle import oracle
    return oracle


def _spec(rng, n):
    h = n // 2 + 1
    return (rng.standard_normal((n, h)) + 1j * rng.standard_normal((n, h))).astype(np.complex64)


@pytest.mark.parametrize("n", [256, 1024])
def test_tables_bit_exact(xfb, orc, n):
    b = xfb.Backend(n)
    o = orc.Oracle(n)
    for which in range(5):
        assert np.array_equal(b.table(which), o.table(which)), f"table {which}"
    # KAT-6 mask census
    i = np.minimum(np.arange(n), n - np.arange(n))[:, None].astype(np.int64)
    j = np.arange(n // 2 + 1)[None, :].astype(np.int64)
    kd = 2 * int(np.ceil(n / 3.0)) ** 2
    assert int(b.table(4).sum()) == int((i * i + j * j < kd).sum())
    b.close()


@pytest.mark.parametrize("n", [256, 512])
def test_pointwise_operators_bit_exact(xfb, orc, n):
    rng = np.random.default_rng(n)
    z = _spec(rng, n)
    b = xfb.Backend(n)
    o = orc.Oracle(n)
    for name_g, name_o in (("gradx", "gradx"), ("grady", "grady"), ("laplacian", "laplacian"),
                           ("invertLaplacian", "invert_laplacian"), ("dealiase", "dealias")):
        got = getattr(b, name_g)(z)
        exp = getattr(o, name_o)(z)
        assert np.array_equal(got.view(np.float32), exp.view(np.float32)), name_g
    b.close()


@pytest.mark.parametrize("n", SIZES + [2048])
def test_r2c_c2r_against_oracle(xfb, orc, n):
    rng = np.random.default_rng(7 + n)
    f = rng.standard_normal((n, n)).astype(np.float32)
    b = xfb.Backend(n)
    o = orc.Oracle(n)
    F = b.r2c(f)
    Fo = o.r2c(f)
    assert rel_l2(F, Fo) < 5e-7
    # non-Hermitian-consistent input: c2r must drop Im at j = 0 and j = N/2 after the x pass
    z = _spec(rng, n)
    g = b.c2r(z)
    go = o.c2r(z)
    assert rel_l2(g, go) < 5e-7
    # round trip, unnormalised like FFTW
    back = b.c2r(F) / np.float32(n * n)
    assert rel_l2(back, f) < 1e-6
    b.close()


def test_kat_derivative(xfb):
    # KAT-1: zeta = sin(2 pi m x / L) -> gradx gives (2 pi m / L) cos, grady gives 0
    n, m = 256, 5
    L = 600000.0
    b = xfb.Backend(n)
    x = (np.arange(n) * (L / n))[:, None] * np.ones((1, n))
    f = np.sin(2 * np.pi * m * x / L).astype(np.float32)
    F = b.r2c(f)
    dx = b.c2r(b.gradx(F)) / (n * n)
    dy = b.c2r(b.grady(F)) / (n * n)
    ref = (2 * np.pi * m / L) * np.cos(2 * np.pi * m * x / L)
    assert rel_l2(dx, ref) < 2e-6
    assert np.abs(dy).max() < 1e-6 * np.abs(ref).max()
    b.close()


@pytest.mark.parametrize("n,gen", [(256, "elliptic"), (512, "kuo2004"), (1024, "gaussian")])
def test_one_step_parity(xfb, orc, n, gen):
    v0 = fields.GENERATORS[gen](n)
    b = xfb.Backend(n)
    o = orc.Oracle(n)
    b.set_vorticity(v0)
    o.set_vorticity(v0)
    assert rel_l2(b.get_spectrum(), o.get_spectrum()) < 5e-7
    b.step(1, 3.0)
    o.step(1, 3.0)
    assert rel_l2(b.get_spectrum(), o.get_spectrum()) < 1e-5
    for which in (xfb.capi.VORT, xfb.capi.PSI, xfb.capi.U, xfb.capi.V):
        assert rel_l2(b.get_field(which), o.get_field(which)) < 1e-5, which
    b.close()


def test_multi_step_and_chained_calls(xfb, orc):
    n = 256
    v0 = fields.elliptic(n)
    b = xfb.Backend(n)
    o = orc.Oracle(n)
    b.set_vorticity(v0)
    o.set_vorticity(v0)
    b.step(3, 3.0)
    b.step(2, 3.0)          # chained call reuses the prologue of the previous one
    b.get_field(xfb.capi.PSI)
    b.step(5, 3.0)
    o.step(10, 3.0)
    assert rel_l2(b.get_field(xfb.capi.VORT), o.get_field(orc.VORT)) < 1e-5


def test_thousand_steps(xfb, orc):
    # north_star: <= 1e-3 after 1000 steps (Kuo 2004 binary vortex, the reference's test/02 input)
    n = 256
    v0 = fields.kuo2004(n)
    b = xfb.Backend(n)
    o = orc.Oracle(n)
    b.set_vorticity(v0)
    o.set_vorticity(v0)
    b.step(1000, 3.0)
    o.step(1000, 3.0)
    err = rel_l2(b.get_field(xfb.capi.VORT), o.get_field(orc.VORT))
    assert err < 1e-3, err


def test_source_term(xfb, orc):
    n = 256
    v0 = fields.gaussian(n)
    src = (fields.kuo2004(n) * np.float32(1e-4)).astype(np.float32)
    b = xfb.Backend(n)
    o = orc.Oracle(n)
    for m in (b, o):
        m.set_vorticity(v0)
        m.set_source(src)
        m.step(2, 3.0)
        m.set_source(None)
        m.step(1, 3.0)
    assert rel_l2(b.get_field(xfb.capi.VORT), o.get_field(orc.VORT)) < 1e-5
    assert np.array_equal(b.get_field(xfb.capi.SRC), np.zeros((n, n), np.float32))


def test_ensemble_members_are_independent(xfb, orc):
    n, nb = 256, 3
    b = xfb.Backend(n, batch=nb)
    gens = ["elliptic", "gaussian", "kuo2004"]
    for m, g in enumerate(gens):
        b.set_vorticity(fields.GENERATORS[g](n), member=m)
    b.step(2, 3.0)
    for m, g in enumerate(gens):
        o = orc.Oracle(n)
        o.set_vorticity(fields.GENERATORS[g](n))
        o.step(2, 3.0)
        assert rel_l2(b.get_field(xfb.capi.VORT, member=m), o.get_field(orc.VORT)) < 1e-5, g
    b.close()


def test_kat_linear_decay_and_frozen_mode(xfb):
    # KAT-3: a single mode inside the mask has zero Jacobian: one RK4 step multiplies it by
    # 1+z+z^2/2+z^3/6+z^4/24, z = -nu |k|^2 dt; a mode outside the mask is frozen (tendency masked)
    n, dt, nu, L = 256, 3.0, 6.5, 600000.0
    b = xfb.Backend(n)
    h = n // 2 + 1
    for (i, j, inside) in ((3, 4, True), (100, 100, False)):
        z = np.zeros((n, h), np.complex64)
        z[i, j] = 1.0 + 0.5j
        b.set_spectrum(z)
        b.step(1, dt)
        out = b.get_spectrum()
        if inside:
            k2 = (2 * np.pi * i / L) ** 2 + (2 * np.pi * j / L) ** 2
            zz = -nu * k2 * dt
            amp = 1 + zz + zz ** 2 / 2 + zz ** 3 / 6 + zz ** 4 / 24
            assert abs(out[i, j] - amp * z[i, j]) < 1e-6 * abs(z[i, j])
        else:
            assert out[i, j] == z[i, j]
        out[i, j] = 0
        assert np.abs(out).max() < 1e-6
    b.close()


def test_diagnostics_against_cpu_restatement(xfb, orc):
    n = 512
    v0 = fields.elliptic(n)
    b = xfb.Backend(n)
    o = orc.Oracle(n)
    b.set_vorticity(v0)
    o.set_vorticity(v0)
    b.step(2, 3.0)
    o.set_spectrum(b.get_spectrum())     # same state: isolates the diagnostic kernels
    tfil, deform, _, _ = o.diagnostics()
    gt = b.get_field(xfb.capi.TFIL)
    gd = b.get_field(xfb.capi.DEFORM)
    gt2, gd2 = b.diagnostics()                      # fused path: COL_DIAG + ROW_DIAG (two launches)
    assert rel_l2(gd2, deform) < 1e-4
    both2 = (tfil > 0) & (gt2 > 0) & (tfil < 1e6)
    assert both2.mean() > 0.3 and rel_l2(gt2[both2], tfil[both2]) < 1e-3
    assert rel_l2(gd, deform) < 1e-4
    # tau = 2/sqrt(Q) is singular where Q -> 0+: compare where both are defined and Q is not tiny
    both = (tfil > 0) & (gt > 0) & (tfil < 1e6)
    assert both.mean() > 0.3
    assert rel_l2(gt[both], tfil[both]) < 1e-3
    cmin, cmax = float(v0.min()) - 1e-6, float(v0.max()) * 1.01
    a_g, g_g = b.keff_hist(64, cmin, cmax)
    a_o, g_o = o.keff_hist(64, cmin, cmax)
    assert abs(a_g.sum() - 600000.0 ** 2) < 1e-3 * 600000.0 ** 2
    assert rel_l2(a_g, a_o) < 1e-3
    assert rel_l2(g_g, g_o) < 1e-3
    b.close()


def test_invert_pres(xfb, orc):
    n = 256
    b = xfb.Backend(n)
    o = orc.Oracle(n)
    b.set_vorticity(fields.elliptic(n))
    psi = b.get_field(xfb.capi.PSI)
    got = b.invert_pres(psi, 3, 5)
    exp = o.invert_pres(psi, 3, 5)
    assert rel_l2(got, exp) < 1e-5
    assert got[5, 3] == 0.0       # pres[ref_x + XPTS*ref_y]
    b.close()


def test_error_paths(xfb):
    with pytest.raises(xfb.XfbError):
        xfb.Backend(254)                      # unsupported size: 2 * 127 (300 = 2^2 3 5^2 runs on the generic path)
    b = xfb.Backend(256)
    with pytest.raises(xfb.XfbError):
        b.step(1, 3.0)                        # step before set_vorticity
    with pytest.raises(xfb.XfbError):
        b.set_vorticity(np.zeros((256, 256), np.float32), member=1)
    b.close()


@pytest.mark.parametrize("n", [4096])
def test_large_grid_properties(xfb, n):
    # size-independent properties at bench scale: round trip, mode (0,0) constant, enstrophy decays
    rng = np.random.default_rng(3)
    b = xfb.Backend(n)
    f = rng.standard_normal((n, n)).astype(np.float32)
    back = b.c2r(b.r2c(f)) / np.float32(n * n)
    assert rel_l2(back, f) < 2e-6
    v0 = fields.const_vortex(n)
    b.set_vorticity(v0)
    z0 = b.get_spectrum()
    b.step(3, 3.0)
    z1 = b.get_spectrum()
    assert z1[0, 0] == z0[0, 0]                                     # KAT-5: mean vorticity constant
    w = np.ones(n // 2 + 1); w[1:-1] = 2
    ens0 = float((np.abs(z0.astype(np.complex128)) ** 2 * w).sum())
    ens1 = float((np.abs(z1.astype(np.complex128)) ** 2 * w).sum())
    assert ens1 <= ens0 * (1 + 1e-6)
    b.close()


def test_async_record_fields_match_blocking_reads(xfb):
    """xfb_get_field_async: fields travel to pinned buffers on a second stream while stepping continues; the values
    are those of the state at the time of the call"""
    import ctypes as C
    n = 512
    b = xfb.Backend(n)
    b.set_vorticity(fields.kuo2004(n))
    b.step(2, 3.0)
    want = {w: b.get_field(w) for w in (xfb.capi.VORT, xfb.capi.PSI, xfb.capi.U, xfb.capi.V, xfb.capi.SRC)}
    L = b._L
    bufs, tickets = {}, {}
    for w in want:
        p = C.c_void_p()
        assert L.xfb_host_alloc(C.byref(p), n * n) == 0
        bufs[w] = p
        tickets[w] = b.get_field_async(w, p.value)
    b.step(5, 3.0)                                      # overlaps the copies; must not change what was recorded
    for w in want:
        b.wait_field(tickets[w])
        got = np.ctypeslib.as_array(C.cast(bufs[w], C.POINTER(C.c_float)), shape=(n, n))
        assert np.array_equal(got, want[w]), w
    # more requests than slots: the ring recycles
    p0 = bufs[xfb.capi.VORT]
    ts = [b.get_field_async(xfb.capi.VORT, p0.value) for _ in range(20)]
    b.wait_field(ts[-1])
    b.wait_field(ts[0])
    with pytest.raises(xfb.XfbError):
        b.wait_field(10 ** 6)
    for p in bufs.values():
        assert L.xfb_host_free(p) == 0
    b.close()
