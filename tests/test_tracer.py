"""Passive tracer: dc/dt = -u c_x - v c_y + kappa lap(c), advanced inside xfb_step.

The reference has no tracer, so parity is against the oracle's restatement (barotropic_oracle.c get_dtrcdt: the
vorticity tendency's own operations with c for vort and kappa for NU) -- "parity unpinned" by the reference, pinned
by the identity below: a tracer equal to the vorticity with kappa == nu stays equal to the vorticity.
Tolerances as for the vorticity: relative L2 <= 1e-5 after one step.
"""
import numpy as np
import pytest

from conftest import rel_l2
import fields


def _blob(n, cx=0.42, cy=0.55, r=0.08):
    x = (np.arange(n) / n)[:, None]
    y = (np.arange(n) / n)[None, :]
    return np.exp(-((x - cx) ** 2 + (y - cy) ** 2) / r ** 2).astype(np.float32)


def test_oracle_tracer_equal_to_vorticity_stays_equal():
    from oracle import oracle as orc
    n = 64
    v0 = fields.gaussian(n)
    o = orc.Oracle(n)
    o.set_vorticity(v0)
    o.set_tracer(v0, 6.5)
    o.step(5, 3.0)
    assert np.array_equal(o.get_tracer(), o.get_field(orc.VORT))
    ref = orc.Oracle(n)
    ref.set_vorticity(v0)
    ref.step(5, 3.0)
    assert np.array_equal(ref.get_field(orc.VORT), o.get_field(orc.VORT))      # the tracer never feeds back


def test_oracle_tracer_conserves_mean_and_bounds():
    from oracle import oracle as orc
    n = 64
    o = orc.Oracle(n)
    o.set_vorticity(fields.elliptic(n))
    c0 = _blob(n)
    o.set_tracer(c0, 0.0)
    o.step(20, 3.0)
    c = o.get_tracer()
    assert abs(float(c.mean()) - float(c0.mean())) < 1e-6
    assert c.min() > -0.05 and c.max() < 1.05


@pytest.mark.gpu
@pytest.mark.parametrize("n,gen", [(256, "elliptic"), (512, "kuo2004"), (1024, "gaussian")])
def test_tracer_one_step_parity(n, gen):
    import xlab_fftbarotropic_b200 as xfb
    from oracle import oracle as orc
    v0 = fields.GENERATORS[gen](n)
    c0 = _blob(n)
    b = xfb.Backend(n)
    o = orc.Oracle(n)
    b.set_vorticity(v0); o.set_vorticity(v0)
    b.set_tracer(c0, 25.0); o.set_tracer(c0, 25.0)
    assert rel_l2(b.get_field(xfb.capi.TRACER), o.get_tracer()) < 5e-7
    b.step(1, 3.0); o.step(1, 3.0)
    assert rel_l2(b.get_field(xfb.capi.TRACER), o.get_tracer()) < 1e-5
    assert rel_l2(b.get_field(xfb.capi.VORT), o.get_field(orc.VORT)) < 1e-5
    b.close()


@pytest.mark.gpu
def test_tracer_many_steps_graph_replay_and_reset():
    import xlab_fftbarotropic_b200 as xfb
    from oracle import oracle as orc
    n = 256
    v0 = fields.elliptic(n)
    b = xfb.Backend(n)
    o = orc.Oracle(n)
    b.set_vorticity(v0); o.set_vorticity(v0)
    b.step(3, 3.0); o.step(3, 3.0)                 # a captured step WITHOUT tracer launches exists now
    c0 = _blob(n)
    b.set_tracer(c0, 10.0); o.set_tracer(c0, 10.0)
    l0 = b.launch_count
    b.step(40, 3.0); o.step(40, 3.0)               # eager + graph replays with the tracer's 8 extra launches
    assert b.launch_count - l0 == 40 * 16 + 1    # + the tracer's prologue
    assert rel_l2(b.get_field(xfb.capi.TRACER), o.get_tracer()) < 1e-4
    assert rel_l2(b.get_field(xfb.capi.VORT), o.get_field(orc.VORT)) < 1e-4
    # another diffusivity: the captured step must not keep the old one
    b.set_tracer(c0, 400.0); o.set_tracer(c0, 400.0)
    b.step(10, 3.0); o.step(10, 3.0)
    assert rel_l2(b.get_field(xfb.capi.TRACER), o.get_tracer()) < 1e-4
    b.close()


@pytest.mark.gpu
@pytest.mark.parametrize("n", [512, 4096])
def test_tracer_equal_to_vorticity_stays_equal(n):
    """size-independent property: c == zeta, kappa == nu, no forcing  =>  c(t) == zeta(t)"""
    import xlab_fftbarotropic_b200 as xfb
    v0 = fields.elliptic(n)
    dt = 3.0 if n <= 1024 else 1.0
    b = xfb.Backend(n)
    b.set_vorticity(v0)
    b.set_tracer(v0, 6.5)
    b.step(6, dt)
    z = b.get_field(xfb.capi.VORT)
    c = b.get_field(xfb.capi.TRACER)
    assert rel_l2(c, z) < 1e-6
    b.close()


@pytest.mark.gpu
def test_tracer_ensemble_members_and_keff():
    import xlab_fftbarotropic_b200 as xfb
    from oracle import oracle as orc
    n = 256
    b = xfb.Backend(n, batch=2)
    gens = ["elliptic", "gaussian"]
    blobs = [_blob(n), _blob(n, 0.6, 0.4, 0.05)]
    for m in range(2):
        b.set_vorticity(fields.GENERATORS[gens[m]](n), member=m)
        b.set_tracer(blobs[m], 15.0, member=m)
    b.step(4, 3.0)
    for m in range(2):
        o = orc.Oracle(n)
        o.set_vorticity(fields.GENERATORS[gens[m]](n))
        o.set_tracer(blobs[m], 15.0)
        o.step(4, 3.0)
        assert rel_l2(b.get_field(xfb.capi.TRACER, member=m), o.get_tracer()) < 1e-5, m
        a_g, g_g = b.tracer_keff_hist(32, -0.01, 1.01, member=m)
        a_o, g_o = o.tracer_keff_hist(32, -0.01, 1.01)
        assert abs(a_g.sum() - 600000.0 ** 2) < 1e-3 * 600000.0 ** 2
        assert rel_l2(a_g, a_o) < 1e-3 and rel_l2(g_g, g_o) < 1e-3
    b.close()


@pytest.mark.gpu
def test_tracer_error_paths():
    import xlab_fftbarotropic_b200 as xfb
    b = xfb.Backend(256)
    b.set_vorticity(fields.gaussian(256))
    with pytest.raises(xfb.XfbError):
        b.get_field(xfb.capi.TRACER)               # before set_tracer
    with pytest.raises(xfb.XfbError):
        b.tracer_keff_hist(16, 0.0, 1.0)
    b.close()


@pytest.mark.gpu
def test_tracer_on_the_default_768_grid():
    """generic mixed-radix path (the reference's default NPTS = 768): same loops, unfused"""
    import xlab_fftbarotropic_b200 as xfb
    from oracle import oracle as orc
    n = 768
    v0 = fields.elliptic(n)
    c0 = _blob(n)
    b = xfb.Backend(n)
    o = orc.Oracle(n)
    b.set_vorticity(v0); o.set_vorticity(v0)
    b.set_tracer(c0, 20.0); o.set_tracer(c0, 20.0)
    b.step(1, 3.0); o.step(1, 3.0)
    assert rel_l2(b.get_field(xfb.capi.TRACER), o.get_tracer()) < 1e-5
    b.step(9, 3.0); o.step(9, 3.0)
    assert rel_l2(b.get_field(xfb.capi.TRACER), o.get_tracer()) < 1e-5
    assert rel_l2(b.get_field(xfb.capi.VORT), o.get_field(orc.VORT)) < 1e-5
    a_g, g_g = b.tracer_keff_hist(32, -0.01, 1.01)
    a_o, g_o = o.tracer_keff_hist(32, -0.01, 1.01)
    assert rel_l2(a_g, a_o) < 1e-3 and rel_l2(g_g, g_o) < 1e-3
    b.close()


@pytest.mark.gpu
def test_main_out_tracer_files():
    """main.out -c: tracer_step_N.bin after the reference's five record files, against the oracle"""
    import os
    import subprocess
    import tempfile
    from oracle import oracle as orc
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "xlab_fftbarotropic_b200", "bin", "main.out")
    n, steps, rec = 256, 5, 2
    v0 = fields.elliptic(n)
    c0 = _blob(n)
    with tempfile.TemporaryDirectory() as d:
        os.makedirs(os.path.join(d, "input")); os.makedirs(os.path.join(d, "output"))
        v0.tofile(os.path.join(d, "input", "initial_vorticity.bin"))
        c0.tofile(os.path.join(d, "input", "tracer.bin"))
        r = subprocess.run([exe, "-n", str(n), "-t", str(steps), "-r", str(rec), "-c", "tracer.bin", "-k", "12.5"], cwd=d,
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        log = open(os.path.join(d, "log")).read().split()
        assert [os.path.basename(p) for p in log[:6]] == ["vort_src_input_step_0.bin", "vort_step_0.bin", "psi_step_0.bin",
                                                           "u_step_0.bin", "v_step_0.bin", "tracer_step_0.bin"]
        o = orc.Oracle(n)
        o.set_vorticity(v0)
        o.set_tracer(c0, 12.5)
        for s in range(0, steps, rec):
            got = np.fromfile(os.path.join(d, "output", f"tracer_step_{s}.bin"), np.float32).reshape(n, n)
            assert rel_l2(got, o.get_tracer()) < 1e-5, s
            o.step(rec, 3.0)


@pytest.mark.gpu
def test_tracer_on_the_16384_grid():
    """the two-level kernels of 16384-point lines carry the tracer too: a tracer equal to the vorticity with kappa = nu and
    no forcing stays bit-identical to it"""
    import xlab_fftbarotropic_b200 as xfb
    n = 16384
    x = (np.arange(n, dtype=np.float32) / n)
    v0 = (5e-3 * np.exp(-((x[:, None] - 0.5) / 0.07) ** 2 - ((x[None, :] - 0.45) / 0.04) ** 2)).astype(np.float32)
    b = xfb.Backend(n)
    b.set_vorticity(v0)
    b.set_tracer(v0, 6.5)
    b.step(1, 0.25)
    assert np.array_equal(b.get_field(xfb.capi.TRACER), b.get_field(xfb.capi.VORT))
    b.close()
