"""GPU: initial conditions generated on the device (csrc/gen.cu).  Opt-in and NOT a parity mode for
the velocities (a counter-based generator replaces the global NumPy stream); what is checked is what the mode promises:
the reference's lattice bit for bit, the reference's Maxwell-Boltzmann sampling scheme (inverse CDF on gen.py's own
table; direction factors incl. the whole-array np.prod of gen.py:83-92), determinism per (seed, particle id) regardless
of the path or the number of ranks, and that a state generated in place behaves exactly like the same state uploaded."""
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

STEP = 0.001


def _ctx(n, L, ndim=3, rc=0.0, **kw):
    from mdsea_b200 import _capi

    return _capi.Context(ndim=ndim, num_particles=n, boxlen=L, mass=1.0, radius=0.5, dt=0.004, pbc=1, algorithm=0, pot_kind=1,
                         precision=0, epsilon=1.0, sigma=1.0, mie_m=12.0, mie_n=6.0, mie_a=0.0, r_cutoff=rc, skin=0.3,
                         force_evals_per_step=1, ring_rows=10, device=0, **kw)


def _box(ndim, n, vf=0.3):
    from mdsea_b200.helpers import nsphere_volume

    return (n * nsphere_volume(ndim, 0.5) / vf) ** (1 / ndim)


def _ks_against_table(speeds, cdf):
    """max |empirical CDF - table CDF| of the sampled speeds"""
    s = np.sort(speeds)
    table = np.interp(s, np.arange(cdf.shape[0]) * STEP, cdf)
    emp_hi = np.arange(1, s.shape[0] + 1) / s.shape[0]
    emp_lo = np.arange(0, s.shape[0]) / s.shape[0]
    return max(np.abs(emp_hi - table).max(), np.abs(emp_lo - table).max())


@pytest.mark.parametrize("ndim,per_row", [(1, 4096), (2, 64), (3, 16)])
def test_lattice_bit_exact_and_speed_distribution(ndim, per_row):
    from mdsea_b200.gen import PosGen, mb_cdf

    n = per_row ** ndim
    L = _box(ndim, n)
    cdf = mb_cdf(1.0, 0.8, 1.0)
    ctx = _ctx(n, L, ndim=ndim)
    mean_speed = ctx.generate_state(cdf, STEP, seed=3, isotropic=True)
    r, v, _, _ = ctx.get_state()
    np.testing.assert_array_equal(r, PosGen(ndim=ndim, boxlen=L, nparticles=n).simplecubic())   # bit for bit
    speeds = np.sqrt((v ** 2).sum(axis=0))
    assert mean_speed == pytest.approx(speeds.mean(), rel=1e-12)
    if ndim > 1:   # |v| is the sampled speed itself (1-D: |s cos(theta)|)
        assert _ks_against_table(speeds, cdf) < 1.95 / np.sqrt(n)   # Kolmogorov-Smirnov at alpha ~ 0.1 %
        assert abs(v.mean(axis=1)).max() < 5.0 * np.sqrt(0.8 / n)   # no drift: components have variance k T / m
        np.testing.assert_allclose((v ** 2).mean(axis=1), 0.8 * np.ones(ndim) * (3.0 / ndim if ndim == 2 else 1.0), rtol=0.12)
    else:
        assert speeds.max() <= 50.0 and 0.3 < speeds.mean() < 1.5
    # deterministic per seed, different across seeds
    ctx.generate_state(cdf, STEP, seed=3, isotropic=True)
    np.testing.assert_array_equal(ctx.get_state()[1], v)
    ctx.generate_state(cdf, STEP, seed=4, isotropic=True)
    assert np.abs(ctx.get_state()[1] - v).max() > 0.1
    # temp == 0 (gen.py:100-103): NULL table -> zero velocities, mean speed 0
    assert ctx.generate_state(None, STEP) == 0.0
    assert np.all(ctx.get_state()[1] == 0.0)
    ctx.close()


def test_reference_direction_factors_in_3d():
    """gen.py:83-92: vx = s cos(theta_0); vy, vz carry the product of sin(theta_0) over ALL particles."""
    from mdsea_b200.gen import mb_cdf

    cdf = mb_cdf(1.0, 1.0, 1.0)
    for per_row, tiny in ((2, False), (16, True)):
        n = per_row ** 3
        ctx = _ctx(n, _box(3, n))
        ctx.generate_state(cdf, STEP, seed=1, isotropic=False)
        v = ctx.get_state()[1]
        ctx.generate_state(cdf, STEP, seed=1, isotropic=True)
        s = np.sqrt((ctx.get_state()[1] ** 2).sum(axis=0))   # the same draws: isotropic |v| is the speed itself
        assert np.all(np.abs(v[0]) <= s * (1 + 1e-12))
        ratio = np.sqrt(v[1] ** 2 + v[2] ** 2) / np.sqrt(s ** 2 - 0.0)
        if tiny:
            assert np.all(v[1] == 0.0) and np.all(v[2] == 0.0)     # the product of 4096 sines underflows
        else:
            assert np.allclose(ratio, ratio[0], rtol=1e-12) and 0.0 < ratio[0] < 1.0   # one common factor for all particles
        ctx.close()


def test_generated_state_equals_uploaded_state_on_both_paths():
    from mdsea_b200.gen import mb_cdf

    cdf = mb_cdf(1.0, 1.0, 1.0)
    n = 12 ** 3
    L = _box(3, n)
    for rc in (0.0, 2.5):
        a = _ctx(n, L, rc=rc)
        a.set_temperature(1.0)
        ms = a.generate_state(cdf, STEP, seed=11, isotropic=True)
        r, v, _, _ = a.get_state()
        b = _ctx(n, L, rc=rc)
        b.set_temperature(1.0)
        b.set_state(r, v)
        ra, rb = a.run(10, kb=1.0), b.run(10, kb=1.0)
        for k in ("pe", "ke", "r", "v"):
            np.testing.assert_array_equal(ra[k], rb[k])
        assert ms > 0
        a.set_dt(0.002)   # takes effect for the following run calls
        a.close(); b.close()


def test_generated_state_does_not_depend_on_the_number_of_ranks():
    from mdsea_b200 import _capi
    from mdsea_b200.gen import mb_cdf

    cdf = mb_cdf(1.0, 1.0, 1.0)
    n = 12 ** 3
    L = _box(3, n)
    one = _ctx(n, L, rc=2.5)
    ms1 = one.generate_state(cdf, STEP, seed=5, isotropic=False)
    r1, v1, _, _ = one.get_state()
    one.set_temperature(1.0)
    ref = one.run(10, kb=1.0)
    one.close()
    out, errs = [None, None], []

    def worker(rank):
        try:
            ctx = _ctx(n, L, rc=2.5, rank=rank, world=2, proc_grid=(2, 1, 1))
            ctx.comm_init_local(4242)
            ctx.set_temperature(1.0)
            ms = ctx.generate_state(cdf, STEP, seed=5, isotropic=False)
            rows = ctx.run(10, kb=1.0)
            out[rank] = (ms, rows["pe"].copy(), rows["ke"].copy(), ctx.get_state())
            ctx.close()
        except Exception as e:  # pragma: no cover
            errs.append((rank, repr(e)))

    th = [threading.Thread(target=worker, args=(k,)) for k in range(2)]
    [t.start() for t in th]
    [t.join(timeout=300) for t in th]
    assert not errs, errs
    for ms, pe, ke, _ in out:
        assert ms == ms1
        np.testing.assert_allclose(pe, ref["pe"], rtol=1e-10)
        np.testing.assert_allclose(ke, ref["ke"], rtol=1e-10)
    owned = np.concatenate([o[3][3] for o in out])
    assert np.array_equal(np.sort(owned), np.arange(n))
    del _capi


def test_device_init_through_the_solver(tmp_path, monkeypatch):
    """ContinuousPotentialSolver(sm, ..., device_init=True): lattice, dt from the generated velocities, a normal run."""
    from test_gpu_cutoff import RC, SKIN

    monkeypatch.chdir(tmp_path)
    from mdsea_b200.config import config_context
    from mdsea_b200.constants import fileman as cfm
    from mdsea_b200.core import SysManager
    from mdsea_b200.gen import PosGen
    from mdsea_b200.helpers import get_dt
    from mdsea_b200.potentials import Potential
    from mdsea_b200.simulator import ContinuousPotentialSolver

    base = str(tmp_path / "simfiles")
    for name, suffix in (("DIR_SIMFILES", ""), ("SIM_PATH", "/{}"), ("VIS_PATH", "/{}/vis"), ("PNG_PATH", "/{}/vis/png"),
                         ("MP4_PATH", "/{}/vis/mp4"), ("BLENDER_PATH", "/{}/vis/blender"), ("DATA_PATH", "/{}/data.h5"),
                         ("SETTINGS_PATH", "/{}/settings.pkl")):
        monkeypatch.setattr(cfm, name, base + suffix)
    with config_context(k_boltzmann=1.0):
        sm = SysManager.new(simid="dev", ndim=3, num_particles=14 ** 3, steps=20, vol_fraction=0.3, radius_particle=0.5)
        sim = ContinuousPotentialSolver(sm, pot=Potential.lennardjones(1, 1), r_cutoff=RC, skin=SKIN, device_init=True,
                                        device_init_seed=9)
        np.testing.assert_array_equal(sim.r_vec, PosGen(ndim=3, boxlen=sm.LEN_BOX, nparticles=14 ** 3).simplecubic())
        speeds = np.sqrt((sim.v_vec ** 2).sum(axis=0))
        assert sim.dt == pytest.approx(get_dt(0.5, float(speeds.mean())), rel=1e-12)
        sim.run_simulation()
    pe, ke = sm.get_ds("pot_energy"), sm.get_ds("kin_energy")
    assert np.all(np.isfinite(pe)) and np.all(np.isfinite(ke)) and pe[-1] < 0
    e = pe + ke
    assert np.abs(e - e[0]).max() < 1e-2 * abs(e[0])
