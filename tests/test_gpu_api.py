"""GPU: the reference-facing Python API end to end -- SysManager.new -> ContinuousPotentialSolver ->
run_simulation -> data.h5 -> SysManager.load -> Analyser -- against the reference's golden rows."""
import numpy as np
import pytest

from conftest import load_golden
from test_host_cpu import in_tmp_cwd  # noqa: F401  (fixture)

pytestmark = pytest.mark.gpu


def _run(kw, solver_kw, pot, seed):
    from mdsea_b200.config import config_context
    from mdsea_b200.core import SysManager
    from mdsea_b200.simulator import ContinuousPotentialSolver

    with config_context(k_boltzmann=1.0):
        sm = SysManager.new(**kw)
        np.random.seed(seed)
        sim = ContinuousPotentialSolver(sm, pot=pot, **solver_kw)
        sim.run_simulation()
    return sm, sim


def test_readme_example_end_to_end(in_tmp_cwd):  # noqa: F811
    from mdsea_b200.analytics import Analyser
    from mdsea_b200.core import SysManager
    from mdsea_b200.potentials import Potential

    g = load_golden("c1_readme_2d_n36")
    sm, sim = _run(dict(simid="_mdsea_docs_example", ndim=2, num_particles=36, steps=500, vol_fraction=0.2,
                        radius_particle=0.5), dict(flush_every=64), Potential.lennardjones(1, 1), 0)
    assert sim.dt == float(g["dt"]) and sim.step == 499
    sm2 = SysManager.load("_mdsea_docs_example")
    pe, ke, temp = sm2.get_ds("pot_energy"), sm2.get_ds("kin_energy"), sm2.get_ds("temp")
    np.testing.assert_allclose(pe[:150], g["pot_energy"][:150], rtol=1e-10)
    np.testing.assert_allclose(ke[:150], g["kin_energy"][:150], rtol=1e-10)
    np.testing.assert_allclose(pe, g["pot_energy"], rtol=1e-6)
    np.testing.assert_array_equal(temp, g["temp"])
    r = sm2.get_ds("r_coords")
    assert r.shape == (500, 2, 36)
    for n, k in enumerate(int(k) for k in g["keep_rows"]):
        if k <= 100:
            assert np.abs(r[k] - g["r_rows"][n]).max() < 1e-9 * sm.LEN_BOX
    # energy drift no worse than the reference's own (7.3e-4 relative spread over the 500 steps)
    E = pe + ke
    Eref = g["pot_energy"] + g["kin_energy"]
    assert (E.max() - E.min()) / abs(E.mean()) <= 1.05 * (Eref.max() - Eref.min()) / abs(Eref.mean())
    an = Analyser(sm2)
    assert an.r_vecs.shape == (500, 36, 2) and an.speeds.shape == (500, 36)
    np.testing.assert_allclose(an.total_energy, E)
    # state attributes after the run
    np.testing.assert_array_equal(sim.r_vec, r[499])
    assert sim.mean_pe == pe[499] and sim.mean_ke == ke[499] and sim.temp == 1.0
    st = sim.stats()
    assert st["steps"] == 500 and st["force_evals"] == 501  # a(t+dt) reused: one evaluation per step + the first


def test_strict_two_evals_and_fp32_modes(in_tmp_cwd):  # noqa: F811
    from mdsea_b200.potentials import Potential

    g = load_golden("lj_3d_n512")
    kw = dict(simid="x", ndim=3, num_particles=512, steps=60, vol_fraction=0.3, radius_particle=0.5)
    sm, sim = _run(kw, dict(force_evals_per_step=2), Potential.lennardjones(1, 1), 0)
    assert sim.stats()["force_evals"] == 120
    np.testing.assert_allclose(sm.get_ds("pot_energy"), g["pot_energy"][:60], rtol=1e-10)
    sm, sim = _run(kw, dict(precision="fp32"), Potential.lennardjones(1, 1), 0)
    np.testing.assert_allclose(sm.get_ds("pot_energy"), g["pot_energy"][:60], rtol=1e-5)
    np.testing.assert_allclose(sm.get_ds("kin_energy"), g["kin_energy"][:60], rtol=1e-5)


def test_hbc_boundedmie_gravity_quench_through_api(in_tmp_cwd):  # noqa: F811
    from mdsea_b200.potentials import Potential

    g = load_golden("initsetup_boundedmie_hbc")  # examples/initsetup.py
    sm, sim = _run(dict(simid="i", ndim=3, num_particles=27, steps=240, vol_fraction=0.1, radius_particle=0.5, pbc=False,
                        temp=1), {}, Potential.boundedmie(a=0.2, epsilon=1, sigma=1, m=12, n=6), 8)
    np.testing.assert_allclose(sm.get_ds("pot_energy")[:100], g["pot_energy"][:100], rtol=1e-10)
    np.testing.assert_allclose(sm.get_ds("kin_energy")[:100], g["kin_energy"][:100], rtol=1e-10)

    g = load_golden("gravity_hbc_3d")
    sm, sim = _run(dict(simid="g", ndim=3, num_particles=27, steps=60, vol_fraction=0.1, radius_particle=0.5, pbc=False),
                   dict(gravity=True), Potential.lennardjones(1, 1), 7)
    np.testing.assert_allclose(sm.get_ds("kin_energy"), g["kin_energy"], rtol=1e-10)

    g = load_golden("quench_3d")
    sm, sim = _run(dict(simid="q", ndim=3, num_particles=64, steps=40, vol_fraction=0.3, radius_particle=0.5,
                        quench_temps=[0.5, 0.1], quench_timings=[0.25, 0.75]), dict(flush_every=7),
                   Potential.lennardjones(1, 1), 6)
    np.testing.assert_allclose(sm.get_ds("temp"), g["temp"], rtol=1e-10)
    np.testing.assert_allclose(sm.get_ds("kin_energy"), g["kin_energy"], rtol=1e-10)

    g = load_golden("isothermal_2d")
    sm, sim = _run(dict(simid="t", ndim=2, num_particles=36, steps=40, vol_fraction=0.2, radius_particle=0.5),
                   dict(isothermal=True, flush_every=16), Potential.lennardjones(1, 1), 4)
    np.testing.assert_allclose(sm.get_ds("temp"), g["temp"], rtol=1e-10)
    np.testing.assert_allclose(sm.get_ds("kin_energy"), g["kin_energy"], rtol=1e-10)


def test_manual_stepping_and_attribute_edits(in_tmp_cwd):  # noqa: F811
    """for sim.step in range(...): sim.advance(); sim.update_files() -- the reference's own loop body."""
    from mdsea_b200.config import config_context
    from mdsea_b200.core import SysManager
    from mdsea_b200.potentials import Potential
    from mdsea_b200.simulator import ContinuousPotentialSolver

    g = load_golden("lj_3d_n64")
    with config_context(k_boltzmann=1.0):
        sm = SysManager.new(simid="m", ndim=3, num_particles=64, steps=10, vol_fraction=0.3, radius_particle=0.5)
        np.random.seed(0)
        sim = ContinuousPotentialSolver(sm, pot=Potential.lennardjones(1, 1), force_evals_per_step=2)
        for sim.step in range(0, 10):
            sim.advance()
            sim.update_files()
        np.testing.assert_allclose(sm.get_ds("pot_energy"), g["pot_energy"][:10], rtol=1e-10)
        acc = sim.update_acc()
        assert acc.shape == (3, 64)
        d = sim.update_dists()
        assert d.shape == (64 * 63 // 2,) and sim.drunits.shape == (64 * 63 // 2, 3)
        # editing the state on the host is picked up by the next device call
        sim.v_vec[:] = 0.0
        sim.advance()
        assert sim.mean_ke < 1e-3


@pytest.mark.parametrize("cutoff", [False, True])
def test_two_batches_in_flight_give_the_same_rows(cutoff):
    """run_async / wait pipelining (two device ring sets): rows and scalars of every batch are identical to the
    ones of the blocking run(), for the all-pairs path (fully asynchronous) and the cut-off path."""
    from mdsea_b200 import _capi
    from oracle.oracle_np import box_length, default_dt, lattice_positions, mb_velocities

    per = 12 if cutoff else 6
    n = per ** 3
    L = box_length(3, n, 0.5, 0.3)
    np.random.seed(3)
    r0, v0 = lattice_positions(3, n, L), mb_velocities(3, n, 1.0, 1.0, 1.0)
    v0 = v0 + np.random.default_rng(3).normal(scale=0.5, size=v0.shape)
    dt = default_dt(0.5, v0)
    kw = dict(ndim=3, num_particles=n, boxlen=L, mass=1.0, radius=0.5, dt=dt, pbc=1, algorithm=0, pot_kind=1, precision=0,
              epsilon=1.0, sigma=1.0, mie_m=12.0, mie_n=6.0, mie_a=0.0, force_evals_per_step=1, ring_rows=8, device=0)
    if cutoff:
        kw.update(r_cutoff=2.5, skin=0.3)
    nb, k = 5, 8

    def fresh():
        ctx = _capi.Context(**kw)
        ctx.set_temperature(1.0)
        ctx.set_state(r0, v0)
        return ctx

    a = fresh()
    ref = [{key: val.copy() for key, val in a.run(k, kb=1.0).items()} for _ in range(nb)]
    ref_pairs = a.stats()["pair_evals_unique"]
    a.close()

    b = fresh()
    bufs = [dict(r=np.empty((k, 3, n)), v=np.empty((k, 3, n)), pe=np.empty(k), ke=np.empty(k), temp=np.empty(k)) for _ in range(2)]
    got, in_flight = [], []
    for i in range(nb):
        b.run(k, kb=1.0, rows=bufs[i % 2], asynchronous=True)
        in_flight.append(i)
        if len(in_flight) == 2:
            b.wait()
            j = in_flight.pop(0)
            got.append({key: val.copy() for key, val in bufs[j % 2].items()})
    while in_flight:
        b.wait()
        j = in_flight.pop(0)
        got.append({key: val.copy() for key, val in bufs[j % 2].items()})
    assert b.stats()["pair_evals_unique"] == ref_pairs
    b.close()
    for x, y in zip(ref, got):
        for key in ("r", "v", "pe", "ke", "temp"):
            np.testing.assert_array_equal(x[key], y[key])


def test_analyser_speed_statistics_on_device(in_tmp_cwd):  # noqa: F811
    """Analyser(sm, device=0) == the reference's NumPy formulation (speeds bit for bit)."""
    from mdsea_b200 import _capi
    from mdsea_b200.analytics import Analyser
    from mdsea_b200.potentials import Potential
    from mdsea_b200.quicker import norm

    sm, _ = _run(dict(simid="an", ndim=3, num_particles=216, steps=40, vol_fraction=0.3, radius_particle=0.5),
                 dict(flush_every=16), Potential.lennardjones(1, 1), 1)
    host, dev = Analyser(sm), Analyser(sm, device=0)
    np.testing.assert_array_equal(dev.speeds, host.speeds)
    assert abs(dev.maxspeed - host.maxspeed) <= 1e-12 * host.maxspeed
    # chunked streaming (several chunks, 1-D / 2-D layouts, no speeds output)
    rng = np.random.default_rng(5)
    for shape in ((3, 1, 1000), (7, 2, 333), (40, 3, 70000)):
        v = rng.normal(size=shape)
        sp, mean, std, mx = _capi.analyse_speeds(v)
        ref = norm(np.moveaxis(v, 1, -1), axis=-1)
        np.testing.assert_array_equal(sp, ref)
        assert abs(mean - ref.mean()) <= 1e-12 * ref.mean() and abs(std - ref.std()) <= 1e-11 * ref.std() and mx == ref.max()
        _, mean2, _, _ = _capi.analyse_speeds(v, want_speeds=False)
        assert mean2 == mean


def test_boundary_and_pair_methods_of_the_base_simulator(in_tmp_cwd):  # noqa: F811
    """apply_bc / apply_pbc / apply_hbc, _quench, apply_field, com, rog, pairs, update_pairs as callable methods
    (mdsea/simulator.py:90,156-231,313-323), each against the NumPy statement of the reference lines."""
    from mdsea_b200 import quicker
    from mdsea_b200.config import Config, config_context
    from mdsea_b200.core import SysManager
    from mdsea_b200.potentials import Potential
    from mdsea_b200.simulator import ContinuousPotentialSolver

    with config_context(k_boltzmann=1.0):
        # hard walls: clamp + reflect on the device, com / rog reduced on the device
        sm = SysManager.new(simid="w", ndim=3, num_particles=27, steps=5, vol_fraction=0.1, radius_particle=0.5, pbc=False, temp=1)
        np.random.seed(3)
        sim = ContinuousPotentialSolver(sm, pot=Potential.lennardjones(1, 1), restitution_coeff=1.0, gravity=True)
        L, a = sm.LEN_BOX, sm.RADIUS_PARTICLE
        r = sim.r_vec.copy()
        v = sim.v_vec.copy()
        r[0, :3] = [-0.2, L + 0.3, 0.1]
        r[2, 5] = L - 0.1
        sim.r_vec = r
        er, ev = r.copy(), v.copy()
        whr = np.where(er - a < 0); er[whr] = a; ev[whr] *= -1.0
        whr = np.where(er + a > L); er[whr] = L - a; ev[whr] *= -1.0
        sim.apply_bc()
        np.testing.assert_array_equal(sim.r_vec, er)
        np.testing.assert_array_equal(sim.v_vec, ev)
        sim.apply_hbc()   # idempotent once inside
        np.testing.assert_array_equal(sim.r_vec, er)
        com = sim.com
        np.testing.assert_allclose(com, er.mean(axis=1), rtol=1e-13)
        d = np.stack(er, axis=-1) - er.mean(axis=1)
        np.testing.assert_allclose(sim.rog, np.sqrt(np.mean(quicker.norm(d, axis=1) ** 2)), rtol=1e-13)
        vz = sim.v_vec[-1].copy()
        sim.apply_field()
        np.testing.assert_allclose(sim.v_vec[-1], vz - Config.gravity_acceleration * sim.dt, rtol=0, atol=0)
        # pairs within a radius: combinations order, the table built on first use
        pr = list(sim.update_pairs(radius=2.0))
        assert sim.all_pairs.shape == (27 * 26 // 2, 2) and tuple(sim.all_pairs[0]) == (0, 1) and tuple(sim.all_pairs[-1]) == (25, 26)
        assert len(pr) == len(sim.pairs) == int(sim.pairs_indexes.sum()) > 0
        (i, j), dist, unit = pr[0]
        rr = sim.r_vec
        assert dist == pytest.approx(np.linalg.norm(rr[:, j] - rr[:, i]), rel=1e-12) and dist < 2.0
        # _quench: v *= sqrt(T / temp) with temp from the current mean kinetic energy
        sim.update_mean_ke()
        v0 = sim.v_vec.copy()
        sim._quench(0.25)
        t_now = (2 / 3) * 0.5 * sm.MASS * float((v0 * v0).sum(axis=0).mean())
        np.testing.assert_allclose(sim.v_vec, v0 * (0.25 / t_now) ** 0.5, rtol=1e-14)

        # periodic: strict one-shot wrap on the device; rog has no centre of mass to refer to (TypeError like the reference)
        sm = SysManager.new(simid="p", ndim=2, num_particles=36, steps=5, vol_fraction=0.2, radius_particle=0.5)
        np.random.seed(4)
        sim = ContinuousPotentialSolver(sm, pot=Potential.lennardjones(1, 1))
        L = sm.LEN_BOX
        r = sim.r_vec.copy()
        r[0, :4] = [-0.5, L + 0.25, 2.5 * L, -1.5 * L]
        sim.r_vec = r
        er = r.copy()
        er[np.where(er < 0)] += L
        er[np.where(er > L)] -= L
        sim.apply_pbc()
        np.testing.assert_array_equal(sim.r_vec, er)
        assert sim.com is None
        with pytest.raises(TypeError):
            sim.rog


@pytest.mark.parametrize("pbc,gravity", [(True, False), (False, True)])
def test_two_dimensional_boxes_take_the_cutoff_path(in_tmp_cwd, pbc, gravity):  # noqa: F811
    """A 2-D system with a cutoff (the README family at sizes the all-pairs kernels cannot hold) runs as the 3-D system (x, L/2, y) on
    the cell / list kernels: every row against the NumPy restatement of the reference (2-D, same plain cutoff)."""
    from mdsea_b200.config import config_context
    from mdsea_b200.core import SysManager
    from mdsea_b200.potentials import Potential
    from mdsea_b200.simulator import ContinuousPotentialSolver
    from oracle.oracle_np import MiePotential, OracleSolver

    steps = 40
    with config_context(k_boltzmann=1.0):
        sm = SysManager.new(simid="d2", ndim=2, num_particles=900, steps=steps, vol_fraction=0.25, radius_particle=0.5, pbc=pbc, temp=1)
        np.random.seed(12)
        sim = ContinuousPotentialSolver(sm, pot=Potential.lennardjones(1, 1), r_cutoff=2.5, gravity=gravity, flush_every=16)
        r0, v0 = sim.r_vec.copy(), sim.v_vec.copy()
        ref = OracleSolver(r=r0, v=v0, boxlen=sm.LEN_BOX, dt=sim.dt, pot=MiePotential(), pbc=pbc, radius=0.5, temp=1.0, kb=1.0,
                           gravity=gravity, g_acc=9.80665, r_cutoff=2.5).run(steps)
        sim.run_simulation()
        st = sim.stats()
        assert st["list_rebuilds"] >= 1 and sim._embed == [0, 2]          # the cell path ran, as a 3-D system
        rr, vv = sm.get_ds("r_coords"), sm.get_ds("v_coords")
        assert rr.shape == (steps, 2, 900)
        np.testing.assert_allclose(sm.get_ds("pot_energy"), ref["pe"], rtol=1e-10, atol=1e-13)
        np.testing.assert_allclose(sm.get_ds("kin_energy"), ref["ke"], rtol=1e-10)
        np.testing.assert_allclose(sm.get_ds("temp"), ref["temp"], rtol=1e-10)
        assert np.abs(rr - np.array(ref["r"])).max() <= 1e-9 * sm.LEN_BOX
        assert np.abs(vv - np.array(ref["v"])).max() <= 1e-9
        # the host views are 2-D as well
        assert sim.r_vec.shape == (2, 900) and sim.acc.shape == (2, 900)
        np.testing.assert_allclose(sim.r_vec, rr[-1], rtol=0, atol=0)
        sim.close()
