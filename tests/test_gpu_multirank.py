"""GPU: the spatial domain decomposition (K6) on ONE GPU through the in-process transport: `world` contexts, one
host thread each, exchanging ghosts / migrants with device-to-device copies.  The NCCL transport shares every
kernel and all of the routing logic; only the four primitives of csrc/comm.cu differ (exercised by bench.py under
torchrun on real multi-GPU boxes).

Contracts: the union of the ranks' owned particles is the full system at every step; per-rank neighbour lists are
the canonical (global) lists of their owned particles, bit for bit; energies / trajectories equal the C oracle's
(and hence the single-GPU path's) to the fp64 tolerances; all ranks rebuild on the same steps as the oracle.
"""
import threading

import numpy as np
import pytest

from oracle import oracle_c
from test_gpu_cutoff import RC, SKIN, _liquid

pytestmark = pytest.mark.gpu

_KEY = [1000]


def _run_ranks(world, grid, n, L, dt, r, v, steps, precision=0, evals=2, batch=None, want_lists=False, quench=None, g=0.0,
               pot=None, pbc=1):
    from mdsea_b200 import _capi

    _KEY[0] += 1
    key = _KEY[0]
    batch = batch or steps
    out = [None] * world
    errs = []

    def worker(rank):
        try:
            pk = dict(epsilon=1.0, sigma=1.0, mie_m=12.0, mie_n=6.0, mie_a=0.0)
            pk.update(pot or {})
            ctx = _capi.Context(ndim=3, num_particles=n, boxlen=L, mass=1.0, radius=0.5, dt=dt, pbc=pbc, algorithm=0, pot_kind=1,
                                precision=precision, r_cutoff=RC,
                                skin=SKIN, force_evals_per_step=evals, ring_rows=batch, device=0, rank=rank, world=world,
                                proc_grid=grid, **pk)
            ctx.comm_init_local(key)
            ctx.set_temperature(1.0)
            ctx.set_state(r, v)
            res = dict(pe=[], ke=[], temp=[], rows=[])
            for b in range(steps // batch):
                qb = None if quench is None else quench[b * batch:(b + 1) * batch]
                if qb is not None and np.isnan(qb).all():
                    qb = None
                rows = ctx.run(batch, kb=1.0, g=g, quench=qb)
                res["pe"].append(rows["pe"].copy()); res["ke"].append(rows["ke"].copy()); res["temp"].append(rows["temp"].copy())
                res["rows"].append({k: rows[k].copy() for k in ("r", "v", "ids", "n_owned")})
            res["pe"] = np.concatenate(res["pe"]); res["ke"] = np.concatenate(res["ke"]); res["temp"] = np.concatenate(res["temp"])
            res["stats"] = ctx.stats()
            if want_lists:
                res["cells"] = ctx.get_cells()
                res["nbrs"] = ctx.get_neighbors()
            res["state"] = ctx.get_state()
            out[rank] = res
            ctx.close()
        except Exception as e:  # pragma: no cover
            errs.append((rank, repr(e)))

    th = [threading.Thread(target=worker, args=(k,)) for k in range(world)]
    [t.start() for t in th]
    [t.join(timeout=300) for t in th]
    assert not errs, errs
    assert all(o is not None for o in out)
    return out


def _assemble_rows(out, steps, n, batch):
    """Scatter the ranks' compact rows by particle id -> (steps, 3, n); checks every particle is owned exactly once."""
    r = np.full((steps, 3, n), np.nan)
    v = np.full((steps, 3, n), np.nan)
    owners = np.zeros((steps, n), dtype=int)
    for o in out:
        for b, rows in enumerate(o["rows"]):
            for k in range(batch):
                m = int(rows["n_owned"][k])
                ids = rows["ids"][k, :m]
                r[b * batch + k][:, ids] = rows["r"][k, :, :m]
                v[b * batch + k][:, ids] = rows["v"][k, :, :m]
                owners[b * batch + k, ids] += 1
    assert np.all(owners == 1)
    return r, v


@pytest.mark.parametrize("world,grid", [(2, (2, 1, 1)), (2, (1, 1, 2)), (4, (2, 2, 1)), (8, (2, 2, 2))])
def test_decomposed_run_matches_oracle(world, grid):
    n, L, r, v, dt = _liquid(16)
    steps, batch = 40, 20
    o = oracle_c.CutoffSystem(r, v, L, RC, SKIN, dt)
    ref = o.run(steps, record=True)
    oc = o.counters()
    out = _run_ranks(world, grid, n, L, dt, r, v, steps, precision=0, evals=2, batch=batch, want_lists=True)
    for res in out:  # energies are global on every rank
        np.testing.assert_allclose(res["pe"], ref["pe"], rtol=1e-10)
        np.testing.assert_allclose(res["ke"], ref["ke"], rtol=1e-10)
        assert res["stats"]["list_rebuilds"] == oc["rebuilds"]
    rr, vv = _assemble_rows(out, steps, n, batch)
    assert np.abs(rr - ref["r"]).max() <= 1e-9 * L
    assert np.abs(vv - ref["v"]).max() <= 1e-9
    assert sum(res["stats"]["pair_evals_unique"] for res in out) * 1.0 == pytest.approx(oc["pairs_unique"], rel=0, abs=world * steps)
    assert sum(res["stats"]["n_local"] for res in out) == n
    assert all(res["stats"]["n_ghost"] > 0 for res in out)
    # the neighbour lists in force at the end: union over ranks == the oracle's list, entry for entry
    ref_off, ref_idx = o.neighbor_list()
    seen = np.zeros(n, dtype=int)
    for res in out:
        cell_of, order, nc = res["cells"]
        off, idx = res["nbrs"]
        owned = np.where(cell_of >= 0)[0]
        seen[owned] += 1
        for i in owned[:: max(1, len(owned) // 400)]:
            np.testing.assert_array_equal(idx[off[i]:off[i + 1]], ref_idx[ref_off[i]:ref_off[i + 1]])
        cnt = np.diff(off)
        np.testing.assert_array_equal(cnt[owned], np.diff(ref_off)[owned])
        assert cnt[np.setdiff1d(np.arange(n), owned)].sum() == 0
    assert np.all(seen == 1)
    # final state: each rank fills its owned entries
    r_fin = np.full((3, n), np.nan)
    for res in out:
        rr_k, vv_k, aa_k, ids = res["state"]
        r_fin[:, ids] = rr_k[:, ids]
    assert np.abs(r_fin - ref["r"][-1]).max() <= 1e-9 * L
    o.close()


def test_decomposed_fp32_reuse_mode():
    n, L, r, v, dt = _liquid(16)
    steps = 40
    o = oracle_c.CutoffSystem(r, v, L, RC, SKIN, dt)
    ref = o.run(steps, record=True)
    out = _run_ranks(4, (2, 1, 2), n, L, dt, r, v, steps, precision=1, evals=1, batch=steps)
    for res in out:
        np.testing.assert_allclose(res["pe"], ref["pe"], rtol=1e-5)
        np.testing.assert_allclose(res["ke"], ref["ke"], rtol=1e-5)
        assert res["stats"]["force_evals"] == steps + 1
    rr, vv = _assemble_rows(out, steps, n, steps)
    assert np.abs(rr[-1] - ref["r"][-1]).max() <= 1e-3
    o.close()


@pytest.mark.parametrize("world,grid", [(2, (2, 1, 1)), (8, (2, 2, 2)), (2, (1, 1, 2)), (4, (1, 2, 2)), (8, (1, 2, 4))])
def test_decomposed_fp32_lists_bit_exact(world, grid):
    """fp32 pair math with world > 1.  Grids that leave x whole (the default): tile lists for every owned slot, windows
    reaching into the ghost cells.  Grids that split x: tile lists for the whole blocks of interior slots, per-particle
    lists for the rest (domain faces).  One step from a melted state: the lists in force are those of the initial
    positions.  32^3 particles (13 cells per dimension) so that every rank has cells to tile."""
    n, L, r, v, dt = _liquid(32, melt_steps=20)
    out = _run_ranks(world, grid, n, L, dt, r, v, 1, precision=1, evals=1, batch=1, want_lists=True)
    ref_off, ref_idx = oracle_c.neighbors(r, L, RC + SKIN)
    seen = np.zeros(n, dtype=int)
    for res in out:
        assert res["stats"]["list_rebuilds"] == 1
        assert res["stats"]["tile_lists"] == 1 and res["stats"]["tile_fallbacks"] == 0, res["stats"]
        cell_of, order, nc = res["cells"]
        off, idx = res["nbrs"]
        owned = np.where(cell_of >= 0)[0]
        seen[owned] += 1
        cnt = np.diff(off)
        np.testing.assert_array_equal(cnt[owned], np.diff(ref_off)[owned])
        mine = np.repeat(cell_of >= 0, np.diff(ref_off))
        np.testing.assert_array_equal(idx, ref_idx[mine])
    assert np.all(seen == 1)


def test_uneven_blocks_and_lattice_start():
    """nc not divisible by the process grid (blocks of different widths) and a perfect-lattice start."""
    n, L, r, v, dt = _liquid(14, melt_steps=0)   # L = 16.86 -> nc = 6 with 2 ranks in x... and 3 along z: widths 2,2,2
    steps = 30
    o = oracle_c.CutoffSystem(r, v, L, RC, SKIN, dt)
    ref = o.run(steps, record=True)
    out = _run_ranks(6, (2, 1, 3), n, L, dt, r, v, steps, evals=2, batch=steps)
    for res in out:
        np.testing.assert_allclose(res["pe"], ref["pe"], rtol=1e-10)
        np.testing.assert_allclose(res["ke"], ref["ke"], rtol=1e-10)
    rr, vv = _assemble_rows(out, steps, n, steps)
    assert np.abs(rr - ref["r"]).max() <= 1e-9 * L
    n2, L2, r2, v2, dt2 = _liquid(15)            # L = 18.06 -> nc = 6; 4 ranks along x -> widths 1,2,1,2
    o2 = oracle_c.CutoffSystem(r2, v2, L2, RC, SKIN, dt2)
    ref2 = o2.run(20, record=True)
    out2 = _run_ranks(4, (4, 1, 1), n2, L2, dt2, r2, v2, 20, evals=1, batch=20)
    for res in out2:
        np.testing.assert_allclose(res["pe"], ref2["pe"], rtol=1e-10)
    o.close(); o2.close()


def test_decomposed_quench_isothermal_gravity_and_mie():
    """world > 1: _quench needs the GLOBAL mean_ke of the previous step (one all-reduced double per step); the
    bounded-Mie list kernels; quench targets in the first step of a batch and in batches that follow plain ones."""
    from oracle.oracle_np import MiePotential, OracleSolver
    from test_gpu_cutoff import _quench_array

    n, L, r, v, dt = _liquid(12)
    steps, batch = 30, 10
    qs, qt = [10, 14, 29], [0.5, 0.8, 0.2]          # step 10 = first step of the second batch (the first batch is plain)
    ref = OracleSolver(r=r, v=v, boxlen=L, dt=dt, pot=MiePotential(), r_cutoff=RC, temp=1.0, kb=1.0, gravity=True,
                       quench_steps=list(qs), quench_temps=list(qt)).run(steps)
    q = _quench_array(steps, 1.0, False, qs, qt)
    out = _run_ranks(4, (2, 2, 1), n, L, dt, r, v, steps, evals=1, batch=batch, quench=q, g=9.80665)
    for res in out:
        np.testing.assert_allclose(res["pe"], ref["pe"], rtol=1e-10)
        np.testing.assert_allclose(res["ke"], ref["ke"], rtol=1e-10)
        np.testing.assert_allclose(res["temp"], ref["temp"], rtol=1e-10)
    rr, vv = _assemble_rows(out, steps, n, batch)
    assert np.abs(rr - ref["r"]).max() <= 1e-9 * L
    assert np.abs(vv - ref["v"]).max() <= 1e-9
    # isothermal + Mie(9, 6) on 2 ranks
    mie = dict(m=9, n=6, a=0.0, epsilon=1.0, sigma=1.0)
    ref2 = OracleSolver(r=r, v=v, boxlen=L, dt=dt, pot=MiePotential(**mie), r_cutoff=RC, temp=1.0, kb=1.0, isothermal=True).run(20)
    out2 = _run_ranks(2, (1, 2, 1), n, L, dt, r, v, 20, evals=2, batch=10, quench=_quench_array(20, 1.0, True, [], []),
                      pot=dict(mie_m=9.0, mie_n=6.0))
    for res in out2:
        np.testing.assert_allclose(res["pe"], ref2["pe"], rtol=1e-10)
        np.testing.assert_allclose(res["ke"], ref2["ke"], rtol=1e-10)
        np.testing.assert_allclose(res["temp"], ref2["temp"], rtol=1e-10)


def test_multirank_argument_errors():
    from mdsea_b200 import _capi

    kw = dict(ndim=3, num_particles=4096, boxlen=19.3, mass=1.0, radius=0.5, dt=0.005, pbc=1, algorithm=0, pot_kind=1, precision=0,
              epsilon=1.0, sigma=1.0, mie_m=12.0, mie_n=6.0, mie_a=0.0, force_evals_per_step=1, ring_rows=4, device=0)
    with pytest.raises(_capi.MdseaError):   # all-pairs runs do not decompose
        _capi.Context(r_cutoff=0.0, skin=0.0, rank=0, world=2, proc_grid=(2, 1, 1), **kw)
    with pytest.raises(_capi.MdseaError):   # grid does not multiply to world
        _capi.Context(r_cutoff=2.5, skin=0.3, rank=0, world=2, proc_grid=(2, 2, 1), **kw)
    ctx = _capi.Context(r_cutoff=2.5, skin=0.3, rank=0, world=2, proc_grid=(2, 1, 1), **kw)
    with pytest.raises(_capi.MdseaError):   # set_state before comm_init
        ctx.set_state(np.zeros((3, 4096)), np.zeros((3, 4096)))
    ctx.close()


@pytest.mark.parametrize("world,grid,precision", [(2, (1, 1, 2), 0), (4, (1, 2, 2), 1), (8, (1, 2, 4), 0)])
def test_decomposed_hard_wall_run_matches_the_numpy_oracle(world, grid, precision):
    """Hard walls on several ranks: the ranks share the padded periodic cell grid (its outer layers are empty); clamp + reflect
    happens on whichever rank owns the particle.  Energies, trajectories and ownership against the NumPy restatement of the
    reference with the same cutoff (gravity on hard-wall cell runs is covered on one rank in test_gpu_cutoff.py)."""
    from oracle.oracle_np import MiePotential, OracleSolver
    from test_gpu_cutoff import _walled_gas

    n, L, r0, v0, dt = _walled_gas(12, vf=0.15)
    steps, batch = 24, 12
    ref = OracleSolver(r=r0, v=v0, boxlen=L, dt=dt, pot=MiePotential(), pbc=False, radius=0.5, temp=1.0, kb=1.0, r_cutoff=RC).run(steps)
    out = _run_ranks(world, grid, n, L, dt, r0, v0, steps, precision=precision, evals=2, batch=batch, pbc=0)
    # 24 steps: in this hot dilute gas the first hard close encounters (around step 25-30) amplify the 1e-16 differences between
    # summation orders by ~1e6 within a few steps (measured: exact to 2e-16 up to step 20, 1.4e-10 .. 2e-9 at step 40, the same
    # on every decomposition) -- the liquids of the other tests hold 1e-10 over 100 steps
    etol, xtol = (1e-10, 1e-9 * L) if precision == 0 else (1e-5, 1e-3)
    for res in out:
        np.testing.assert_allclose(res["pe"], ref["pe"], rtol=etol, atol=etol * 1e-3)
        np.testing.assert_allclose(res["ke"], ref["ke"], rtol=etol)
        assert res["stats"]["list_rebuilds"] >= 1
    rr, vv = _assemble_rows(out, steps, n, batch)
    assert np.abs(rr - np.array(ref["r"])).max() <= xtol
    assert sum(res["stats"]["n_local"] for res in out) == n
