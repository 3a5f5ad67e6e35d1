"""CPU: the C oracle (oracle/lj_oracle.c) -- pinned to the reference golden rows through the all-pairs limit,
and self-consistent between its brute-force definition and its cell-list implementation."""
import numpy as np
import pytest

from conftest import load_golden, oracle_from_golden
from oracle import oracle_c


def _random_state(n, L, seed):
    rng = np.random.default_rng(seed)
    return rng.uniform(-0.05 * L, 1.05 * L, size=(3, n))  # slightly outside the box on purpose


@pytest.mark.parametrize("name", ["lj_3d_n64", "lj_3d_n125_odd", "lj_3d_n512", "lj_eps_sigma_mass"])
def test_c_oracle_reproduces_reference_in_the_all_pairs_limit(name):
    """rc >= sqrt(3) L / 2 puts every min-image pair inside the cutoff: the cut-off code path must then give the
    reference's own rows (this is what pins lj_oracle.c to the reference)."""
    g = load_golden(name)
    L = float(g["boxlen"])
    eps = float(g["pot_epsilon"]); sig = float(g["pot_sigma"])
    mass = float(g["sm_mass"]) if "sm_mass" in g else 1.0
    s = oracle_c.CutoffSystem(g["r0"], g["v0"], L, rc=L, skin=0.0, dt=float(g["dt"]), eps=eps, sigma=sig, mass=mass)
    steps = min(int(g["sm_steps"]), 60)
    rows = s.run(steps, record=True)
    np.testing.assert_allclose(rows["pe"], g["pot_energy"][:steps], rtol=1e-10)
    np.testing.assert_allclose(rows["ke"], g["kin_energy"][:steps], rtol=1e-10)
    for n, k in enumerate(int(k) for k in g["keep_rows"]):
        if k < steps:
            assert np.abs(rows["r"][k] - g["r_rows"][n]).max() < 1e-9 * L
    n = g["r0"].shape[1]
    assert s.counters()["pairs_unique"] == 2 * steps * n * (n - 1) // 2
    s.close()


def test_numpy_and_c_oracles_agree_with_a_cutoff():
    g = load_golden("lj_3d_n512")
    L = float(g["boxlen"])
    o = oracle_from_golden(g, r=g["r_rows"][-1], v=g["v_rows"][-1], r_cutoff=2.5)
    s = oracle_c.CutoffSystem(g["r_rows"][-1], g["v_rows"][-1], L, rc=2.5, skin=0.3, dt=float(g["dt"]))
    rows_np = o.run(25)
    rows_c = s.run(25, record=True)
    np.testing.assert_allclose(rows_c["pe"], rows_np["pe"], rtol=1e-11)
    np.testing.assert_allclose(rows_c["ke"], rows_np["ke"], rtol=1e-11)
    assert np.abs(rows_c["r"][-1] - rows_np["r"][-1]).max() < 1e-11 * L
    assert s.counters()["pairs_unique"] == o.pairs_evaluated
    s.close()


@pytest.mark.parametrize("n,L,rl", [(300, 9.0, 2.8), (1000, 14.0, 2.8), (2000, 12.0, 3.1), (64, 6.0, 2.8)])
def test_cell_list_equals_brute_force_definition(n, L, rl):
    r = _random_state(n, L, seed=n)
    off_b, idx_b = oracle_c.neighbors(r, L, rl, brute=True)
    off_c, idx_c = oracle_c.neighbors(r, L, rl, brute=False)
    np.testing.assert_array_equal(off_b, off_c)
    np.testing.assert_array_equal(idx_b, idx_c)
    # symmetric and ascending
    for i in range(0, n, max(1, n // 50)):
        row = idx_b[off_b[i]:off_b[i + 1]]
        assert np.all(np.diff(row) > 0)
        for j in row[:5]:
            assert i in idx_b[off_b[j]:off_b[j + 1]]


def test_canonical_cells():
    L, rl = 10.0, 2.8
    r = np.array([[0.0, 9.999999, -0.1, 10.2, 3.34, 3.33], [0.0] * 6, [0.0, 0.0, 0.0, 0.0, 9.99, 5.0]])
    cell_of, order, nc = oracle_c.cells(r, L, rl)
    assert nc == 3
    # x cells: 0, 2, 2 (wrapped from -0.1), 0 (wrapped from 10.2), 1 (3.34 > 10/3), 0
    np.testing.assert_array_equal(cell_of % 3, [0, 2, 2, 0, 1, 0])
    np.testing.assert_array_equal(cell_of // 9, [0, 0, 0, 0, 2, 1])
    # stable by (cell, id)
    keys = cell_of[order].astype(np.int64) * 10 + order
    assert np.all(np.diff(keys) > 0)


def test_skin_trigger_and_rebuild_counter():
    g = load_golden("lj_3d_n512")
    L = float(g["boxlen"])
    s = oracle_c.CutoffSystem(g["r_rows"][-1], g["v_rows"][-1], L, rc=2.5, skin=0.3, dt=float(g["dt"]))
    s.run(60)
    c = s.counters()
    assert 2 <= c["rebuilds"] <= 30 and c["force_evals"] == 120
    s.close()


def test_c_oracle_with_hard_walls_matches_the_numpy_restatement():
    """lj_oracle.c with walls = the padded-period construction of the product (cells / lists / minimum image at period
    L + 2 (rc + skin), clamp + reflect instead of the wrap): pinned to oracle_np (pbc=False, same plain cutoff), which is pinned
    to the reference-generated hard-wall rows (gravity_hbc_3d, initsetup_boundedmie_hbc)."""
    import numpy as np

    from oracle import oracle_c
    from oracle.oracle_np import MiePotential, OracleSolver, box_length, default_dt, lattice_positions, mb_velocities

    n = 6 ** 3
    L = box_length(3, n, 0.5, 0.15)
    np.random.seed(5)
    r0 = lattice_positions(3, n, L)
    v0 = mb_velocities(3, n, 1.0, 1.0, 1.0) + np.random.default_rng(5).normal(scale=0.9, size=(3, n))
    dt = default_dt(0.5, v0)
    steps = 60
    ref = OracleSolver(r=r0, v=v0, boxlen=L, dt=dt, pot=MiePotential(), pbc=False, radius=0.5, temp=1.0, kb=1.0, r_cutoff=2.5).run(steps)
    o = oracle_c.CutoffSystem(r0, v0, L, 2.5, 0.3, dt, walls=(0.5, 1.0))
    assert o.period == L + 2.0 * 2.8
    out = o.run(steps, record=True)
    np.testing.assert_allclose(out["pe"], ref["pe"], rtol=1e-10, atol=1e-13)
    np.testing.assert_allclose(out["ke"], ref["ke"], rtol=1e-10)
    assert np.abs(out["r"] - np.array(ref["r"])).max() <= 1e-9 * L
    ref_r = np.array(ref["r"])
    assert int(((ref_r < 0.5) | (ref_r > L - 0.5)).sum()) > 0, "no particle reached a wall in this run"
    # the list in force is the non-periodic one: no pair across the period
    off, idx = o.neighbor_list()
    r, _, _ = o.state()
    i = np.repeat(np.arange(n), np.diff(off))
    d = r[:, idx] - r[:, i]
    assert (np.sqrt((d * d).sum(axis=0)) < 2.8 + 0.3).all()   # built within rc + skin, moved by < skin / 2 each since
    o.close()
