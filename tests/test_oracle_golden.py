"""CPU: the NumPy oracle (oracle/oracle_np.py) against the reference-generated golden fixtures."""
import numpy as np
import pytest

from conftest import force_rel_err, golden_names, load_golden, oracle_from_golden

SLOW = {"lj_3d_n4096"}


@pytest.mark.parametrize("name", [n for n in golden_names()])
def test_numpy_oracle_matches_reference_rows(name):
    g = load_golden(name)
    steps = int(g["sm_steps"])
    o = oracle_from_golden(g)
    rows = o.run(steps, record=True, record_acc=True)
    keep = [int(k) for k in g["keep_rows"]]
    # energies over the whole run
    np.testing.assert_allclose(rows["pe"], g["pot_energy"], rtol=1e-10, atol=1e-13)
    np.testing.assert_allclose(rows["ke"], g["kin_energy"], rtol=1e-10, atol=1e-13)
    np.testing.assert_allclose(rows["temp"], g["temp"], rtol=1e-12, atol=0)
    L = float(g["boxlen"])
    for n, k in enumerate(keep):
        assert np.abs(rows["r"][k] - g["r_rows"][n]).max() <= 1e-9 * L
        assert np.abs(rows["v"][k] - g["v_rows"][n]).max() <= 1e-9 * max(1.0, np.abs(g["v_rows"][n]).max())
        if "acc_rows" in g:
            assert force_rel_err(rows["acc"][k], g["acc_rows"][n]) <= 1e-10


def test_row_semantics_readme():
    """Row 0 is the state AFTER the first step; temp stays sm.TEMP; positions may sit just outside [0, L]."""
    g = load_golden("c1_readme_2d_n36")
    assert not np.array_equal(g["r_rows"][0], g["r0"])
    assert np.all(g["temp"] == 1.0)
    assert g["pot_energy"].shape == (500,)
    E = g["pot_energy"] + g["kin_energy"]
    assert (E.max() - E.min()) / abs(E.mean()) < 1e-3
