"""CPU: host-side mirror of the reference interface (no compute calls into the CUDA library)."""
import ctypes
import os
import pickle
import re

import numpy as np
import pytest

from conftest import ROOT, golden_names, load_golden


@pytest.fixture()
def in_tmp_cwd(tmp_path, monkeypatch):
    """Simfile paths are cwd-relative and frozen at first import (constants/fileman.py); repoint them."""
    monkeypatch.chdir(tmp_path)
    from mdsea_b200.constants import fileman as cfm

    base = str(tmp_path / "simfiles")
    monkeypatch.setattr(cfm, "DIR_SIMFILES", base)
    monkeypatch.setattr(cfm, "SIM_PATH", base + "/{}")
    monkeypatch.setattr(cfm, "VIS_PATH", base + "/{}/vis")
    monkeypatch.setattr(cfm, "PNG_PATH", base + "/{}/vis/png")
    monkeypatch.setattr(cfm, "MP4_PATH", base + "/{}/vis/mp4")
    monkeypatch.setattr(cfm, "BLENDER_PATH", base + "/{}/vis/blender")
    monkeypatch.setattr(cfm, "DATA_PATH", base + "/{}/data.h5")
    monkeypatch.setattr(cfm, "SETTINGS_PATH", base + "/{}/settings.pkl")
    return tmp_path


@pytest.mark.parametrize("name", golden_names())
def test_initial_conditions_bit_exact(name):
    """gen.py parity: lattice, Maxwell-Boltzmann velocities (global RNG call order, np.prod quirk), LEN_BOX, dt."""
    from mdsea_b200 import quicker
    from mdsea_b200.config import config_context
    from mdsea_b200.gen import PosGen, VelGen
    from mdsea_b200.helpers import get_dt, nsphere_volume

    g = load_golden(name)
    nd, n = int(g["sm_ndim"]), int(g["sm_num_particles"])
    radius, vf = float(g["sm_radius_particle"]), float(g["sm_vol_fraction"])
    L = (n * nsphere_volume(nd, radius) / vf) ** (1 / nd)
    assert L == float(g["boxlen"])
    r0 = PosGen(ndim=nd, boxlen=L, nparticles=n).simplecubic()
    np.testing.assert_array_equal(r0, g["r0"])
    with config_context(k_boltzmann=1.0):
        from mdsea_b200.config import Config

        np.random.seed(int(g["seed"]))
        v0 = VelGen(ndim=nd, nparticles=n).mb(float(g["sm_mass"]) if "sm_mass" in g else 1,
                                               float(g["sm_temp"]) if "sm_temp" in g else 1, Config.k_boltzmann)
    np.testing.assert_array_equal(v0, g["v0"])
    if "solver_dt" not in g:
        assert get_dt(radius, float(quicker.norm(v0, axis=0).mean())) == float(g["dt"])


def test_sysmanager_contract(in_tmp_cwd):
    from mdsea_b200.core import SysManager

    sm = SysManager.new(simid="abc", ndim=2, num_particles=36, steps=5, vol_fraction=0.2, radius_particle=0.5,
                        quench_temps=[0.5], quench_timings=[0.4])
    root = in_tmp_cwd / "simfiles" / "abc"
    assert (root / "data.h5").exists() and (root / "settings.pkl").exists()
    for d in ("vis", "vis/png", "vis/mp4", "vis/blender"):
        assert (root / d).is_dir()
    assert sm.all_dsnames == ["r_coords", "v_coords", "pot_energy", "kin_energy", "temp"]
    assert sm.QUENCH_STEP == [2] and abs(sm.LEN_BOX - 11.889981892818033) < 1e-12
    with open(root / "settings.pkl", "rb") as f:
        st = pickle.load(f)
    assert sorted(st) == sorted(["radius_particle", "num_particles", "vol_fraction", "steps", "ndim", "mass", "temp", "pbc",
                                 "quench_temps", "quench_steps", "quench_timings", "new_sim", "simid"])
    # the file is closed after initfilesys; the first update_ds reopens it and retries (core.py:329-340)
    assert not sm.dfile.id.valid
    sm.update_ds("pot_energy", [1.5], 0)
    sm.update_ds("r_coords", [np.ones((2, 36))], 4)
    assert sm.dfile.id.valid
    sm.update_ds_block("kin_energy", np.arange(3.0), 1)
    sm2 = SysManager.load("abc")
    assert sm2.NUM_PARTICLES == 36 and sm2.STEPS == 5 and not sm2.new_sim
    assert sm2.get_ds("r_coords").shape == (5, 2, 36) and sm2.get_ds("pot_energy")[0] == 1.5
    np.testing.assert_array_equal(sm2.get_ds("kin_energy"), [0, 0, 1, 2, 0])
    np.testing.assert_array_equal(sm2.get_ds("r_coords")[4], 1.0)
    with pytest.raises(SystemExit):
        sm2.initfilesys()
    with pytest.raises(FileNotFoundError):
        SysManager.load("missing")
    sm.delete()
    assert not root.exists()


def test_potentials_match_oracle_and_map_to_device():
    from mdsea_b200.potentials import Potential
    from oracle.oracle_np import MiePotential

    r = np.linspace(0.8, 3.0, 50)
    lj = Potential.lennardjones(0.7, 1.1)
    o = MiePotential(epsilon=0.7, sigma=1.1)
    np.testing.assert_allclose(lj.force(r), o.dvdr(r), rtol=1e-14)
    np.testing.assert_allclose(lj.potential(r), o.energy(r), rtol=1e-14)
    bm = Potential.boundedmie(a=0.2, epsilon=1, sigma=1, m=12, n=6)
    ob = MiePotential(a=0.2)
    np.testing.assert_allclose(bm.force(r), ob.dvdr(r), rtol=1e-14)
    assert abs(lj.potminimum() - 1.1 * 2 ** (1 / 6)) < 1e-6
    assert lj.device_params() == dict(kind=1, epsilon=0.7, sigma=1.1, m=12.0, n=6.0, a=0.0)
    assert Potential.mie(1, 1, 9, 6).device_params()["m"] == 9.0
    assert bm.device_params()["a"] == 0.2
    assert Potential.ideal().device_params()["kind"] == 0 and Potential.ideal().force(r) == 0
    with pytest.raises(NotImplementedError):
        Potential(force=lambda r: -r, potential=lambda r: r * r / 2).device_params()


def test_config_context():
    from mdsea_b200.config import Config, config_context

    kb = Config.k_boltzmann
    assert abs(kb - 1.380649e-23) < 1e-30
    with config_context(k_boltzmann=1.0):
        assert Config.k_boltzmann == 1.0
    assert Config.k_boltzmann == kb
    with pytest.raises(AttributeError):
        Config.update({"nope": 1})


def test_solver_argument_errors_without_gpu(in_tmp_cwd):
    """Error behaviour of the reference that does not need the device."""
    from mdsea_b200.config import config_context
    from mdsea_b200.core import SysManager
    from mdsea_b200.potentials import Potential
    from mdsea_b200.simulator import ContinuousPotentialSolver

    sm = SysManager.new(simid="e", ndim=2, num_particles=16, steps=3, vol_fraction=0.2, radius_particle=0.5)
    with config_context(k_boltzmann=1.0):
        with pytest.raises(KeyError):
            ContinuousPotentialSolver(sm, pot=Potential.lennardjones(1, 1), algorithm="leapfrog")
        sim = ContinuousPotentialSolver(sm, pot=Potential.lennardjones(1, 1), restitution_coeff=0.5)
        with pytest.raises(NotImplementedError):
            sim.advance()
        sim = ContinuousPotentialSolver(sm, pot=Potential(force=lambda r: r, potential=lambda r: r))
        with pytest.raises(NotImplementedError):
            sim.run_simulation()
        with pytest.raises(ValueError):
            ContinuousPotentialSolver(sm, pot=Potential.lennardjones(1, 1), precision="fp16")
    with pytest.raises(AssertionError):  # simplecubic needs a perfect ndim-th power (gen.py:147-148)
        ContinuousPotentialSolver(SysManager.new(simid="f", ndim=2, num_particles=15, steps=1, vol_fraction=0.2,
                                                 radius_particle=0.5, initfilesys=False))


def test_c_abi_library_exports_every_declared_symbol():
    """The shared library builds for sm_100a without a GPU, loads, and exports what include/mdsea_b200.h declares."""
    from mdsea_b200 import _build, _capi

    lib_path = _build.build()
    assert os.path.exists(lib_path)
    lib = ctypes.CDLL(lib_path)
    hdr = open(os.path.join(ROOT, "include", "mdsea_b200.h")).read()
    declared = set(re.findall(r"\b(mdsea_b200_[a-z0-9_]+)\s*\(", hdr))
    assert declared and declared == set(_capi.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name
    lib.mdsea_b200_version.restype = ctypes.c_char_p
    assert b"sm_100a" in lib.mdsea_b200_version()
    # struct layout agreement between the header and the ctypes mirror
    assert ctypes.sizeof(_capi.Params) == 200 and ctypes.sizeof(_capi.RunArgs) == 32 and ctypes.sizeof(_capi.Rows) == 64


def test_product_fails_loudly_without_gpu():
    """No CPU fallback: without a CUDA device, creating a context raises."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from mdsea_b200 import _capi

    with pytest.raises(_capi.MdseaError) as ei:
        _capi.Context(ndim=2, num_particles=4, boxlen=4.0, mass=1.0, radius=0.5, dt=0.01, pbc=1, pot_kind=1,
                      epsilon=1.0, sigma=1.0, mie_m=12.0, mie_n=6.0)
    assert ei.value.status == -2


def test_device_init_needs_the_gpu_at_construction(in_tmp_cwd):
    """device_init=True generates the state on the device: without a GPU the constructor fails loudly (no host fallback),
    while the default keeps generating on the host and needs no GPU until the first device call."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from mdsea_b200 import _capi
    from mdsea_b200.config import config_context
    from mdsea_b200.core import SysManager
    from mdsea_b200.potentials import Potential
    from mdsea_b200.simulator import ContinuousPotentialSolver

    with config_context(k_boltzmann=1.0):
        sm = SysManager.new(simid="devinit", ndim=3, num_particles=27, steps=2, vol_fraction=0.2, radius_particle=0.5)
        sim = ContinuousPotentialSolver(sm, pot=Potential.lennardjones(1, 1))
        assert sim.r_vec.shape == (3, 27) and sim.dt > 0
        with pytest.raises(_capi.MdseaError) as ei:
            ContinuousPotentialSolver(sm, pot=Potential.lennardjones(1, 1), device_init=True)
        assert ei.value.status == -2


def test_quench_targets_follow_the_reference_call_order(in_tmp_cwd):
    """Host logic of apply_special (simulator.py:185-190): per step, column 0 = the isothermal target (temp_init), column 1
    = the scheduled quench of that step (QUENCH_STEP = round(STEPS * timing), targets consumed in order); NaN = no call;
    None when a batch has no target at all.  No GPU needed: the solver only builds the array here."""
    from mdsea_b200.config import config_context
    from mdsea_b200.core import SysManager
    from mdsea_b200.potentials import Potential
    from mdsea_b200.simulator import ContinuousPotentialSolver

    with config_context(k_boltzmann=1.0):
        sm = SysManager.new(simid="q", ndim=2, num_particles=16, steps=40, vol_fraction=0.2, radius_particle=0.5, temp=0.9,
                            quench_temps=[0.5, 0.1], quench_timings=[0.25, 0.75])
        assert sm.QUENCH_STEP == [10, 30]
        sim = ContinuousPotentialSolver(sm, pot=Potential.lennardjones(1, 1))
        assert sim._quench_targets(0, 10) is None                     # steps 0..9: nothing scheduled
        q = sim._quench_targets(10, 10)                               # step 10 is the first step of this batch
        assert q.shape == (10, 2) and q[0, 1] == 0.5 and np.isnan(q[1:]).all() and np.isnan(q[:, 0]).all()
        assert sim._quench_targets(20, 10) is None
        q = sim._quench_targets(30, 10)
        assert q[0, 1] == 0.1 and np.isnan(q[1:, 1]).all()
        sm2 = SysManager.new(simid="q2", ndim=2, num_particles=16, steps=8, vol_fraction=0.2, radius_particle=0.5, temp=0.9,
                             quench_temps=[0.3], quench_timings=[0.5])
        iso = ContinuousPotentialSolver(sm2, pot=Potential.lennardjones(1, 1), isothermal=True)
        q = iso._quench_targets(0, 8)
        assert np.all(q[:, 0] == 0.9) and q[4, 1] == 0.3 and np.isnan(np.delete(q[:, 1], 4)).all()


def test_config_rejects_unknown_fields_and_restores_on_exit():
    from mdsea_b200.config import Config, config_context

    before = Config.snapshot()
    with pytest.raises(AttributeError):
        Config.update({"no_such_field": 1})
    with config_context(k_boltzmann=2.0, gravity_acceleration=1.0) as cfg:
        assert cfg.k_boltzmann == 2.0 and Config.gravity_acceleration == 1.0
    assert Config.snapshot() == before
    assert set(before) == {"k_boltzmann", "gravity_acceleration"}


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "mdsea_b200")):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, fn)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), fn
                assert "/root/reference" not in src, fn


def test_bench_reads_dram_traffic_from_the_tracked_ncu_summaries(tmp_path, monkeypatch):
    """bench.py fills roofline.traffic from profiles/*.txt (written by profiles/summarize.py): read + write bytes of the
    FIRST block of the named kernel, None when the summary or the kernel is missing."""
    import bench

    prof = tmp_path / "profiles"
    prof.mkdir()
    (prof / "s.txt").write_text(
        "# ncu --set full\n\n== void k_other<1>(int)\ndram__bytes_read.sum   1.0 Gbyte\ndram__bytes_write.sum  1.0 Gbyte\n"
        "\n== void k_nbr_force_f32<1, 1>(const uint4 *)\ngpu__time_duration.sum  114.7 us\n"
        "dram__bytes_read.sum                     244.5 Mbyte\ndram__bytes_write.sum                     18.5 Mbyte\n"
        "\n== void k_nbr_force_f32<1, 1>(const uint4 *)\ndram__bytes_read.sum   9 Gbyte\ndram__bytes_write.sum  9 Gbyte\n")
    monkeypatch.setattr(bench, "ROOT", str(tmp_path))
    assert bench.ncu_traffic("s.txt", "k_nbr_force_f32") == pytest.approx(263.0e6)
    assert bench.ncu_traffic("s.txt", "k_missing") is None
    assert bench.ncu_traffic("missing.txt", "k_nbr_force_f32") is None
    # the tracked r01 summaries carry the kernels bench.py asks for
    monkeypatch.undo()
    assert bench.ncu_traffic("r01_ncu_c3_kernels.txt", "k_nbr_force_f32") > 1e8
    assert bench.ncu_traffic("r01_ncu_c3_force_f64.txt", "k_nbr_force_f64") > 1e8
    assert bench.ncu_traffic("r01_ncu_c2_allpairs_f64.txt", "k_allpairs") is not None


def test_shared_datafile_backend_is_the_memory_mapped_one(in_tmp_cwd):
    """Multi-GPU runs open data.h5 from every rank: that always goes through hdf5min (shared memory maps), whether or not
    real h5py is installed -- libhdf5's file locking forbids several non-MPI read-write opens."""
    from mdsea_b200 import hdf5min
    from mdsea_b200.core import SysManager

    sm = SysManager.new(simid="sh", ndim=3, num_particles=8, steps=3, vol_fraction=0.2, radius_particle=0.5)
    a = sm._open_datafile(shared=True)
    assert isinstance(a, hdf5min.File)
    b = SysManager.load("sh")
    b._open_datafile(shared=True)
    a[sm.rcoord_dsname].raw[1][:, [0, 3]] = 7.0     # two "ranks" writing disjoint columns of the same row
    b.dfile[sm.rcoord_dsname].raw[1][:, [1, 2]] = 9.0
    a.close(); b.dfile.close()
    row = SysManager.load("sh").get_ds(sm.rcoord_dsname)[1]
    assert (row[:, [0, 3]] == 7.0).all() and (row[:, [1, 2]] == 9.0).all()
    # a reopen after a close keeps the shared backend
    sm._open_datafile()
    assert isinstance(sm.dfile, hdf5min.File)


def test_bench_row_hashes_detect_any_change_of_a_neighbour_row():
    """bench.py compares neighbour lists of decomposed runs through one 64-bit hash per particle row (all-reduced over the ranks):
    equal lists hash equal, a changed / missing / extra entry or a moved row boundary changes the hash of the rows it touches."""
    import bench

    rng = np.random.default_rng(0)
    n = 500
    cnt = rng.integers(0, 9, size=n)
    cnt[[0, 17, n - 1]] = 0                                   # empty rows, also at the ends
    off = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int64)
    idx = np.concatenate([np.sort(rng.choice(n, size=c, replace=False)) for c in cnt] + [np.empty(0, dtype=np.int64)]).astype(np.int32)
    h = bench.row_hashes(off, idx, n)
    assert h.shape == (n,) and h.dtype == np.uint64
    np.testing.assert_array_equal(h, bench.row_hashes(off.copy(), idx.copy(), n))
    assert (h[cnt == 0] == 0).all()
    k = int(np.flatnonzero(cnt > 1)[3])
    idx2 = idx.copy()
    idx2[off[k]] = (idx2[off[k]] + 1) % n if (idx2[off[k]] + 1) % n not in idx[off[k]:off[k + 1]] else (idx2[off[k]] + 250) % n
    h2 = bench.row_hashes(off, idx2, n)
    assert h2[k] != h[k] and (np.delete(h2, k) == np.delete(h, k)).all()
    # the sum over "ranks" of per-rank hashes (each rank holds only the rows it owns) equals the global hashes
    own = rng.integers(0, 2, size=n).astype(bool)
    parts = []
    for mask in (own, ~own):
        c = np.where(mask, cnt, 0)
        o = np.concatenate([[0], np.cumsum(c)]).astype(np.int64)
        i = np.concatenate([idx[off[p]:off[p + 1]] for p in np.flatnonzero(mask)] + [np.empty(0, dtype=np.int32)]).astype(np.int32)
        parts.append(bench.row_hashes(o, i, n))
    with np.errstate(over="ignore"):
        np.testing.assert_array_equal(parts[0] + parts[1], h)
