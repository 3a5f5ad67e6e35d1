"""TEST INFRASTRUCTURE ONLY.  Generates tests/golden/*.npz by EXECUTING THE UNMODIFIED
REFERENCE (/root/reference/mdsea, through oracle/refshim.py) in the build container.

    cd /root/repo && python oracle/gen_golden.py

The reference has no golden vectors of its own (tests/test_mdsea.py:1-5), so these
files are the pin for both the CPU oracle (oracle/oracle_np.py, oracle/lj_oracle.c)
and the CUDA path.  Each fixture stores the run settings, the reference's initial
state (r0, v0 from gen.py under np.random.seed(seed), dt, LEN_BOX) and the
reference's outputs (dataset rows exactly as SysManager.update_ds received them,
plus sim.acc after selected steps).  Config.k_boltzmann is 1.0 for every case
(the SI default diverges to NaN at step 2).

The reference's step loop body (simulator.py:583-586) is driven one step at a time
here only so that sim.acc can be sampled between steps; `full_run` cases call
run_simulation() itself.
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")

CASES = {
    # name: (sysmanager kwargs, solver kwargs, potential spec, seed, keep_rows, full_run)
    "c1_readme_2d_n36": (
        dict(ndim=2, num_particles=36, steps=500, vol_fraction=0.2, radius_particle=0.5),
        {}, ("lj", dict(epsilon=1, sigma=1)), 0, [0, 1, 2, 99, 499], True),
    "lj_1d_n8": (
        dict(ndim=1, num_particles=8, steps=50, vol_fraction=0.3, radius_particle=0.5),
        {}, ("lj", dict(epsilon=1, sigma=1)), 1, [0, 49], False),
    "lj_3d_n64": (
        dict(ndim=3, num_particles=64, steps=100, vol_fraction=0.3, radius_particle=0.5),
        {}, ("lj", dict(epsilon=1, sigma=1)), 0, [0, 1, 49, 99], False),
    "lj_3d_n125_odd": (
        dict(ndim=3, num_particles=125, steps=60, vol_fraction=0.3, radius_particle=0.5),
        {}, ("lj", dict(epsilon=1, sigma=1)), 3, [0, 59], False),
    "lj_3d_n512": (
        dict(ndim=3, num_particles=512, steps=100, vol_fraction=0.3, radius_particle=0.5),
        {}, ("lj", dict(epsilon=1, sigma=1)), 0, [0, 9, 99], False),
    "lj_3d_n4096": (
        dict(ndim=3, num_particles=4096, steps=3, vol_fraction=0.3, radius_particle=0.5),
        {}, ("lj", dict(epsilon=1, sigma=1)), 0, [0, 2], False),
    "lj_eps_sigma_mass": (
        dict(ndim=3, num_particles=216, steps=40, vol_fraction=0.25, radius_particle=0.6, mass=2.5, temp=0.7),
        {}, ("lj", dict(epsilon=0.8, sigma=1.2)), 5, [0, 39], False),
    "simple_algo_2d": (
        dict(ndim=2, num_particles=49, steps=60, vol_fraction=0.25, radius_particle=0.5),
        dict(algorithm="simple"), ("lj", dict(epsilon=1, sigma=1)), 2, [0, 59], False),
    "isothermal_2d": (
        dict(ndim=2, num_particles=36, steps=40, vol_fraction=0.2, radius_particle=0.5),
        dict(isothermal=True), ("lj", dict(epsilon=1, sigma=1)), 4, [0, 1, 39], False),
    "quench_3d": (
        dict(ndim=3, num_particles=64, steps=40, vol_fraction=0.3, radius_particle=0.5,
             quench_temps=[0.5, 0.1], quench_timings=[0.25, 0.75]),
        {}, ("lj", dict(epsilon=1, sigma=1)), 6, [0, 10, 30, 39], False),
    "gravity_hbc_3d": (
        dict(ndim=3, num_particles=27, steps=60, vol_fraction=0.1, radius_particle=0.5, pbc=False),
        dict(gravity=True), ("lj", dict(epsilon=1, sigma=1)), 7, [0, 59], False),
    "initsetup_boundedmie_hbc": (  # examples/initsetup.py:15-48
        dict(ndim=3, num_particles=27, steps=240, vol_fraction=0.1, radius_particle=0.5, pbc=False, temp=1),
        {}, ("boundedmie", dict(a=0.2, epsilon=1, sigma=1, m=12, n=6)), 8, [0, 239], False),
    "mie_9_6_2d": (
        dict(ndim=2, num_particles=64, steps=50, vol_fraction=0.3, radius_particle=0.5),
        {}, ("mie", dict(epsilon=1, sigma=1, m=9, n=6)), 9, [0, 49], False),
    "ideal_2d": (
        dict(ndim=2, num_particles=16, steps=20, vol_fraction=0.2, radius_particle=0.5),
        {}, ("ideal", {}), 10, [0, 19], False),
    "fixed_dt_t0": (
        dict(ndim=3, num_particles=64, steps=20, vol_fraction=0.45, radius_particle=0.5, temp=0),
        dict(dt=0.002), ("lj", dict(epsilon=1, sigma=1)), 11, [0, 19], False),
}


def make_pot(ref, spec):
    kind, kw = spec
    P = ref.potentials.Potential
    return {"lj": P.lennardjones, "mie": P.mie, "boundedmie": P.boundedmie, "ideal": P.ideal}[kind](**kw)


def run_case(ref, name, sm_kw, solver_kw, potspec, seed, keep, full_run):
    sm_kw = {k: (list(v) if isinstance(v, list) else v) for k, v in sm_kw.items()}
    sm = ref.core.SysManager.new(simid=name, **sm_kw)
    np.random.seed(seed)
    sim = ref.simulator.ContinuousPotentialSolver(sm, pot=make_pot(ref, potspec), **solver_kw)
    out = dict(
        seed=seed, boxlen=sm.LEN_BOX, dt=sim.dt, r0=sim.r_vec.copy(), v0=sim.v_vec.copy(),
        keep_rows=np.array(keep), pot_kind=potspec[0],
    )
    for k, v in sm_kw.items():
        out["sm_" + k] = np.array(v)
    for k, v in solver_kw.items():
        out["solver_" + k] = np.array(v)
    for k, v in potspec[1].items():
        out["pot_" + k] = np.array(v)
    accs = {}
    if full_run:
        sim.run_simulation()
    else:
        for sim.step in range(sim.step, sm.STEPS):  # simulator.py:583-586
            sim.advance()
            sim.update_files()
            if sim.step in keep:
                accs[sim.step] = sim.acc.copy()
    r = sm.get_ds("r_coords")
    v = sm.get_ds("v_coords")
    out["pot_energy"] = sm.get_ds("pot_energy").copy()
    out["kin_energy"] = sm.get_ds("kin_energy").copy()
    out["temp"] = sm.get_ds("temp").copy()
    out["r_rows"] = np.stack([r[i] for i in keep])
    out["v_rows"] = np.stack([v[i] for i in keep])
    if accs:
        out["acc_rows"] = np.stack([accs[i] for i in keep])
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(f"{name}: N={sm.NUM_PARTICLES} steps={sm.STEPS} dt={sim.dt:.6g} L={sm.LEN_BOX:.6g} "
          f"pe[-1]={out['pot_energy'][-1]:.12g} ke[-1]={out['kin_energy'][-1]:.12g}")


def main():
    os.makedirs(OUT, exist_ok=True)
    work = tempfile.mkdtemp(prefix="mdsea_golden_")
    os.chdir(work)  # reference paths are cwd-relative and frozen at import (constants/fileman.py:8)
    from oracle.refshim import load_reference

    ref = load_reference()
    ref.config.Config.update({"k_boltzmann": 1.0})
    only = sys.argv[1:]
    for name, spec in CASES.items():
        if only and name not in only:
            continue
        run_case(ref, name, *spec)


if __name__ == "__main__":
    main()
