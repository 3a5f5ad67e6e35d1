import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
    return {k: z[k] for k in z.files}


def golden_names():
    return sorted(f[:-4] for f in os.listdir(GOLDEN_DIR) if f.endswith(".npz"))


def oracle_from_golden(g, **over):
    """Build an OracleSolver from a golden fixture's stored settings + initial state."""
    from oracle.oracle_np import MiePotential, OracleSolver

    kind = str(g["pot_kind"])
    pk = {k[4:]: float(g[k]) for k in g if k.startswith("pot_") and k != "pot_kind" and k != "pot_energy"}
    if kind == "ideal":
        pot = MiePotential(ideal=True)
    else:
        pot = MiePotential(epsilon=pk.get("epsilon", 1.0), sigma=pk.get("sigma", 1.0),
                           m=pk.get("m", 12), n=pk.get("n", 6), a=pk.get("a", 0.0))
    steps = int(g["sm_steps"])
    timings = [float(t) for t in g["sm_quench_timings"]] if "sm_quench_timings" in g else []
    kw = dict(
        r=g["r0"], v=g["v0"], boxlen=float(g["boxlen"]), dt=float(g["dt"]), pot=pot,
        mass=float(g["sm_mass"]) if "sm_mass" in g else 1.0,
        pbc=bool(g["sm_pbc"]) if "sm_pbc" in g else True,
        radius=float(g["sm_radius_particle"]),
        temp=float(g["sm_temp"]) if "sm_temp" in g else 1.0,
        kb=1.0,
        gravity=bool(g["solver_gravity"]) if "solver_gravity" in g else False,
        isothermal=bool(g["solver_isothermal"]) if "solver_isothermal" in g else False,
        algorithm=str(g["solver_algorithm"]) if "solver_algorithm" in g else "verlet",
        quench_steps=[round(steps * t) for t in timings],  # mdsea/core.py:86
        quench_temps=[float(t) for t in g["sm_quench_temps"]] if "sm_quench_temps" in g else [],
    )
    kw.update(over)
    return OracleSolver(**kw)


def force_rel_err(a, b):
    """Per-particle force error metric: |dF_i| / max(|F_i|, median|F|); a,b are (ndim,N)."""
    d = np.sqrt(((a - b) ** 2).sum(axis=0))
    nb = np.sqrt((b ** 2).sum(axis=0))
    scale = np.maximum(nb, np.median(nb))
    scale = np.where(scale > 0, scale, 1.0)
    return float((d / scale).max())
