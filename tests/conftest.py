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


@pytest.fixture(autouse=True)
def _no_sticky_cuda_error(request):
    """GPU tests: a test must not leave a CUDA error behind (cudaGetLastError is per thread and sticky: the next test's
    first launch check would report it as its own)."""
    yield
    if request.node.get_closest_marker("gpu") is None:
        return
    import ctypes

    try:
        rt = ctypes.CDLL("libcudart.so.12")
    except OSError:
        return
    err = rt.cudaGetLastError()
    if err in (35, 100):   # cudaErrorInsufficientDriver / cudaErrorNoDevice: a GPU test selected on a box without a GPU
        return
    assert err == 0, f"{request.node.name} left CUDA error {err} behind"


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


def force_rel_err_pairscale(a, b, pair_scale):
    """fp32 metric: |dF_i| / max(|F_i|, median|F|, S_i) with S_i = sum_j |f_ij| (oracle.pair_force_scale)."""
    d = np.sqrt(((a - b) ** 2).sum(axis=0))
    nb = np.sqrt((b ** 2).sum(axis=0))
    scale = np.maximum(np.maximum(nb, np.median(nb)), pair_scale)
    scale = np.where(scale > 0, scale, 1.0)
    return float((d / scale).max())


# ---------------------------------------------------------------------------------------------------
# GPU-side helpers (only used by -m gpu tests)
# ---------------------------------------------------------------------------------------------------
def has_cuda():
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


def golden_settings(g):
    """Settings of a golden fixture as plain Python values."""
    steps = int(g["sm_steps"])
    timings = [float(t) for t in g["sm_quench_timings"]] if "sm_quench_timings" in g else []
    kind = str(g["pot_kind"])
    pk = {k[4:]: float(g[k]) for k in g if k.startswith("pot_") and k not in ("pot_kind", "pot_energy")}
    return dict(
        ndim=int(g["sm_ndim"]), n=int(g["sm_num_particles"]), steps=steps, boxlen=float(g["boxlen"]), dt=float(g["dt"]),
        mass=float(g["sm_mass"]) if "sm_mass" in g else 1.0, pbc=bool(g["sm_pbc"]) if "sm_pbc" in g else True,
        radius=float(g["sm_radius_particle"]), temp=float(g["sm_temp"]) if "sm_temp" in g else 1.0,
        gravity=bool(g["solver_gravity"]) if "solver_gravity" in g else False,
        isothermal=bool(g["solver_isothermal"]) if "solver_isothermal" in g else False,
        algorithm=str(g["solver_algorithm"]) if "solver_algorithm" in g else "verlet",
        quench_steps=[round(steps * t) for t in timings],
        quench_temps=[float(t) for t in g["sm_quench_temps"]] if "sm_quench_temps" in g else [],
        pot_kind=0 if kind == "ideal" else 1, epsilon=pk.get("epsilon", 1.0), sigma=pk.get("sigma", 1.0),
        m=pk.get("m", 12.0), n_exp=pk.get("n", 6.0), a=pk.get("a", 0.0),
    )


def context_from_settings(s, precision=0, evals=2, ring=None, r_cutoff=0.0, skin=0.3, **over):
    from mdsea_b200 import _capi

    kw = dict(
        ndim=s["ndim"], num_particles=s["n"], boxlen=s["boxlen"], mass=s["mass"], radius=s["radius"], dt=s["dt"],
        pbc=1 if s["pbc"] else 0, algorithm=0 if s["algorithm"] == "verlet" else 1, restitution=1.0,
        pot_kind=s["pot_kind"], precision=precision, epsilon=s["epsilon"], sigma=s["sigma"], mie_m=s["m"],
        mie_n=s["n_exp"], mie_a=s["a"], r_cutoff=r_cutoff, skin=skin, force_evals_per_step=evals,
        ring_rows=ring or min(s["steps"], 128), device=0,
    )
    kw.update(over)
    return _capi.Context(**kw)


def run_context(ctx, s, steps, kb=1.0, g_acc=9.80665, record=True, start_step=0, qt=None):
    """Drive ctx for `steps` steps in ring-sized batches; returns stacked rows like OracleSolver.run."""
    ring = int(ctx.params.ring_rows)
    out = {k: [] for k in ("r", "v", "pe", "ke", "temp")}
    qt = list(s["quench_temps"]) if qt is None else qt
    done = 0
    while done < steps:
        n = min(ring, steps - done)
        q = None
        if s["isothermal"] or s["quench_steps"]:
            q = np.full((n, 2), np.nan)
            for k in range(n):
                st = start_step + done + k
                if s["isothermal"]:
                    q[k, 0] = s["temp"]
                if st in s["quench_steps"]:
                    q[k, 1] = qt.pop(0)
        rows = ctx.run(n, kb=kb, g=g_acc if s["gravity"] else 0.0, quench=q, record=record)
        for k in out:
            if rows[k] is not None:
                out[k].append(np.array(rows[k][:n]))
        done += n
    return {k: (np.concatenate(v) if v else None) for k, v in out.items()}
