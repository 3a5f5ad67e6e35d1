"""One GPU at the per-GPU size of BASELINE configs[4] (64 M particles / 8 GPUs = 8 M each): 200^3 particles generated on
the device (csrc/gen.cu), fp32 pair math, cell + Verlet list.  Checks that the sizes fit (HBM, 32-bit indices) and reports
the per-kernel-family times; the lattice energy per particle must equal the 1e6-particle lattice's (same density).
usage: python profiles/probe_8m_per_gpu.py [per_row=200] [md_steps=40]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from bench import WORKLOADS, make_ctx  # noqa: E402
from mdsea_b200.gen import mb_cdf  # noqa: E402
from mdsea_b200.helpers import get_dt, nsphere_volume  # noqa: E402

per_row = int(sys.argv[1]) if len(sys.argv) > 1 else 200
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 40
wl = WORKLOADS["c3"]
out = {}
for pr in (100, per_row):
    n = pr ** 3
    L = (n * nsphere_volume(3, 0.5) / wl["vf"]) ** (1 / 3)
    ctx = make_ctx(wl, n, L, 0.004, 0, "fp32", 20)
    ctx.set_temperature(1.0)
    t0 = time.perf_counter()
    ms = ctx.generate_state(mb_cdf(1.0, 1.0, 1.0), 0.001, seed=0, isotropic=False)   # reference semantics: x-only velocities
    t_gen = time.perf_counter() - t0
    ctx.set_dt(get_dt(0.5, ms))
    _, pe0 = ctx.update_acc()
    ctx.run(20, kb=1.0, record=False)
    ctx.set_profiling(True)
    s0 = ctx.stats()
    done = 0
    while done < steps:
        rows = ctx.run(20, kb=1.0, record=False)
        done += 20
    s1 = ctx.stats()
    d = {k: s1[k] - s0[k] for k in ("ms_force", "ms_rebuild", "ms_integrate", "force_evals", "list_rebuilds", "pair_evals_unique")}
    out[n] = dict(generate_s=t_gen, mean_speed=ms, lattice_pe=pe0, force_us=1e3 * d["ms_force"] / max(d["force_evals"], 1),
                  rebuild_us=1e3 * d["ms_rebuild"] / max(d["list_rebuilds"], 1), integrate_us_per_step=1e3 * d["ms_integrate"] / steps,
                  rebuilds=d["list_rebuilds"], pairs_per_eval=d["pair_evals_unique"] / max(d["force_evals"], 1),
                  pe_last=float(rows["pe"][-1]), ke_last=float(rows["ke"][-1]), nbr_entries=s1["nbr_entries"])
    ctx.close()
a, b = out[100 ** 3], out[per_row ** 3]
assert np.isfinite(b["pe_last"]) and np.isfinite(b["ke_last"])
assert abs(a["lattice_pe"] - b["lattice_pe"]) <= 1e-5 * abs(a["lattice_pe"]), (a["lattice_pe"], b["lattice_pe"])
print(json.dumps(out))
