"""Per-kernel-family timings of the cut-off path at configs[2] (1e6 particles), from the library's own CUDA-event
timers (profiling mode: one sync per family, so the sum is NOT the step time -- use bench.py for that).
usage: python profiles/ab_cutoff.py [fp32|fp64] [md_steps] [workload] [warmup_md_steps]
env: the A/B switches of INTEGRATION.md section 6 (MDSEA_NBR_BUILD, MDSEA_NBW_VARIANT, MDSEA_CUT_FUSE, ...)"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import WORKLOADS, initial_state, make_ctx  # noqa: E402

precision = sys.argv[1] if len(sys.argv) > 1 else "fp32"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 40
wl = WORKLOADS[sys.argv[3] if len(sys.argv) > 3 else "c3"]
n, L, r0, v0, dt = initial_state(wl["ndim"], wl["per_row"], wl["vf"])
ctx = make_ctx(wl, n, L, dt, 0, precision, 20)
ctx.set_temperature(1.0)
ctx.set_state(r0, v0)
warm = int(sys.argv[4]) if len(sys.argv) > 4 else 20   # move off the perfect lattice (the list build slows down as it melts)
while warm > 0:
    ctx.run(min(20, warm), kb=1.0, record=False)
    warm -= 20
ctx.set_profiling(True)
s0 = ctx.stats()
done = 0
while done < steps:
    k = min(20, steps - done)
    rows = ctx.run(k, kb=1.0, record=False)
    done += k
s1 = ctx.stats()
d = {k: s1[k] - s0[k] for k in s1 if k not in ("n_local", "n_ghost", "nbr_entries")}
out = dict(precision=precision, n=n, build=os.environ.get("MDSEA_NBR_BUILD", "warp"), md_steps=steps, nbr_entries=s1["nbr_entries"],
           force_us=1e3 * d["ms_force"] / max(d["force_evals"], 1), rebuild_us=1e3 * d["ms_rebuild"] / max(d["list_rebuilds"], 1),
           integrate_us_per_step=1e3 * d["ms_integrate"] / steps, rebuilds=d["list_rebuilds"], force_evals=d["force_evals"],
           unique_pairs_per_eval=d["pair_evals_unique"] / max(d["force_evals"], 1), pe_last=float(rows["pe"][-1]))
print(json.dumps(out))
ctx.close()
