import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np
from bench import WORKLOADS, initial_state, make_ctx
wl = WORKLOADS["c2"]
n, L, r0, v0, dt = initial_state(3, 16, 0.3)
for prec in ("fp64", "fp32"):
    ctx = make_ctx(wl, n, L, dt, 0, prec, 200)
    ctx.set_temperature(1.0); ctx.set_state(r0, v0)
    done = 0
    for target in (0, 200, 600, 2000, 6000):
        while done < target:
            ctx.run(200, kb=1.0, record=False); done += 200
        ctx.set_profiling(True); q0 = ctx.stats()
        for _ in range(5): ctx.update_acc()
        q1 = ctx.stats(); ctx.set_profiling(False)
        r, v, a, _ = ctx.get_state()
        print(prec, "steps", done, "force_us", 1e3*(q1["ms_force"]-q0["ms_force"])/5, "vy_rms", float(np.sqrt((v[1]**2).mean())), flush=True)
    ctx.close()
