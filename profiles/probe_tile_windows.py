"""Window sizes of the tile lists (entries and slot ranges per block of 128 slots) at configs[2], on the lattice and after a warm-up.
usage: python profiles/probe_tile_windows.py [warmup_md_steps] [workload]"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import WORKLOADS, initial_state, make_ctx  # noqa: E402

warm = int(sys.argv[1]) if len(sys.argv) > 1 else 200
wl = WORKLOADS[sys.argv[2] if len(sys.argv) > 2 else "c3"]
n, L, r0, v0, dt = initial_state(wl["ndim"], wl["per_row"], wl["vf"])
ctx = make_ctx(wl, n, L, dt, 0, "fp32", 20)
ctx.set_temperature(1.0)
ctx.set_state(r0, v0)
for label, steps in (("lattice", 0), ("warm", warm)):
    done = 0
    while done < steps:
        ctx.run(min(20, steps - done), kb=1.0, record=False)
        done += 20
    nw, nr = ctx.get_tile_windows()
    st = ctx.stats()
    q = lambda a: [int(x) for x in np.percentile(a, [0, 50, 90, 99, 100])] if len(a) else []
    big = np.argsort(nw)[-5:] if len(nw) else []
    print(json.dumps(dict(state=label, blocks=len(nw), win_pct_0_50_90_99_100=q(nw), rng_pct=q(nr), biggest_blocks=[int(b) for b in big],
                          biggest_win=[int(nw[b]) for b in big], biggest_rng=[int(nr[b]) for b in big],
                          tile_lists=st["tile_lists"], tile_window_max=st["tile_window_max"], rebuilds=st["list_rebuilds"])))
ctx.close()
