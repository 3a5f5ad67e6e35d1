"""Small driver for ncu captures of the force kernels (run under gpurun; see profiles/README.md).
usage: python profiles/prof_force.py {c2|c3} {fp64|fp32} [n_evals]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import WORKLOADS, initial_state, make_ctx  # noqa: E402

wl = WORKLOADS[sys.argv[1]]
precision = sys.argv[2]
evals = int(sys.argv[3]) if len(sys.argv) > 3 else 5
n, L, r0, v0, dt = initial_state(wl["ndim"], wl["per_row"], wl["vf"])
ctx = make_ctx(wl, n, L, dt, 0, precision, 8)
ctx.set_temperature(1.0)
ctx.set_state(r0, v0)
rows = ctx.run(8, kb=1.0, record=False)   # move off the perfect lattice
for _ in range(evals):
    acc, pe = ctx.update_acc()
print("pe", pe, "launches", ctx.stats()["kernel_launches"])
ctx.close()
