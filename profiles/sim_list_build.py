"""CPU model of the warp-cooperative list build's work (no GPU): candidate tests, staged candidates, chunks and append-loop
iterations per tile for different tile shapes, from a melted LJ liquid of the C oracle (TEST INFRASTRUCTURE).

The kernel (csrc/cells_nbr_build.cuh) gives every lane group of a warp its own candidate stream per stencil row: the cells
[first cell of the group - 1, last cell + 1] of that row, walked in chunks of 32 candidates; every lane of the group tests
every candidate of the chunk (8.75 instructions per test in r01), then appends its hits (one loop iteration per hit, the
warp waits for the lane with the most hits of the chunk).  Per tile the model reports

    tests      candidate tests per PARTICLE                      (x 8.75 instructions / 32 lanes per warp instruction)
    staged     candidates staged into shared memory per particle (x ~16 instructions, one lane each)
    chunks     chunk iterations per tile                         (x ~35 instructions of per-chunk overhead)
    append     append-loop iterations per tile = sum over chunks of max over lanes of the lane's hits (x ~11 instructions)
    /row       the same if the appends were done once per stencil ROW (max over lanes of the lane's hits in the row)
    ideal      hits of the tile / lanes: what the append loop would cost without divergence

    python profiles/sim_list_build.py [per_row=32] [melt_steps=300]
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle_c  # noqa: E402
from oracle.oracle_np import box_length, default_dt, lattice_positions, mb_velocities  # noqa: E402

per_row = int(sys.argv[1]) if len(sys.argv) > 1 else 32
melt = int(sys.argv[2]) if len(sys.argv) > 2 else 300
RC, SKIN = 2.5, 0.3
n = per_row ** 3
L = box_length(3, n, 0.5, 0.3)
np.random.seed(0)
r0 = lattice_positions(3, n, L)
v0 = mb_velocities(3, n, 1.0, 1.0, 1.0) + np.random.default_rng(0).normal(scale=0.7, size=(3, n))
s = oracle_c.CutoffSystem(r0, v0, L, RC, SKIN, default_dt(0.5, v0))
s.run(melt)
r, _, _ = s.state()
s.close()
cell_of, order, nc = oracle_c.cells(r, L, RC + SKIN)
off, idx = oracle_c.neighbors(r, L, RC + SKIN)
slot_of = np.empty(n, dtype=np.int64)
slot_of[order] = np.arange(n)
cell_start = np.searchsorted(cell_of[order], np.arange(nc ** 3 + 1))
cx_of, cy_of, cz_of = cell_of % nc, (cell_of // nc) % nc, cell_of // (nc * nc)
nbr_slots = [np.sort(slot_of[idx[off[i]:off[i + 1]]]) for i in range(n)]


# The append loop runs per WARP: the chunks of the groups of a warp are walked in lock step, so the iterations of a chunk
# step are the max over all lanes of the warp.  Model that directly (groups of one tile advance together).
def model_warp(tile, group, chunk=32):
    lanes = 32
    per_lane = tile // lanes
    tests = staged = chunks = append = append_row = hits_tot = parts = 0
    for t0 in range(0, n - tile + 1, tile * 5):
        ids = order[t0:t0 + tile]
        parts += tile
        streams = []   # per group: list of (candidates of one row, members)
        for g0 in range(0, tile, group):
            gi = ids[g0:g0 + group]
            cy, cz = cy_of[gi[0]], cz_of[gi[0]]
            members = [i for i in gi if cy_of[i] == cy and cz_of[i] == cz]   # the first pass of the group (the rest: another pass)
            rest = [i for i in gi if not (cy_of[i] == cy and cz_of[i] == cz)]
            for mem in (members, rest):
                if not mem:
                    continue
                cy, cz = cy_of[mem[0]], cz_of[mem[0]]
                cxs = [cx_of[i] for i in mem]
                lo, hi = min(cxs) - 1, max(cxs) + 1
                rows = []
                for dz in (-1, 0, 1):
                    for dy in (-1, 0, 1):
                        row = (((cz + dz) % nc) * nc + (cy + dy) % nc) * nc
                        rows.append(np.concatenate([np.arange(cell_start[row + x % nc], cell_start[row + x % nc + 1]) for x in range(lo, hi + 1)]))
                streams.append((rows, mem))
        # lock step over the 9 rows; per row the warp iterates max-over-groups chunks
        for k in range(9):
            longest = max(st[0][k].shape[0] for st in streams)
            # append once per ROW instead of once per chunk: max over lanes of the lane's hits in the whole row
            worst_row = 0
            for rows, mem in streams:
                for l0 in range(0, len(mem), per_lane):
                    worst_row = max(worst_row, sum(int(np.isin(rows[k], nbr_slots[i], assume_unique=True).sum()) for i in mem[l0:l0 + per_lane]))
            append_row += worst_row
            for c0 in range(0, longest, chunk):
                chunks += 1
                worst = 0
                for rows, mem in streams:
                    ch = rows[k][c0:c0 + chunk]
                    if ch.shape[0] == 0:
                        continue
                    staged += ch.shape[0]
                    tests += ch.shape[0] * len(mem)
                    # hits per LANE: a lane owns per_lane consecutive particles of the tile
                    for l0 in range(0, len(mem), per_lane):
                        h = sum(int(np.isin(ch, nbr_slots[i], assume_unique=True).sum()) for i in mem[l0:l0 + per_lane])
                        hits_tot += h
                        worst = max(worst, h)
                append += worst
    tiles = parts / tile
    return dict(tests=tests / parts, staged=staged / parts, chunks=chunks / tiles, append=append / tiles, append_row=append_row / tiles, ideal=hits_tot / lanes / tiles,
                tiles_per_1e6=1e6 / tile)


print(f"N = {n}, nc = {nc}, mean list length {np.diff(off).mean():.1f}")
print(f"{'shape':34s} {'tests':>7s} {'staged':>7s} {'chunks':>7s} {'append':>7s} {'/row':>6s} {'ideal':>6s} {'~Minstr per 1e6 particles':>26s}")
for label, tile, group in (("32 slots, 4 groups of 8 (r01 start)", 32, 8), ("32 slots, 2 groups of 16 (now)", 32, 16),
                           ("32 slots, 1 group", 32, 32), ("64 slots (2 per lane), 4 groups of 16", 64, 16),
                           ("64 slots (2 per lane), 2 groups of 32", 64, 32)):
    m = model_warp(tile, group)
    per_lane = tile // 32
    # warp instructions per tile: tests are shared by the lanes (x per_lane particles each), LDS + FMA part ~ 8.75 / test for
    # one particle per lane, ~ 1 + 7.75 x per_lane for several (one broadcast load serves all of a lane's particles)
    instr = (m["tests"] * tile / 32 / per_lane) * (1.0 + 7.75 * per_lane) + m["staged"] * tile / 32 * 16 + m["chunks"] * 35 + m["append"] * 11
    print(f"{label:34s} {m['tests']:7.0f} {m['staged']:7.1f} {m['chunks']:7.1f} {m['append']:7.1f} {m['append_row']:6.1f} {m['ideal']:6.1f} {instr * m['tiles_per_1e6'] / 1e6:26.0f}")
