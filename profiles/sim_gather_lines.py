"""CPU model of the list force kernel's gather pattern (no GPU): how many distinct 128-byte lines of the packed position
array (8 slots of 16 B per line) does one warp touch per list row, for different list-row layouts?  The kernel is bound
by L1 line wavefronts (profiles/README.md section 3), so lines x rows per tile is the quantity a layout has to reduce.

The state is a melted LJ liquid from the C oracle (TEST INFRASTRUCTURE, like everything under oracle/); slots are the
canonical (cell, id) order; a warp = 32 consecutive slots; entries of a particle are ascending in slot order.

    python profiles/sim_gather_lines.py [per_row=40] [melt_steps=300]

layouts:
  plain      row k = k-th neighbour of every lane (what k_nbr_force_* reads today)
  by_z       the list of a particle is cut into 3 segments by the neighbour's cell layer (dz = -1, 0, +1); inside a tile
             every segment is padded to the longest one, so that row k of all lanes points into the same layer
  by_yz      same with the 9 (dy, dz) cell rows
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle_c  # noqa: E402
from oracle.oracle_np import box_length, default_dt, lattice_positions, mb_velocities  # noqa: E402

per_row = int(sys.argv[1]) if len(sys.argv) > 1 else 40
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
cz = cell_of // (nc * nc)
cy = (cell_of // nc) % nc


def tile_rows(layout):
    """yields, per tile of 32 slots, a (rows, 32) array of neighbour slots (-1 = padding)"""
    for t0 in range(0, n - 31, 32):
        ids = order[t0:t0 + 32]
        lists = []
        for i in ids:
            nb = idx[off[i]:off[i + 1]]
            sl = slot_of[nb]
            o = np.argsort(sl)
            nb, sl = nb[o], sl[o]
            if layout == "plain":
                segs = [sl]
            else:
                dz = (cz[nb] - cz[i] + 1 + nc) % nc      # 0, 1, 2 for the layers below / same / above (periodic)
                if layout == "by_z":
                    key, nseg = dz, 3
                else:
                    dy = (cy[nb] - cy[i] + 1 + nc) % nc
                    key, nseg = dz * 3 + dy, 9
                segs = [sl[key == q] for q in range(nseg)]
            lists.append(segs)
        nseg = len(lists[0])
        width = [max(len(l[q]) for l in lists) for q in range(nseg)]
        rows = np.full((sum(width), 32), -1, dtype=np.int64)
        base = 0
        for q in range(nseg):
            for lane, l in enumerate(lists):
                rows[base:base + len(l[q]), lane] = l[q]
            base += width[q]
        yield rows


print(f"N = {n}, L = {L:.2f}, nc = {nc}, mean list length {np.diff(off).mean():.1f}")
print(f"{'layout':8s} {'rows/tile':>10s} {'lines/row':>10s} {'lines/tile':>11s} {'useful entries/row':>19s}")
for layout in ("plain", "by_z", "by_yz"):
    nrows = nlines = nuse = ntiles = 0
    for k, rows in enumerate(tile_rows(layout)):
        if k % 7:   # sample every 7th tile
            continue
        ntiles += 1
        nrows += rows.shape[0]
        for row in rows:
            live = row[row >= 0]
            nuse += live.shape[0]
            nlines += np.unique(live // 8).shape[0]
    print(f"{layout:8s} {nrows / ntiles:10.1f} {nlines / nrows:10.2f} {nlines / ntiles:11.1f} {nuse / nrows:19.1f}")

# ---- two particles per lane: one gather serves both if the lane walks the UNION of their lists (the r^2 < rc^2 test
# decides per particle, so no per-particle membership is needed).  How large is the union for slot neighbours?
def mean_union(pair_order):
    tot = cnt = 0
    for a, b in zip(pair_order[0::2], pair_order[1::2]):
        la, lb = idx[off[a]:off[a + 1]], idx[off[b]:off[b + 1]]
        tot += np.union1d(la, lb).shape[0] + 0   # (a is in b's list and vice versa: each still needs the other's position)
        cnt += 1
    return tot / cnt


sample = order[: (min(n, 20000) // 2) * 2]
print(f"\ntwo particles per lane, gathers per PAIR of particles (2 x {np.diff(off).mean():.1f} = {2 * np.diff(off).mean():.1f} separately):")
print(f"  canonical (cell, id) order, slots 2l and 2l+1 : {mean_union(sample):.1f}")
# in-cell order by x (an internal layout that keeps the cells but not the id order inside them)
key = cell_of.astype(np.float64) * 1e3 + (r[0] % L) / L
byx = np.argsort(key, kind="stable")
print(f"  cells kept, particles ordered by x inside a cell: {mean_union(byx[: sample.shape[0]]):.1f}")
# greedy nearest-neighbour pairing inside each cell (upper bound of what an in-cell spatial order could give)
pairs = []
for c in np.unique(cell_of[sample]):
    members = list(np.where(cell_of == c)[0])
    while len(members) > 1:
        a = members.pop(0)
        d = r[:, members] - r[:, [a]]
        d -= np.rint(d / L) * L
        j = int(np.argmin((d ** 2).sum(axis=0)))
        pairs += [a, members.pop(j)]
    if len(pairs) >= sample.shape[0]:
        break
print(f"  greedy nearest pairs inside a cell               : {mean_union(np.array(pairs)):.1f}")
