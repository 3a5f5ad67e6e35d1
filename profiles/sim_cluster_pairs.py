"""CPU model (no GPU): how many pair evaluations would a CLUSTER-PAIR list (i-clusters x j-clusters selected by bounding
boxes, the GROMACS nbnxm / VERDICT r01 item 1 design) execute at the density and cutoff of BASELINE configs[2], against
the per-particle Verlet list the tile kernels walk?

The state is a melted LJ liquid from the C oracle (TEST INFRASTRUCTURE, like everything under oracle/).  Clusters are
built the way a cluster-pair code would: spatially sorted particles, consecutive groups of
`ci` (i side) / `cj` (j side) particles -- spatially compact (recursive median bisection), every cluster full.  A cluster pair is listed when the
distance between the two axis-aligned bounding boxes (minimum image) is below rc + skin; it is evaluated as ci x cj pair
interactions.  Reported per particle:

  listed     pair evaluations per i particle with a DIRECTED cluster list (no reaction forces, no atomics: what the
             deterministic kernels of this repository would need) and with a HALF list (Newton's third law, needs a
             reaction scatter)
  useful     the in-cutoff neighbours among them (the 32.5 the metric credits twice per unique pair)
  tiles8x4   8x4 sub-tiles (a warp's 32 lanes) of the listed cluster pairs that contain at least one in-range pair: what a
             kernel that skips empty sub-tiles with a warp vote still executes

    python profiles/sim_cluster_pairs.py [per_row=32] [melt_steps=300]
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
RL = RC + SKIN
n = per_row ** 3
L = box_length(3, n, 0.5, 0.3)
np.random.seed(0)
r0 = lattice_positions(3, n, L)
v0 = mb_velocities(3, n, 1.0, 1.0, 1.0) + np.random.default_rng(0).normal(scale=0.7, size=(3, n))
s = oracle_c.CutoffSystem(r0, v0, L, RC, SKIN, default_dt(0.5, v0))
s.run(melt)
r, _, _ = s.state()
s.close()
r = (r % L).T                                          # (n, 3) in [0, L)
off, idx = oracle_c.neighbors(r.T.copy(), L, RL)
listed_verlet = (off[-1] / n)
off_c, idx_c = oracle_c.neighbors(r.T.copy(), L, RC)
in_cut = off_c[-1] / n


def compact_order(pos, leaf):
    """Recursive median bisection along the longest extent down to groups of `leaf` particles: balanced, near-cubic
    clusters -- at least as compact as the grid-column + z-sort clusters of a production cluster-pair search."""
    out = []

    def rec(ids):
        if len(ids) <= leaf:
            out.append(ids)
            return
        p = pos[ids]
        d = int(np.argmax(p.max(axis=0) - p.min(axis=0)))
        ids = ids[np.argsort(p[:, d], kind="stable")]
        cut = max(leaf, (len(ids) // leaf // 2) * leaf)
        rec(ids[:cut])
        rec(ids[cut:])

    rec(np.arange(len(pos)))
    return np.concatenate(out)


def boxes(pos, size):
    m = len(pos) // size * size
    p = pos[:m].reshape(-1, size, 3)
    # periodic: unwrap every cluster around its first particle, so that a cluster across the box edge stays compact
    ref = p[:, :1, :]
    p = ref + ((p - ref + L / 2) % L - L / 2)
    return p.min(axis=1), p.max(axis=1), p


def model(ci, cj):
    order = compact_order(r, max(ci, cj))   # j clusters of cj < ci particles: halves of the i clusters
    pos = r[order]
    lo_i, hi_i, pi = boxes(pos, ci)
    lo_j, hi_j, pj = boxes(pos, cj)
    ni, nj = len(lo_i), len(lo_j)
    ctr_j = 0.5 * (lo_j + hi_j)
    half_j = 0.5 * (hi_j - lo_j)
    listed = useful = tiles = 0
    sample = np.random.default_rng(1).choice(ni, size=min(ni, 400), replace=False)
    for a in sample:
        ctr = 0.5 * (lo_i[a] + hi_i[a])
        half = 0.5 * (hi_i[a] - lo_i[a])
        d = ctr_j - ctr
        d -= np.rint(d / L) * L
        gap = np.maximum(np.abs(d) - half - half_j, 0.0)
        near = np.nonzero((gap ** 2).sum(axis=1) < RL * RL)[0]
        listed += len(near) * cj
        # exact pair distances of the listed cluster pairs
        dd = pj[near][None, :, :, :] - pi[a][:, None, None, :]          # (ci, pairs, cj, 3)
        dd -= np.rint(dd / L) * L
        d2 = (dd ** 2).sum(axis=-1)
        inr = d2 < RC * RC
        useful += inr.sum() / ci
        # 8x4 sub-tiles: groups of 8 i x 4 j (for ci < 8 the whole i cluster is one row group)
        gi = max(ci // 8, 1)
        inr_t = (d2 < RL * RL).reshape(gi, ci // gi, len(near), cj // 4, 4)
        tiles += inr_t.any(axis=(1, 4)).sum() * (ci // gi) * 4 / ci
    k = len(sample)
    ext = float((hi_i - lo_i).mean())
    return listed / k, useful / k, tiles / k, ext


print(f"N = {n} melted liquid, rc = {RC}, skin = {SKIN}: per-particle Verlet list {listed_verlet:.1f} entries, "
      f"{in_cut:.1f} neighbours inside rc")
print(f"{'clusters (i x j)':>18s} {'listed, directed':>17s} {'listed, half':>13s} {'8x4 tiles kept, directed':>25s} {'useful fraction':>16s}")
for ci, cj in ((8, 8), (8, 4), (4, 4)):
    listed, useful, tiles, ext = model(ci, cj)
    print(f"{ci:>8d} x {cj:<7d} {listed:17.1f} {listed / 2:13.1f} {tiles:25.1f} {in_cut / listed:16.3f}   (mean i-box edge {ext:.2f} sigma)")
print(f"{'per-particle list':>18s} {listed_verlet:17.1f} {listed_verlet / 2:13.1f} {'-':>25s} {in_cut / listed_verlet:16.3f}")
