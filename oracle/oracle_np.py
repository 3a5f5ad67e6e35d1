"""TEST INFRASTRUCTURE ONLY -- NumPy float64 restatement of mdsea's
ContinuousPotentialSolver step loop (the hot path this repo accelerates).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module; the product (mdsea_b200/) never does and fails loudly
when its CUDA library is missing.

Parity pinning: the upstream reference ships no golden vectors for this path
(its only test checks the version string, tests/test_mdsea.py:1-5), so this oracle
is pinned against OUTPUTS OF THE UNMODIFIED REFERENCE executed in the build
container (oracle/gen_golden.py -> tests/golden/*.npz, via oracle/refshim.py) and,
when /root/reference is present, against the live reference in
tests/test_oracle_golden.py.

Every function cites the reference lines (relative to /root/reference/) it follows.
The restatement keeps the reference's evaluation structure -- explicit i<j pair
table, minimum image through rint of a TRUE division, dV/dr through two integer
powers, bincount scatter -- because that structure is what the CPU baseline timing
is meant to represent.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np
from scipy import special
from scipy.interpolate import interp1d

F64 = np.float64


# ----------------------------------------------------------------------------
# system geometry / initial conditions
# ----------------------------------------------------------------------------

def box_length(ndim: int, num_particles: int, radius: float, vol_fraction: float) -> float:
    """mdsea/core.py:113-115 with mdsea/helpers.py:54-56 (n-sphere volume)."""
    vol_particle = math.pi ** (ndim / 2) * (radius ** ndim) / special.gamma(ndim / 2 + 1)
    return (num_particles * vol_particle / vol_fraction) ** (1 / ndim)


def lattice_positions(ndim: int, num_particles: int, boxlen: float) -> np.ndarray:
    """mdsea/gen.py:144-156 -- simple cubic lattice, x fastest; returns (ndim, N) f64."""
    per_row = round(num_particles ** (1 / ndim))
    if per_row ** ndim != num_particles:
        raise AssertionError("num_particles must be a perfect ndim-th power")
    ticks = 0.5 + np.arange(per_row, dtype=int)
    span = boxlen / per_row
    out = np.zeros((ndim, num_particles), dtype=F64)
    for d in range(ndim):
        out[d] = span * np.tile(np.repeat(ticks, per_row ** d), per_row ** (ndim - d - 1))
    return out


_SPEED_GRID = np.arange(0, 50, 0.001)


def _mb_cdf(mass: float, temp: float, kb: float) -> np.ndarray:
    """mdsea/gen.py:27-37 -- Maxwell-Boltzmann speed CDF on the 0.001 grid."""
    a = math.sqrt(kb * temp / mass)
    t1 = special.erf(_SPEED_GRID / (math.sqrt(2) * a))
    t2 = np.sqrt(2 / math.pi) * _SPEED_GRID
    t3 = np.exp(-(_SPEED_GRID ** 2) / (2 * a ** 2)) / a
    return t1 - t2 * t3


def mb_velocities(ndim: int, num_particles: int, mass: float, temp: float, kb: float) -> np.ndarray:
    """mdsea/gen.py:98-123 and :83-92 -- inverse-CDF speeds times hyperspherical
    direction cosines.  Uses the GLOBAL numpy RNG in the reference's call order
    (random(N); then ndim-2 draws of uniform(0,pi,N); then uniform(0,2pi,N)) and keeps
    the whole-array np.prod of gen.py:85,87,91 (a scalar product over ALL particles'
    sines, which collapses vy,vz towards 0 in 3-D)."""
    if temp == 0:
        return np.zeros((ndim, num_particles), dtype=F64)
    inv_cdf = interp1d(_mb_cdf(mass, temp, kb), _SPEED_GRID)
    speeds = inv_cdf(np.random.random(num_particles))
    thetas = [np.random.uniform(0, math.pi, num_particles) for _ in range(ndim - 2)]
    thetas.append(np.random.uniform(0, 2 * math.pi, num_particles))
    comps = []
    for i in range(ndim):
        if i == ndim - 1:
            c = speeds * np.cos(thetas[ndim - 2]) * np.prod([np.sin(thetas[j]) for j in range(ndim - 2)])
        elif i == ndim - 2:
            c = speeds * np.sin(thetas[ndim - 2]) * np.prod([np.sin(thetas[j]) for j in range(ndim - 2)])
        elif i == 0:
            c = speeds * np.cos(thetas[0])
        else:
            c = speeds * np.cos(thetas[i]) * np.prod([np.sin(thetas[j]) for j in range(i)])
        comps.append(c)
    return np.array(comps).astype(F64)


def default_dt(radius: float, v: np.ndarray) -> float:
    """mdsea/simulator.py:86-87 with mdsea/helpers.py:89-100 and mdsea/quicker.py:44-46."""
    mean_speed = float(np.sqrt(np.add.reduce(v ** 2, axis=0, dtype=F64), dtype=F64).mean())
    if mean_speed == 0:
        mean_speed = 1
    return 0.01 * radius / mean_speed


# ----------------------------------------------------------------------------
# pair potential (bounded Mie family; LJ = m 12, n 6, a 0)
# ----------------------------------------------------------------------------

@dataclass
class MiePotential:
    """mdsea/potentials.py:15-19 (pf_boundedmie) and :43-48 (ff_boundedmie)."""

    epsilon: float = 1.0
    sigma: float = 1.0
    m: int = 12
    n: int = 6
    a: float = 0.0
    ideal: bool = False

    def _lambda(self) -> float:
        m, n = self.m, self.n
        return (m / (m - n)) * (m / n) ** (n / (m - n))

    def energy(self, r):
        if self.ideal:
            return 0
        term = self.sigma / np.sqrt(self.a * self.a + r * r)
        return self._lambda() * self.epsilon * (term ** self.m - term ** self.n)

    def dvdr(self, r):
        if self.ideal:
            return 0
        aprs = self.a * self.a + r * r
        term = self.sigma / np.sqrt(aprs)
        return self._lambda() * self.epsilon * r * ((self.n * term ** self.n) - (self.m * term ** self.m)) / aprs


# ----------------------------------------------------------------------------
# the step loop
# ----------------------------------------------------------------------------

@dataclass
class OracleSolver:
    """State + step loop of mdsea/simulator.py (_BaseSimulator :51-137,
    ContinuousPotentialSolver :511-589), all-pairs, float64.

    r, v are (ndim, N) SoA exactly like the reference's r_vec / v_vec.
    `r_cutoff` is the documented EXTENSION used by the large-N path: plain
    truncation, a pair contributes to force and energy iff dist < r_cutoff
    (strict, as mdsea/simulator.py:283), no shift, no tail correction.  The
    reference itself cannot run with a cutoff (simulator.py:48-49, :289-309).
    """

    r: np.ndarray
    v: np.ndarray
    boxlen: float
    dt: float
    pot: MiePotential = field(default_factory=MiePotential)
    mass: float = 1.0
    pbc: bool = True
    radius: float = 0.5
    temp: float = 1.0
    kb: float = 1.0
    gravity: bool = False
    g_acc: float = 9.80665
    isothermal: bool = False
    restitution: float = 1.0
    quench_steps: List[int] = field(default_factory=list)
    quench_temps: List[float] = field(default_factory=list)
    algorithm: str = "verlet"
    r_cutoff: Optional[float] = None

    def __post_init__(self):
        self.r = np.array(self.r, dtype=F64)
        self.v = np.array(self.v, dtype=F64)
        self.ndim, self.n = self.r.shape
        self.acc = np.zeros_like(self.r)
        self.temp_init = self.temp
        self.mean_ke = None
        self.mean_pe = None
        self.step = 0
        self.dists = None
        self.quench_temps = list(self.quench_temps)
        # i<j table in itertools.combinations order (simulator.py:122-137)
        iu = np.triu_indices(self.n, k=1)
        self._i = iu[0].astype(np.int32)
        self._j = iu[1].astype(np.int32)
        self.pairs_evaluated = 0
        if self.algorithm not in ("verlet", "simple"):
            raise KeyError(f"Algorithm '{self.algorithm}' not found in ('verlet', 'simple')")

    # -- boundary conditions ------------------------------------------------
    def apply_bc(self):
        if self.pbc:
            # simulator.py:156-166 -- one-shot strict wrap
            self.r[np.where(self.r < 0)] += self.boxlen
            self.r[np.where(self.r > self.boxlen)] -= self.boxlen
        else:
            # simulator.py:168-179 -- clamp + reflect
            w = np.where(self.r - self.radius < 0)
            self.r[w] = self.radius
            self.v[w] *= -self.restitution
            w = np.where(self.r + self.radius > self.boxlen)
            self.r[w] = self.boxlen - self.radius
            self.v[w] *= -self.restitution

    # -- pair geometry + forces --------------------------------------------
    def pair_geometry(self):
        """simulator.py:257-280: dr = r_j - r_i for i<j; dr -= rint(dr/L)*L; dist."""
        rt = np.stack(self.r, axis=-1)
        dr = rt[self._j] - rt[self._i]
        if self.pbc:
            dr -= np.rint(dr / self.boxlen) * self.boxlen
        dists = np.sqrt(np.add.reduce(dr ** 2, axis=1, dtype=F64), dtype=F64)
        return dr, dists

    def update_acc(self):
        """simulator.py:295-311: acc_p = unit * dV/dr / m; +on i, -on j (bincount)."""
        dr, dists = self.pair_geometry()
        ii, jj = self._i, self._j
        if self.r_cutoff is not None:
            keep = dists < self.r_cutoff
            dr, dists, ii, jj = dr[keep], dists[keep], ii[keep], jj[keep]
        self.dists = dists
        self.pairs_evaluated += dists.shape[0]
        acc = np.zeros_like(self.r)
        if dists.shape[0] and not self.pot.ideal:
            unit = dr / dists[:, None]
            acc_p = unit * self.pot.dvdr(dists[:, None]) / self.mass
            for d in range(self.ndim):
                acc[d] += np.bincount(ii, acc_p[:, d], minlength=self.n)
                acc[d] -= np.bincount(jj, acc_p[:, d], minlength=self.n)
        self.acc = acc
        return acc

    def pair_force_scale(self):
        """S_i = sum_j |f_ij| / m: magnitude of the terms update_acc adds up for particle i.  Used as the
        denominator of the fp32 force tolerance (backward-error sense): near a lattice |F_i| << S_i, and
        no fp32 pair arithmetic can resolve F_i to 1e-5 of itself."""
        dr, dists = self.pair_geometry()
        ii, jj = self._i, self._j
        if self.r_cutoff is not None:
            keep = dists < self.r_cutoff
            dists, ii, jj = dists[keep], ii[keep], jj[keep]
        mag = np.abs(self.pot.dvdr(dists)) / self.mass
        return np.bincount(ii, mag, minlength=self.n) + np.bincount(jj, mag, minlength=self.n)

    # -- energies -------------------------------------------------------------
    def update_energies(self):
        """simulator.py:241-255: PE = sum_pairs V / N (dists of the LAST force eval);
        KE = 0.5 m mean_i |v_i|^2."""
        if self.pot.ideal:
            self.mean_pe = 0 / self.n
        else:
            self.mean_pe = np.add.reduce(self.pot.energy(self.dists)) / self.n
        vt = np.stack(self.v, axis=-1)
        self.mean_ke = 0.5 * self.mass * np.einsum("ij,ij->i", vt, vt).mean()

    def _quench(self, target):
        """simulator.py:192-195 with :233-239 (T from the PREVIOUS step's KE, 2/3 factor)."""
        if self.mean_ke is not None:
            self.temp = (2 / 3) * self.mean_ke / self.kb
        if self.temp != 0.0:
            self.v *= (target / self.temp) ** 0.5

    # -- one step -------------------------------------------------------------
    def advance(self):
        """simulator.py:560-568 and :530-544."""
        self.apply_bc()
        if self.algorithm == "verlet":
            self.v += 0.5 * self.update_acc() * self.dt
            self.r += self.v * self.dt
            self.v += 0.5 * self.update_acc() * self.dt
        else:
            self.r += self.v * self.dt
            self.v += 0.5 * self.update_acc() * self.dt
        if self.gravity:
            self.v[-1] -= self.g_acc * self.dt  # simulator.py:201-205
        if self.restitution < 1:
            raise NotImplementedError  # simulator.py:546-548,565-566
        if self.isothermal:
            self._quench(self.temp_init)  # simulator.py:185-190
        if self.step in self.quench_steps:
            self._quench(self.quench_temps.pop(0))
        self.update_energies()

    def run(self, steps: int, record: bool = True, record_acc: bool = False):
        """simulator.py:570-589 -- row i is the state AFTER step i.  Returns dict of rows."""
        rows = {"r": [], "v": [], "pe": [], "ke": [], "temp": [], "acc": []}
        for self.step in range(self.step, self.step + steps):
            self.advance()
            if record:
                rows["r"].append(self.r.copy())
                rows["v"].append(self.v.copy())
                rows["pe"].append(float(self.mean_pe))
                rows["ke"].append(float(self.mean_ke))
                rows["temp"].append(float(self.temp))
                if record_acc:
                    rows["acc"].append(self.acc.copy())
        self.step += 1 if steps > 0 else 0
        return {k: np.array(val) for k, val in rows.items()}
