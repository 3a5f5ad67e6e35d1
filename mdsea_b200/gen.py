"""Initial-condition generators, bit-compatible with mdsea/gen.py (runs once, on the host).

Parity "on the same initial conditions" means reproducing gen.py's output bit for bit, including
its use of the GLOBAL numpy RNG in a fixed call order and the whole-array ``np.prod`` in the
direction cosines (gen.py:83-92), which in 3-D leaves vy and vz near zero.  The arithmetic below
therefore keeps the reference's operation order wherever a rounding could differ; everything
else (structure, naming of internals) is this package's own.  `csrc/gen.cu` samples the same
table on the device (opt-in, different random stream).
"""
import logging
import math

import numpy as np
from scipy import special
from scipy.interpolate import interp1d

from mdsea_b200.constants import DTYPE

log = logging.getLogger(__name__)

MONTECARLO_SPEEDRANGE = np.arange(0, 50, 0.001)


class _SpeedTable:
    """Maxwell-Boltzmann speed law tabulated on the Monte-Carlo grid (gen.py:19-37)."""

    def __init__(self, mass, temp, k_boltzmann, grid=MONTECARLO_SPEEDRANGE):
        self.mass, self.temp, self.kb, self.grid = mass, temp, k_boltzmann, grid
        self.kt = k_boltzmann * temp

    def density(self) -> np.ndarray:
        s2 = self.grid ** 2
        norm = (self.mass / (2 * math.pi * self.kb * self.temp)) ** 1.5   # products left to right, like gen.py:21-23
        shell = 4 * math.pi * s2
        return norm * shell * np.exp(-self.mass * s2 / (2 * self.kb * self.temp))

    def cumulative(self) -> np.ndarray:
        a = math.sqrt(self.kt / self.mass)
        head = special.erf(self.grid / (math.sqrt(2) * a))
        ramp = np.sqrt(2 / math.pi) * self.grid
        bell = np.exp(-(self.grid ** 2) / (2 * a ** 2)) / a
        return head - ramp * bell

    def sample(self, uniforms: np.ndarray) -> np.ndarray:
        """Inverse-CDF transform of uniform deviates (gen.py:104-105; scipy's linear interp1d)."""
        return interp1d(self.cumulative(), self.grid)(uniforms)


def mb(mass, temp, k_boltzmann) -> np.ndarray:
    """Maxwell-Boltzmann speed density on the Monte-Carlo grid (gen.py:19-24)."""
    return _SpeedTable(mass, temp, k_boltzmann).density()


def mb_cdf(mass, temp, k_boltzmann) -> np.ndarray:
    """Maxwell-Boltzmann speed CDF on the Monte-Carlo grid (gen.py:27-37)."""
    return _SpeedTable(mass, temp, k_boltzmann).cumulative()


class _Gen:
    """Holds one (ndim, nparticles) coordinate array and whether a generator has filled it."""

    def __init__(self, nparticles: int, ndim: int) -> None:
        self.nparticles, self.ndim = nparticles, ndim
        self.generated = False
        self.coords = np.zeros((ndim, nparticles), dtype=DTYPE)

    def _done(self, coords) -> np.ndarray:
        self.coords, self.generated = coords, True
        return self._get()

    def _get(self) -> np.ndarray:
        if not self.generated:
            log.warning("Coordinates haven't been generated yet!")
        return self.coords.astype(DTYPE)


def _sine_product(thetas, upto):
    """gen.py:83-92 multiplies by np.prod over a LIST of per-particle arrays, i.e. over every particle."""
    return np.prod([np.sin(thetas[j]) for j in range(upto)])


class VelGen(_Gen):
    def __init__(self, nparticles: int, ndim: int) -> None:
        super().__init__(nparticles, ndim)
        self.speeds = None

    def _component(self, i, thetas):
        """Hyperspherical component i of the velocity (gen.py:83-92), operation order preserved."""
        nd, s = self.ndim, self.speeds
        last = nd - 2
        if i == nd - 1:
            return s * np.cos(thetas[last]) * _sine_product(thetas, last)
        if i == last:
            return s * np.sin(thetas[last]) * _sine_product(thetas, last)
        if i == 0:
            return s * np.cos(thetas[0])
        return s * np.cos(thetas[i]) * _sine_product(thetas, i)

    def mb(self, mass, temp, k_boltzmann) -> np.ndarray:
        """Inverse-CDF Maxwell-Boltzmann speeds times random directions (gen.py:98-123).

        RNG call order (global NumPy stream): random(N) for the speeds, then ndim - 2 times
        uniform(0, pi, N), then uniform(0, 2 pi, N)."""
        n, nd = self.nparticles, self.ndim
        if temp == 0:
            return self._done(np.zeros((nd, n), dtype=DTYPE))
        self.speeds = _SpeedTable(mass, temp, k_boltzmann).sample(np.random.random(n))
        polar = [np.random.uniform(0, math.pi, n) for _ in range(nd - 2)]
        thetas = polar + [np.random.uniform(0, 2 * math.pi, n)]
        return self._done(np.array([self._component(i, thetas) for i in range(nd)]))


class PosGen(_Gen):
    def __init__(self, nparticles: int, ndim: int, boxlen: float) -> None:
        super().__init__(nparticles, ndim)
        self.boxlen = boxlen

    def simplecubic(self) -> np.ndarray:
        """Simple cubic lattice, x fastest (gen.py:144-156)."""
        nd = self.ndim
        per_row = round(self.nparticles ** (1 / nd))
        assert self.nparticles == per_row ** nd
        ticks = 0.5 + np.arange(per_row, dtype=int)
        span = self.boxlen / per_row
        for d in range(nd):
            self.coords[d] = span * np.tile(np.repeat(ticks, per_row ** d), per_row ** (nd - d - 1))
        return self._done(self.coords)

    def random(self, pradius) -> np.ndarray:
        """Uniform random positions at least one radius from the walls (gen.py:158-168)."""
        usable = self.boxlen - 2 * pradius
        coords = usable * np.random.random((self.ndim, self.nparticles))
        coords += pradius
        return self._done(coords)
