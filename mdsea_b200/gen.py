"""Initial-condition generators, bit-compatible with mdsea/gen.py (runs once, on the host).

Parity "on the same initial conditions" means reproducing gen.py's output bit for bit, including
its use of the GLOBAL numpy RNG in a fixed call order and the whole-array ``np.prod`` in the
direction cosines (gen.py:83-92), which in 3-D leaves vy and vz near zero.
"""
import logging
import math

import numpy as np
from scipy import special
from scipy.interpolate import interp1d

from mdsea_b200.constants import DTYPE

log = logging.getLogger(__name__)

MONTECARLO_SPEEDRANGE = np.arange(0, 50, 0.001)


def mb(mass, temp, k_boltzmann) -> np.ndarray:
    """Maxwell-Boltzmann speed density on the Monte-Carlo grid (gen.py:19-24)."""
    s2 = MONTECARLO_SPEEDRANGE ** 2
    pref = (mass / (2 * math.pi * k_boltzmann * temp)) ** 1.5
    return pref * (4 * math.pi * s2) * np.exp(-mass * s2 / (2 * k_boltzmann * temp))


def mb_cdf(mass, temp, k_boltzmann) -> np.ndarray:
    """Maxwell-Boltzmann speed CDF on the Monte-Carlo grid (gen.py:27-37)."""
    a = math.sqrt(k_boltzmann * temp / mass)
    erf_part = special.erf(MONTECARLO_SPEEDRANGE / (math.sqrt(2) * a))
    lin = np.sqrt(2 / math.pi) * MONTECARLO_SPEEDRANGE
    gauss = np.exp(-(MONTECARLO_SPEEDRANGE ** 2) / (2 * a ** 2)) / a
    return erf_part - lin * gauss


class _Gen:
    def __init__(self, nparticles: int, ndim: int) -> None:
        self.nparticles, self.ndim = nparticles, ndim
        self.generated = False
        self.coords = np.zeros((ndim, nparticles), dtype=DTYPE)

    def _get(self) -> np.ndarray:
        if not self.generated:
            log.warning("Coordinates haven't been generated yet!")
        return self.coords.astype(DTYPE)


class VelGen(_Gen):
    def __init__(self, nparticles: int, ndim: int) -> None:
        super().__init__(nparticles, ndim)
        self.speeds = None

    def _direction(self, i, thetas):
        """Hyperspherical direction factor of component i, with gen.py's np.prod over all particles."""
        nd = self.ndim
        lead_sines = lambda upto: np.prod([np.sin(thetas[j]) for j in range(upto)])
        if i == nd - 1:
            return np.cos(thetas[nd - 2]), lead_sines(nd - 2)
        if i == nd - 2:
            return np.sin(thetas[nd - 2]), lead_sines(nd - 2)
        if i == 0:
            return np.cos(thetas[0]), None
        return np.cos(thetas[i]), lead_sines(i)

    def mb(self, mass, temp, k_boltzmann) -> np.ndarray:
        """Inverse-CDF Maxwell-Boltzmann speeds times random directions (gen.py:98-123)."""
        if temp == 0:
            self.coords = np.zeros((self.ndim, self.nparticles), dtype=DTYPE)
            self.generated = True
            return self._get()
        inv_cdf = interp1d(mb_cdf(mass, temp, k_boltzmann), MONTECARLO_SPEEDRANGE)
        self.speeds = inv_cdf(np.random.random(self.nparticles))
        thetas = [np.random.uniform(0, math.pi, self.nparticles) for _ in range(self.ndim - 2)]
        thetas.append(np.random.uniform(0, 2 * math.pi, self.nparticles))
        rows = []
        for i in range(self.ndim):
            trig, sines = self._direction(i, thetas)
            comp = self.speeds * trig
            rows.append(comp if sines is None else comp * sines)
        self.coords = np.array(rows)
        self.generated = True
        return self._get()


class PosGen(_Gen):
    def __init__(self, nparticles: int, ndim: int, boxlen: float) -> None:
        super().__init__(nparticles, ndim)
        self.boxlen = boxlen

    def simplecubic(self) -> np.ndarray:
        """Simple cubic lattice, x fastest (gen.py:144-156)."""
        per_row = round(self.nparticles ** (1 / self.ndim))
        assert self.nparticles == per_row ** self.ndim
        ticks = 0.5 + np.arange(per_row, dtype=int)
        span = self.boxlen / per_row
        for d in range(self.ndim):
            self.coords[d] = span * np.tile(np.repeat(ticks, per_row ** d), per_row ** (self.ndim - d - 1))
        self.generated = True
        return self._get()

    def random(self, pradius) -> np.ndarray:
        """Uniform random positions at least one radius from the walls (gen.py:158-168)."""
        usable = self.boxlen - 2 * pradius
        self.coords = usable * np.random.random((self.ndim, self.nparticles))
        self.coords += pradius
        self.generated = True
        return self._get()
