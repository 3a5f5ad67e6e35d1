"""Pair potentials: the reference's bounded-Mie family behind the same ``Potential`` interface
(mdsea/potentials.py:15-147).

The NumPy callables below are the user-facing definition (plotting, ``potminimum``); the solver does
NOT call them in the step loop -- it recognises ``pot.ff`` / ``pot.kwargs`` and maps them to device
constants (``device_params``).  A potential built from unknown callables is rejected by the solver
with NotImplementedError: there is no CPU fallback by design.
"""
import logging
from typing import Callable, Optional

import numpy as np
from scipy import optimize

log = logging.getLogger(__name__)


def _mie_prefactor(m, n):
    return (m / (m - n)) * (m / n) ** (n / (m - n))


def pf_boundedmie(r, a, epsilon, sigma, m, n):
    """V(r) = lambda eps [(sigma/sqrt(a^2+r^2))^m - (...)^n]  (potentials.py:15-19)."""
    t = sigma / np.sqrt(a * a + r * r)
    return _mie_prefactor(m, n) * epsilon * (t ** m - t ** n)


def pf_mie(r, epsilon, sigma, m, n):
    return pf_boundedmie(r, a=0, epsilon=epsilon, sigma=sigma, m=m, n=n)


def pf_lennardjones(r, epsilon, sigma):
    return pf_mie(r, epsilon=epsilon, sigma=sigma, m=12, n=6)


def pf_ideal(r, **kwargs):
    return 0


def ff_boundedmie(r, a, epsilon, sigma, m, n):
    """dV/dr = lambda eps r (n t^n - m t^m) / (a^2 + r^2)  (potentials.py:43-48)."""
    s2 = a * a + r * r
    t = sigma / np.sqrt(s2)
    return _mie_prefactor(m, n) * epsilon * r * ((n * t ** n) - (m * t ** m)) / s2


def ff_mie(r, epsilon, sigma, m, n):
    return ff_boundedmie(r, a=0, epsilon=epsilon, sigma=sigma, m=m, n=n)


def ff_lennardjones(r, epsilon, sigma):
    return ff_mie(r, epsilon=epsilon, sigma=sigma, m=12, n=6)


def ff_ideal(r, **kwargs):
    return 0


class Potential:
    """Pair of callables ``ff(r, **kwargs)`` / ``pf(r, **kwargs)`` plus their kwargs (potentials.py:70-99)."""

    def __init__(self, force: Optional[Callable], potential: Optional[Callable], step: bool = False,
                 kwargs: Optional[dict] = None) -> None:
        self.ff = force
        self.pf = potential
        self.is_step = step
        self.kwargs: dict = kwargs if kwargs is not None else {}

    @classmethod
    def ideal(cls) -> "Potential":
        return cls(force=ff_ideal, potential=pf_ideal)

    @classmethod
    def boundedmie(cls, a, epsilon, sigma, m, n) -> "Potential":
        return cls(force=ff_boundedmie, potential=pf_boundedmie, kwargs=dict(a=a, epsilon=epsilon, sigma=sigma, m=m, n=n))

    @classmethod
    def mie(cls, epsilon, sigma, m, n) -> "Potential":
        return cls(force=ff_mie, potential=pf_mie, kwargs=dict(epsilon=epsilon, sigma=sigma, m=m, n=n))

    @classmethod
    def lennardjones(cls, epsilon, sigma) -> "Potential":
        return cls(force=ff_lennardjones, potential=pf_lennardjones, kwargs=dict(epsilon=epsilon, sigma=sigma))

    def potential(self, r):
        return self.pf(r, **self.kwargs)

    def force(self, r):
        return self.ff(r, **self.kwargs)

    def potminimum(self, xtol: float = 1e-8):
        """Equilibrium distance via scipy.optimize.fmin, host side (potentials.py:136-147)."""
        return optimize.fmin(self.pf, 0.5, tuple(self.kwargs.values()), xtol=xtol, disp=False)[0]

    # -- mapping onto the device kernels -------------------------------------------------------
    def device_params(self) -> dict:
        """{kind, epsilon, sigma, m, n, a} for the C ABI, or NotImplementedError for foreign callables."""
        kw = self.kwargs
        if self.ff is ff_ideal:
            return dict(kind=0, epsilon=0.0, sigma=1.0, m=12.0, n=6.0, a=0.0)
        if self.ff is ff_lennardjones:
            return dict(kind=1, epsilon=float(kw["epsilon"]), sigma=float(kw["sigma"]), m=12.0, n=6.0, a=0.0)
        if self.ff is ff_mie:
            return dict(kind=1, epsilon=float(kw["epsilon"]), sigma=float(kw["sigma"]), m=float(kw["m"]), n=float(kw["n"]), a=0.0)
        if self.ff is ff_boundedmie:
            return dict(kind=1, epsilon=float(kw["epsilon"]), sigma=float(kw["sigma"]), m=float(kw["m"]), n=float(kw["n"]),
                        a=float(kw["a"]))
        raise NotImplementedError(
            "mdsea_b200 runs the step loop on the GPU and only knows the bounded-Mie family "
            "(Potential.ideal/lennardjones/mie/boundedmie); arbitrary Python force callables have no device "
            "implementation and there is no CPU fallback.")
