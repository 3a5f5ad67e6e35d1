"""Host-side numeric helpers kept for API compatibility (mdsea/quicker.py:37-46).

On the device the Euclidean norm is fused into the force kernels; these NumPy versions serve user
scripts and the analytics consumer.  ``cdist_`` of the reference is dead code and not provided.
"""
import numpy as np

from mdsea_b200.constants import DTYPE


def norm(x, axis=None):
    """sqrt(sum(x**2)) in float64 (mdsea/quicker.py:44-46)."""
    return np.sqrt(np.add.reduce(np.square(x), axis=axis, dtype=DTYPE), dtype=DTYPE)


def flipid(n):
    """Anti-diagonal identity (mdsea/quicker.py:37-41)."""
    return np.fliplr(np.eye(n, dtype=DTYPE))
