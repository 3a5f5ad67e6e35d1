"""Physical constants consumed by Config (same keys as mdsea/f_constants.py:4-7), from SciPy's CODATA table."""
from scipy import constants as _sc

_CODATA_NAMES = (
    ("k_boltzmann", "Boltzmann constant"),
    ("gravity_acceleration", "standard acceleration of gravity"),
)
physical_constants = {key: _sc.value(name) for key, name in _CODATA_NAMES}
