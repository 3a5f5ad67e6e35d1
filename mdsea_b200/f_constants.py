"""Physical constants consumed by Config (same keys as mdsea/f_constants.py:4-7)."""
from scipy import constants as _sc

physical_constants = {
    "k_boltzmann": _sc.value("Boltzmann constant"),
    "gravity_acceleration": _sc.value("standard acceleration of gravity"),
}
