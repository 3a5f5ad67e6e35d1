"""Run-time configuration singleton (interface of mdsea/config.py:7-35).

``Config.k_boltzmann`` is read when a solver is built (initial velocities) and at every run()
(quench temperature); ``Config.gravity_acceleration`` at every run() when gravity is on.
"""
from contextlib import contextmanager

from mdsea_b200.f_constants import physical_constants


class _Config:
    """Attribute bag restricted to the known physical constants; ``update`` rejects unknown keys
    with the reference's AttributeError (mdsea/config.py:13-17)."""

    __slots__ = tuple(physical_constants)

    def __init__(self) -> None:
        for name, default in physical_constants.items():
            object.__setattr__(self, name, default)

    def update(self, new: dict) -> None:
        for name in new:
            if name not in self.__slots__:
                raise AttributeError(f"Not a valid config field: {name}")
        for name, value in new.items():
            setattr(self, name, value)

    def snapshot(self) -> dict:
        return {name: getattr(self, name) for name in self.__slots__}

    def __repr__(self) -> str:
        return "_Config(" + ", ".join(f"{k}={v!r}" for k, v in self.snapshot().items()) + ")"


Config = _Config()


@contextmanager
def config_context(**overrides):
    """Temporarily override configuration fields (the reference's experimental helper, config.py:23-35)."""
    saved = Config.snapshot()
    Config.update(overrides)
    try:
        yield Config
    finally:
        Config.update(saved)
