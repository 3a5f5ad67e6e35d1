"""Run-time configuration singleton (interface of mdsea/config.py:7-35).

``Config.k_boltzmann`` is read when a solver is built (initial velocities) and at every run()
(quench temperature); ``Config.gravity_acceleration`` at every run() when gravity is on.
"""
from contextlib import contextmanager
from dataclasses import dataclass, fields

from mdsea_b200.f_constants import physical_constants


@dataclass
class _Config:
    k_boltzmann: float = physical_constants["k_boltzmann"]
    gravity_acceleration: float = physical_constants["gravity_acceleration"]

    def update(self, new: dict) -> None:
        known = {f.name for f in fields(self)}
        for key, value in new.items():
            if key not in known:
                raise AttributeError(f"Not a valid config field: {key}")
            setattr(self, key, value)

    def snapshot(self) -> dict:
        return {f.name: getattr(self, f.name) for f in fields(self)}


Config = _Config()


@contextmanager
def config_context(**overrides):
    saved = Config.snapshot()
    Config.update(overrides)
    try:
        yield Config
    finally:
        Config.update(saved)
