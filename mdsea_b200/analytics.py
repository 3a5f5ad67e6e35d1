"""Consumer of the simfile layout (mdsea/analytics.py:44-79): loads the five datasets and derives
per-step vectors, speeds and energies.  ``Analyser(sm, device=0)`` computes the speeds and their statistics -- the
only O(STEPS * N) arithmetic in here -- on the GPU (csrc/analytics.cu); the default
(device=None) is the reference's NumPy formulation, so a simfile can still be inspected on a machine without a
GPU (this class is a consumer of the solver's output, not part of the solver's path)."""
import numpy as np

from mdsea_b200.core import SysManager
from mdsea_b200.quicker import norm


class Analyser:
    def __init__(self, sm: SysManager, device=None) -> None:
        self.sm = sm
        self.r_coords = sm.get_ds(sm.rcoord_dsname)                     # (STEPS, NDIM, N)
        self.r_vecs = np.moveaxis(self.r_coords, 1, -1)                 # (STEPS, N, NDIM)
        self.v_coords = sm.get_ds(sm.vcoord_dsname)
        self.v_vecs = np.moveaxis(self.v_coords, 1, -1)
        if device is None:
            self.speeds = norm(self.v_vecs, axis=-1)
            self.maxspeed = float(np.mean(self.speeds) + 3 * np.std(self.speeds))
        else:
            from mdsea_b200 import _capi

            self.speeds, mean, std, _ = _capi.analyse_speeds(self.v_coords, device=int(device))
            self.maxspeed = float(mean + 3 * std)
        self.mean_pot_energies = sm.get_ds(sm.potenergy_dsname)
        self.mean_kin_energies = sm.get_ds(sm.kinenergy_dsname)
        self.total_energy = self.mean_kin_energies + self.mean_pot_energies
        self.temp = sm.get_ds(sm.temp_dsname)
        self.mean_pot_energy = np.mean(self.mean_pot_energies)
        self.mean_kin_energy = np.mean(self.mean_kin_energies)
        self.mean_total_energy = np.mean(self.total_energy)
        self.mean_temp = np.mean(self.temp)


class Vis(Analyser):
    def __init__(self, sm: SysManager, frame_step: int = 1) -> None:
        super().__init__(sm)
        self.frame_step = frame_step
        self.num_frames = int(sm.STEPS / frame_step)
        self.step = 0
