"""Consumer of the simfile layout (mdsea/analytics.py:44-79): loads the five datasets and derives
per-step vectors, speeds and energies.  Kept small on purpose -- it is the compatibility check for the
file layout, not an accelerated component."""
import numpy as np

from mdsea_b200.core import SysManager
from mdsea_b200.quicker import norm


class Analyser:
    def __init__(self, sm: SysManager) -> None:
        self.sm = sm
        self.r_coords = sm.get_ds(sm.rcoord_dsname)                     # (STEPS, NDIM, N)
        self.r_vecs = np.moveaxis(self.r_coords, 1, -1)                 # (STEPS, N, NDIM)
        self.v_coords = sm.get_ds(sm.vcoord_dsname)
        self.v_vecs = np.moveaxis(self.v_coords, 1, -1)
        self.speeds = norm(self.v_vecs, axis=-1)
        self.maxspeed = float(np.mean(self.speeds) + 3 * np.std(self.speeds))
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
