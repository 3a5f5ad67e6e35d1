"""SysManager -- run parameters, simfile tree and HDF5 datasets; the boundary that stays intact.

Interface and on-disk layout of mdsea/core.py: ``SysManager.new/load/initfilesys/update_ds/get_ds/delete``,
UPPER_CASE run attributes, ``simfiles/<simid>/{data.h5, settings.pkl, vis/...}``, five root float64
datasets (core.py:230-244).  New here: ``update_ds_block`` -- the solver's k-step flush writes a block of
rows per dataset instead of one h5py call per row and dataset (core.py:316-321), and the storage backend
falls back to the built-in ``hdf5min`` writer when h5py is not installed (same file format).
"""
import logging
import os
import pickle
import shutil
import time
from typing import Optional, Union

import numpy as np

from mdsea_b200.constants import DTYPE
from mdsea_b200.constants import fileman as cfm
from mdsea_b200.helpers import nsphere_volume

try:  # pragma: no cover - h5py is absent from the build image
    import h5py  # type: ignore
except ImportError:  # pragma: no cover
    from mdsea_b200 import hdf5min as h5py

log = logging.getLogger(__name__)

_SETTING_KEYS = ("radius_particle", "num_particles", "vol_fraction", "steps", "ndim", "mass", "temp", "pbc",
                 "quench_temps", "quench_steps", "quench_timings", "new_sim", "simid")


def _gen_newid() -> str:
    return str(time.time()).replace(".", "")


class SysManager:
    def __init__(self, new_sim: bool, simid: Optional[str], ndim: int, num_particles: int, vol_fraction: float,
                 radius_particle: float, temp: float = 1, mass: float = 1, pbc: bool = True, steps: int = 1,
                 quench_temps: list = [], quench_steps: list = [], quench_timings: list = []) -> None:
        # run parameters (core.py:64-90); the settings dict is what settings.pkl holds (13 keys)
        given = dict(radius_particle=radius_particle, num_particles=num_particles, vol_fraction=vol_fraction,
                     steps=steps, ndim=ndim, mass=mass, temp=temp, pbc=pbc, quench_temps=quench_temps,
                     quench_steps=quench_steps, quench_timings=quench_timings, new_sim=new_sim, simid=simid)
        self._settings = {k: given[k] for k in _SETTING_KEYS}
        self.RADIUS_PARTICLE, self.NUM_PARTICLES, self.VOL_FRACTION = radius_particle, num_particles, vol_fraction
        self.STEPS, self.NDIM, self.MASS, self.TEMP, self.PBC = steps, ndim, mass, temp, pbc
        self.QUENCH_T = quench_temps
        self.QUENCH_TIMING = quench_timings
        self.QUENCH_STEP = [round(self.STEPS * t) for t in self.QUENCH_TIMING]  # core.py:86

        self.SIM_ID, self.new_sim = simid, new_sim
        if self.SIM_ID is None:
            self.new_sim = True
            self.SIM_ID = _gen_newid()

        # box edge from the packing fraction (core.py:113-115)
        box_volume = self.NUM_PARTICLES * nsphere_volume(self.NDIM, self.RADIUS_PARTICLE) / self.VOL_FRACTION
        self.LEN_BOX = box_volume ** (1 / self.NDIM)

        # paths (constants/fileman.py)
        sid = self.SIM_ID
        self.sim_path = cfm.SIM_PATH.format(sid)
        self.dir_tree = tuple(p.format(sid) for p in (cfm.SIM_PATH, cfm.VIS_PATH, cfm.PNG_PATH, cfm.MP4_PATH, cfm.BLENDER_PATH))
        self.settings_path = cfm.SETTINGS_PATH.format(sid)
        self.blender_path = cfm.BLENDER_PATH.format(sid)
        self.png_path = cfm.PNG_PATH.format(sid)
        self.mp4_path = cfm.MP4_PATH.format(sid)
        self.data_path = cfm.DATA_PATH.format(sid)

        # dataset names, in the order update_files writes them (core.py:174-185)
        self.rcoord_dsname, self.vcoord_dsname = cfm.RCOORDS_DSNAME, cfm.VCOORDS_DSNAME
        self.coord_dsnames = [self.rcoord_dsname, self.vcoord_dsname]
        self.potenergy_dsname, self.kinenergy_dsname, self.temp_dsname = (
            cfm.POTENERGY_DSNAME, cfm.KINENERGY_DSNAME, cfm.TEMP_DSNAME)
        self.oned_dsnames = [self.potenergy_dsname, self.kinenergy_dsname, self.temp_dsname]
        self.all_dsnames = [*self.coord_dsnames, *self.oned_dsnames]

        self.dfile = None
        if not self.new_sim:
            self._open_datafile()

    # -- file system -------------------------------------------------------------------------------
    def _create_tree(self) -> None:
        simfiles = os.path.join(os.getcwd(), "simfiles")
        if not os.path.exists(simfiles):
            os.mkdir(simfiles)
            log.debug("Couldn't find a 'simulations-folder'. A new one was created: %s", simfiles)
        for directory in self.dir_tree:
            if os.path.exists(directory):
                log.warning("Directory exists and will be overwritten: %s", directory)
                shutil.rmtree(directory)
            os.mkdir(directory)
        log.debug("New directory tree created: %s", self.sim_path)

    def _open_datafile(self, mode: str = "a", shared: Optional[bool] = None):
        """Open data.h5.  ``shared=True`` (multi-GPU runs: every rank writes the columns of the particles it owns into the
        same file) always goes through the built-in ``hdf5min`` backend, whose datasets are shared memory maps: several
        non-MPI processes holding the file open read-write is exactly what real h5py / libhdf5 file locking forbids."""
        if shared is not None:
            self._shared_file = bool(shared)
        if getattr(self, "_shared_file", False):
            from mdsea_b200 import hdf5min

            self.dfile = hdf5min.File(self.data_path, mode=mode)
        else:
            self.dfile = h5py.File(self.data_path, mode=mode)
        return self.dfile

    def _close_datafile(self):
        try:
            self.dfile.close()
        except AttributeError:
            log.warning("You tried to close a data file (%s) that was already closed.", self.data_path)

    def _init_datasets(self) -> None:
        if self.dfile is None:
            raise FileNotFoundError(f"Datafile does not exist! Expected at: {self.data_path}")
        for name in self.coord_dsnames:
            self.dfile.create_dataset(name, dtype=DTYPE, shape=(self.STEPS, self.NDIM, self.NUM_PARTICLES))
        for name in self.oned_dsnames:
            self.dfile.create_dataset(name, dtype=DTYPE, shape=(self.STEPS,))
        log.debug("HDF datasets initiated/created.")

    def _save_settings(self) -> None:
        with open(self.settings_path, "wb") as f:
            pickle.dump(self._settings, f, pickle.HIGHEST_PROTOCOL)

    # -- constructors ------------------------------------------------------------------------------
    @classmethod
    def new(cls, simid: Optional[str] = None, initfilesys: bool = True, **kwargs) -> "SysManager":
        if simid is None:
            simid = _gen_newid()
        log.info("New simulation: %s", {"id": simid, "path": cfm.SIM_PATH.format(simid)})
        sm = cls(new_sim=True, simid=simid, **kwargs)
        if initfilesys:
            sm.initfilesys()
        return sm

    @classmethod
    def load(cls, simid: str) -> "SysManager":
        path_sim, path_settings = cfm.SIM_PATH.format(simid), cfm.SETTINGS_PATH.format(simid)
        if not (os.path.exists(path_sim) and os.path.exists(path_settings)):
            raise FileNotFoundError(
                f"Could not load simulation: { {'id': simid, 'path': path_sim, 'settings': path_settings} }")
        log.info("Loading simulation: %s", simid)
        with open(path_settings, "rb") as f:
            kwargs: dict = pickle.load(f)
        kwargs.pop("new_sim")
        return cls(new_sim=False, **kwargs)

    # -- public --------------------------------------------------------------------------------------
    def initfilesys(self) -> None:
        if not self.new_sim:
            raise SystemExit("You were about to initialize an existing simulation, which is not allowed. "
                             f"Try loading this simulation: { {'sim ID': self.SIM_ID, 'new simulation': self.new_sim} }")
        self._create_tree()
        self._save_settings()
        self._open_datafile()
        self._init_datasets()
        self._close_datafile()  # the reference leaves the file CLOSED until the first update_ds (core.py:314)

    def update_ds(self, dsname: str, x: Union[list, np.ndarray], i: int) -> None:
        """Write row i of a dataset; error behaviour of core.py:316-344 (reopen-and-retry on a closed file)."""
        try:
            self.dfile[dsname][i] = x
        except TypeError as e:
            if self.dfile is None:
                raise FileNotFoundError("The dataset's datafile you are trying to update might be closed or not exist.")
            raise e
        except ValueError as e:
            if self.dfile.id.valid:
                raise ValueError(f"Could not update dataset: {e}")
            log.warning("A ValueError ('%s') was caught while trying to update a dataset. However, we noticed that the "
                        "datafile in question wasn't open. We'll try to fix this and retry...", e)
            self._open_datafile()
            self.update_ds(dsname, x, i)
        except Exception as e:
            self._close_datafile()
            raise e

    def update_ds_block(self, dsname: str, block: np.ndarray, i0: int) -> None:
        """Rows i0 .. i0+len(block)-1 of a dataset in one assignment (the solver's buffered flush)."""
        if self.dfile is None:
            raise FileNotFoundError("The dataset's datafile you are trying to update might be closed or not exist.")
        if not self.dfile.id.valid:
            log.warning("The datafile wasn't open when a block of rows arrived; reopening it.")
            self._open_datafile()
        try:
            self.dfile[dsname][i0:i0 + len(block)] = block
        except Exception as e:
            self._close_datafile()
            raise e

    def get_ds(self, dsname: str) -> np.ndarray:
        try:
            return self.dfile[dsname][:]
        except Exception as e:
            self._close_datafile()
            raise e

    def delete(self) -> None:
        shutil.rmtree(self.sim_path, ignore_errors=True)
        log.info("Directory tree deleted: %s", self.sim_path)
