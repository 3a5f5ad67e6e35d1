"""Simfile tree and dataset names -- the on-disk contract of mdsea/constants/fileman.py:8-47.

simfiles/<simid>/{data.h5, settings.pkl, vis/{png,mp4,blender}/}, rooted at the working directory
in effect when this module is first imported (the reference freezes it the same way).  Every
``*_PATH`` is a ``str.format`` template whose single field is the simulation id.
"""
import os


def _below(parent: str, leaf: str) -> str:
    return parent + "/" + leaf


# leaf names
VIS_DNAME, PNG_DNAME, MP4_DNAME, BLENDER_DNAME = "vis", "png", "mp4", "blender"
DATA_FNAME, SETTINGS_FNAME = "data.h5", "settings.pkl"

# templates
DIR_SIMFILES = os.path.join(os.getcwd(), "simfiles")
SIM_PATH = _below(DIR_SIMFILES, "{}")
VIS_PATH = _below(SIM_PATH, VIS_DNAME)
PNG_PATH, MP4_PATH, BLENDER_PATH = (_below(VIS_PATH, leaf) for leaf in (PNG_DNAME, MP4_DNAME, BLENDER_DNAME))
DATA_PATH = _below(SIM_PATH, DATA_FNAME)
SETTINGS_PATH = _below(SIM_PATH, SETTINGS_FNAME)

# datasets of data.h5: two of shape (STEPS, NDIM, NUM_PARTICLES), three of shape (STEPS,)
RCOORDS_DSNAME, VCOORDS_DSNAME = "r_coords", "v_coords"
POTENERGY_DSNAME, KINENERGY_DSNAME, TEMP_DSNAME = "pot_energy", "kin_energy", "temp"
