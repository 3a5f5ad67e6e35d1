"""Simfile tree and dataset names -- the on-disk contract of mdsea/constants/fileman.py:8-47.

simfiles/<simid>/{data.h5, settings.pkl, vis/{png,mp4,blender}/}, rooted at the working directory
in effect when this module is first imported (the reference freezes it the same way).
"""
import os

DIR_SIMFILES = os.path.join(os.getcwd(), "simfiles")
SIM_PATH = DIR_SIMFILES + "/{}"

VIS_DNAME, PNG_DNAME, MP4_DNAME, BLENDER_DNAME = "vis", "png", "mp4", "blender"
VIS_PATH = SIM_PATH + "/" + VIS_DNAME
PNG_PATH = VIS_PATH + "/" + PNG_DNAME
MP4_PATH = VIS_PATH + "/" + MP4_DNAME
BLENDER_PATH = VIS_PATH + "/" + BLENDER_DNAME

DATA_FNAME = "data.h5"
DATA_PATH = SIM_PATH + "/" + DATA_FNAME
SETTINGS_FNAME = "settings.pkl"
SETTINGS_PATH = SIM_PATH + "/" + SETTINGS_FNAME

# (STEPS, NDIM, NUM_PARTICLES)
RCOORDS_DSNAME = "r_coords"
VCOORDS_DSNAME = "v_coords"
# (STEPS,)
POTENERGY_DSNAME = "pot_energy"
KINENERGY_DSNAME = "kin_energy"
TEMP_DSNAME = "temp"
