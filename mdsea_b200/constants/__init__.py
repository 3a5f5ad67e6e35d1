"""Constants shared with mdsea (mdsea/constants/__init__.py:10,17)."""
import numpy

DTYPE = numpy.float64
SPRTD_ANIMPLATFORMS = ("mpl", "blender", "vapory", "vpy", "mayavi")
