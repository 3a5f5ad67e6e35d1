"""mdsea_b200 -- B200-native (sm_100a) implementation of mdsea's ContinuousPotentialSolver step loop.

Same Python surface as tpvasconcelos/mdsea for this path::

    from mdsea_b200.core import SysManager
    from mdsea_b200.potentials import Potential
    from mdsea_b200.simulator import ContinuousPotentialSolver

    sm = SysManager.new(simid="demo", ndim=2, num_particles=36, steps=500, vol_fraction=0.2, radius_particle=0.5)
    sim = ContinuousPotentialSolver(sm, pot=Potential.lennardjones(1, 1))
    sim.run_simulation()

The step loop runs in hand-written CUDA kernels behind a C ABI (include/mdsea_b200.h);
there is no CPU fallback -- without the built library or without a GPU the solver raises.
"""
from mdsea_b200._version import __version__

__all__ = ["__version__"]
