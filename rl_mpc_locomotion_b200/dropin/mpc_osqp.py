"""``import mpc_osqp`` shim: put this directory on sys.path and the reference's
``ConvexMPCLocomotion`` (ConvexMPCLocomotion.py:17-22) imports the B200 solver unchanged."""
from rl_mpc_locomotion_b200.mpc_osqp import *  # noqa: F401,F403
from rl_mpc_locomotion_b200.mpc_osqp import ConvexMpc, QPSolverName, OSQP, QPOASES, TEST, __version__  # noqa: F401
