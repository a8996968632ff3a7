"""Robot constant table (host side).

Mirrors the constants the reference keeps in ``Quadruped.__init__``
(reference/MPC_Controller/common/Quadruped.py:14-94) and ``getHipLocation`` (:96-107).
Values are data, not code: masses, diagonal inertias, link lengths, hip offsets, nominal
body height, friction coefficient and the default 13 MPC weights per robot type.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from enum import IntEnum

import numpy as np


class RobotType(IntEnum):
    """Same members as reference Quadruped.RobotType (Quadruped.py:6-10); the integer
    value doubles as the row of the device constant table."""
    ALIENGO = 0
    A1 = 1
    GO1 = 2


@dataclass(frozen=True)
class RobotModel:
    name: str
    mass: float
    inertia_diag: tuple          # body inertia (diagonal, kg m^2)
    abad_link: float
    hip_link: float
    knee_link: float
    abad_location: tuple         # (x, y, z) of the front-left hip
    body_height: float
    mpc_weights: tuple           # 13 default weights
    friction: float = 0.4

    @property
    def inertia9(self) -> np.ndarray:
        return np.diag(np.asarray(self.inertia_diag, dtype=np.float64)).reshape(9)

    def hip_location(self, leg: int) -> np.ndarray:
        """reference Quadruped.getHipLocation (Quadruped.py:96-107)."""
        x, y, z = self.abad_location
        return np.array([x if leg in (0, 1) else -x, y if leg in (0, 2) else -y, z])

    def params_row(self) -> np.ndarray:
        """Row of the device constant table: mass, inertia[9] row-major, 2 pad
        (include/mpc_b200.h MPC_ROBOT_PARAMS_STRIDE)."""
        row = np.zeros(12, dtype=np.float32)
        row[0] = self.mass
        row[1:10] = self.inertia9
        return row


_W_ALIENGO = (1.0, 1.5, 0.0, 0.0, 0.0, 50.0, 0.0, 0.0, 0.1, 1.0, 1.0, 0.1, 0.0)

ROBOTS = {
    RobotType.ALIENGO: RobotModel(
        "aliengo", 9.041 * 2, (0.033260231, 0.16117211, 0.17460442),
        0.083, 0.25, 0.25, (0.2399, 0.051, 0.0), 0.35, _W_ALIENGO),
    RobotType.A1: RobotModel(
        "a1", 8.5 * 3, tuple(10 * v for v in (0.017, 0.057, 0.064)),
        0.08505, 0.2, 0.2, (0.183, 0.047, 0.0), 0.26,
        (0.25, 0.25, 10.0, 2.0, 2.0, 50.0, 0.0, 0.0, 0.3, 0.5, 0.5, 0.1, 0.0)),
    RobotType.GO1: RobotModel(
        "go1", 5.204 * 2, tuple(5 * v for v in (0.0168128557, 0.063009565, 0.0716547275)),
        0.08, 0.213, 0.213, (0.1881, 0.04675, 0.0), 0.26,
        tuple(10 * v for v in _W_ALIENGO)),
}

SIDE_SIGN = (1.0, -1.0, 1.0, -1.0)   # reference utils.py:7
GEOM_STRIDE = 24        # MPC_GEOM_STRIDE (include/mpc_b200.h)
TICK_STATE_STRIDE = 96  # MPC_TICK_STATE_STRIDE


def geom_row(m: RobotModel) -> np.ndarray:
    """Geometry / default-weight row of the device tick (MPC_GEOM_* layout)."""
    row = np.zeros(GEOM_STRIDE, dtype=np.float32)
    row[0:3] = (m.abad_link, m.hip_link, m.knee_link)
    row[3:6] = m.abad_location
    row[6], row[7] = m.body_height, m.friction
    row[8:21] = m.mpc_weights
    return row


def geom_table() -> np.ndarray:
    return np.stack([geom_row(ROBOTS[t]) for t in RobotType])


def initial_tick_state(robot_type) -> np.ndarray:
    """Fresh per-robot tick state (MPC_ST_* layout) = a just-initialised ConvexMPCLocomotion +
    StateEstimator (ConvexMPCLocomotion.py:89-114, StateEstimator.py:27-48): counter 0,
    firstRun, firstSwing x4, estimator height = nominal body height, ground normal = +z."""
    rt = np.atleast_1d(np.asarray(robot_type, dtype=np.int64))
    st = np.zeros((len(rt), TICK_STATE_STRIDE), dtype=np.float32)
    sti = st.view(np.int32)
    sti[:, 1] = 1
    sti[:, 2:6] = 1
    st[:, 70] = [ROBOTS[RobotType(int(t))].body_height for t in rt]
    st[:, 89] = 1.0
    return st


def params_table() -> np.ndarray:
    """[n_types, 12] float32 constant table in RobotType order."""
    return np.stack([ROBOTS[t].params_row() for t in RobotType])
