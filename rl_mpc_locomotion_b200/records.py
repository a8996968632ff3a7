"""Problem-record packing (host side).

One record = the 13 arguments of the reference's ``compute_contact_forces``
(reference/MPC_Controller/convex_MPC/mpc_osqp.cc:578-591) in argument order, followed by
the flattened horizon contact table.  Offsets are the ``MPC_REC_*`` macros of
``include/mpc_b200.h``; keep the two in sync (tests/test_host.py checks them).
"""
from __future__ import annotations

import numpy as np

REC_WEIGHTS = 0
REC_COM_POS = 13
REC_COM_VEL = 16
REC_RPY = 19
REC_NORMAL = 22
REC_ANGVEL = 25
REC_FOOTPOS = 28
REC_MU = 40
REC_DES_POS = 44
REC_DES_VEL = 47
REC_DES_RPY = 50
REC_DES_ANGVEL = 53
REC_CONTACT = 56
ROBOT_PARAMS_STRIDE = 12
MAX_HORIZON = 16

# (name, offset, length) in reference argument order; contact table handled separately
FIELDS = (
    ("qp_weights", REC_WEIGHTS, 13),
    ("com_position", REC_COM_POS, 3),
    ("com_velocity", REC_COM_VEL, 3),
    ("com_roll_pitch_yaw", REC_RPY, 3),
    ("ground_normal_vec", REC_NORMAL, 3),
    ("com_angular_velocity", REC_ANGVEL, 3),
    ("foot_positions_body_frame", REC_FOOTPOS, 12),
    ("foot_friction_coeffs", REC_MU, 4),
    ("desired_com_position", REC_DES_POS, 3),
    ("desired_com_velocity", REC_DES_VEL, 3),
    ("desired_com_roll_pitch_yaw", REC_DES_RPY, 3),
    ("desired_com_angular_velocity", REC_DES_ANGVEL, 3),
)


def rec_stride(horizon: int) -> int:
    """Record stride in floats (MPC_REC_STRIDE): 56 + 4h rounded up to 16 bytes."""
    return (REC_CONTACT + 4 * horizon + 3) & ~3


def pack_records(horizon: int, *, foot_contact_states, **fields) -> np.ndarray:
    """Pack keyword arrays (each ``[N, k]`` or ``[k]``) into ``[N, stride]`` float32."""
    n = 1
    for v in list(fields.values()) + [foot_contact_states]:
        a = np.asarray(v)
        if a.ndim == 2:
            n = max(n, a.shape[0])
    rec = np.zeros((n, rec_stride(horizon)), dtype=np.float32)
    for name, off, ln in FIELDS:
        a = np.asarray(fields[name], dtype=np.float32).reshape(-1, ln)
        rec[:, off:off + ln] = a
    c = np.asarray(foot_contact_states, dtype=np.float32).reshape(-1, 4 * horizon)
    rec[:, REC_CONTACT:REC_CONTACT + 4 * horizon] = c
    return rec


def unpack_record(rec: np.ndarray, horizon: int) -> dict:
    """Inverse of :func:`pack_records` for ONE record -> dict of float64 arrays in the
    reference's argument names (plus ``foot_contact_states``)."""
    rec = np.asarray(rec, dtype=np.float64)
    out = {name: rec[off:off + ln].copy() for name, off, ln in FIELDS}
    out["foot_contact_states"] = rec[REC_CONTACT:REC_CONTACT + 4 * horizon].copy()
    return out
