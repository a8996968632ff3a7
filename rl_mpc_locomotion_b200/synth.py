"""Seeded synthetic problem batches (SURVEY.md §8d "Synthetic inputs").

The same float32 values go to the CPU oracle (promoted to double) and to the GPU.  The
distributions follow the reference's own operating ranges: command ranges from
reference/RL_Environment/cfg/task/Aliengo.yaml:24-27, RL weight ranges from
reference/MPC_Controller/Parameters.py:25-33 (`U(-1,1)*scale + const`, 13th weight 0 as
in reference/RL_Environment/tasks/aliengo.py:253), robot constants from
reference/MPC_Controller/common/Quadruped.py:16-73, gait tables from
reference/MPC_Controller/convex_MPC/Gait.py:69-84 evaluated at a solve tick (fractional
iteration k+0.5).
"""
from __future__ import annotations

import numpy as np

from . import gait as _gait
from .records import pack_records
from .robots import ROBOTS, SIDE_SIGN, RobotType, params_table

MPC_PARAM_SCALE = np.array([4, 4, 4, 20, 20, 20, 1, 1, 1, 1, 1, 1], dtype=np.float64)
MPC_PARAM_CONST = np.array([5, 5, 5, 50, 50, 50, 1, 1, 1, 1, 1, 1], dtype=np.float64)


def make_batch(n: int, horizon: int = 10, seed: int = 0, gaits=(0,), robot_types=(RobotType.ALIENGO,),
               flat_ground: bool = False):
    """Return ``(records[N,stride] f32, robot_type[N] i32, params[T,12] f32, meta)``.

    ``gaits`` / ``robot_types``: each robot draws uniformly from the given tuples
    (BASELINE configs 2/3/4: trot only; trot/walk/bound; Go1/A1/Aliengo).
    """
    rng = np.random.default_rng(seed)
    rtypes = np.asarray([int(t) for t in robot_types], dtype=np.int32)
    rt = rtypes[rng.integers(0, len(rtypes), size=n)] if len(rtypes) > 1 else np.full(n, rtypes[0], np.int32)
    gids = np.asarray(gaits, dtype=np.int64)
    gid = gids[rng.integers(0, len(gids), size=n)] if len(gids) > 1 else np.full(n, gids[0], np.int64)

    body_h = np.array([ROBOTS[RobotType(t)].body_height for t in rt])
    abad = np.array([ROBOTS[RobotType(t)].abad_link for t in rt])
    hips = np.stack([[ROBOTS[RobotType(t)].hip_location(l) for l in range(4)] for t in rt])  # [N,4,3]
    defw = np.stack([ROBOTS[RobotType(t)].mpc_weights for t in rt]).astype(np.float64)
    mu = np.array([ROBOTS[RobotType(t)].friction for t in rt])

    rpy = np.stack([rng.uniform(-0.2, 0.2, n), rng.uniform(-0.2, 0.2, n), rng.uniform(-0.3, 0.3, n)], 1)
    z = rng.uniform(0.25, 0.40, n) * (body_h / 0.35)
    pos = np.stack([np.zeros(n), np.zeros(n), z], 1)
    omega = rng.uniform(-1, 1, (n, 3))
    vel = np.stack([rng.uniform(-2.5, 2.5, n), rng.uniform(-1, 1, n), rng.uniform(-0.3, 0.3, n)], 1)
    if flat_ground:
        normal = np.tile([0.0, 0.0, 1.0], (n, 1))
        rng.normal(0, 0.05, (n, 2))  # keep the stream aligned
    else:
        nn = np.concatenate([rng.normal(0, 0.05, (n, 2)), np.ones((n, 1))], 1)
        normal = nn / np.linalg.norm(nn, axis=1, keepdims=True)
    feet = hips.copy()
    feet[:, :, 1] += np.asarray(SIDE_SIGN)[None, :] * abad[:, None]
    feet[:, :, 0:2] += rng.uniform(-0.1, 0.1, (n, 4, 2))
    feet[:, :, 2] = -z[:, None] + rng.uniform(-0.02, 0.02, (n, 4))

    use_rl = rng.random(n) < 0.5
    rlw = np.concatenate([rng.uniform(-1, 1, (n, 12)) * MPC_PARAM_SCALE + MPC_PARAM_CONST,
                          np.zeros((n, 1))], 1)
    weights = np.where(use_rl[:, None], rlw, defw)

    des_vel = np.stack([rng.uniform(-2.5, 2.5, n), rng.uniform(-1, 1, n), np.zeros(n)], 1)
    des_angvel = np.stack([np.zeros(n), np.zeros(n), rng.uniform(-2.5, 2.5, n)], 1)
    des_pos = np.stack([np.zeros(n), np.zeros(n), body_h], 1)
    des_rpy = np.zeros((n, 3))

    iteration = rng.integers(0, 10, n) + 0.5
    table = _gait.mpc_table(gid, iteration, horizon)

    rec = pack_records(
        horizon,
        qp_weights=weights, com_position=pos, com_velocity=vel, com_roll_pitch_yaw=rpy,
        ground_normal_vec=normal, com_angular_velocity=omega,
        foot_positions_body_frame=feet.reshape(n, 12), foot_friction_coeffs=np.repeat(mu[:, None], 4, 1),
        desired_com_position=des_pos, desired_com_velocity=des_vel,
        desired_com_roll_pitch_yaw=des_rpy, desired_com_angular_velocity=des_angvel,
        foot_contact_states=table)
    meta = {"gait": gid, "iteration": iteration, "use_rl_weights": use_rl}
    return rec, rt.astype(np.int32), params_table(), meta
