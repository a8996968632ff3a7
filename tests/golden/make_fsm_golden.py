"""Golden FSM-shell fixture from the UNMODIFIED reference Python.

Drives the reference's own ``RobotRunnerFSM`` (reference/MPC_Controller/robot_runner/
RobotRunnerFSM.py:55-69 -> ControlFSM.runFSM -> Passive / RecoveryStand / Locomotion states) on a
simulator-free kinematic replay, with the oracle's ``mpc_osqp`` module standing in for the native
one, through a scripted sequence of mode requests and orientation upsets:

  RecoveryStand (stand-up from the constructor's q = 0) -> request LOCOMOTION -> trot ->
  roll beyond 40 deg (locomotionSafe() fails -> RecoveryStand, the main loop's fake gamepad event
  RL_MPC_Locomotion.py:103-105 is mimicked by setting Parameters.control_mode) -> flipped upside
  down during StandUp -> FoldLegs -> RollOver -> FoldLegs -> request PASSIVE -> "bad request"
  LOCOMOTION from Passive -> RecoveryStand -> LOCOMOTION again.

Logs, per tick, the inputs, the requested control mode, the FSM state / operating mode after the
tick and the joint torques.  Run in the build container only (needs /root/reference).
"""
import contextlib
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
os.environ["MPC_ORACLE_MODE"] = "cold"
sys.path.insert(0, os.path.join(ROOT, "oracle"))   # provides `import mpc_osqp`
sys.path.insert(0, "/root/reference")
sys.path.insert(0, HERE)

with contextlib.redirect_stdout(io.StringIO()):
    from MPC_Controller.Parameters import Parameters
    from MPC_Controller.common.Quadruped import RobotType
    from MPC_Controller.robot_runner.RobotRunnerFSM import RobotRunnerFSM
    from MPC_Controller.utils import FSM_OperatingMode, FSM_StateName, GaitType
from make_tick_golden import quat_xyzw  # noqa: E402

Parameters.bridge_MPC_to_RL = True     # array inputs (LegController.py:96-98, StateEstimator.py:58-71)
assert Parameters.control_mode is FSM_StateName.RECOVERY_STAND and Parameters.FSM_check_safety
assert Parameters.operatingMode is FSM_OperatingMode.NORMAL


def main(name="aliengo", robot=RobotType.ALIENGO, ticks=420, seed=5):
    rng = np.random.default_rng(seed)
    Parameters.cmpc_gait = GaitType.TROT
    Parameters.flat_ground = False
    Parameters.control_mode = FSM_StateName.RECOVERY_STAND
    Parameters.locomotionUnsafe = False
    sink = io.StringIO()
    with contextlib.redirect_stdout(sink):
        runner = RobotRunnerFSM()
        runner.init(robot)
    # scripted requests: tick -> control mode ("what the gamepad would set")
    script = {70: FSM_StateName.LOCOMOTION, 300: FSM_StateName.PASSIVE, 330: FSM_StateName.LOCOMOTION,
              345: FSM_StateName.RECOVERY_STAND, 380: FSM_StateName.LOCOMOTION}
    log = {k: [] for k in ("dof", "body", "cmd", "torques", "state", "op_mode", "request", "unsafe")}
    stand = np.tile([0.0, 0.8, -1.6], 4)
    ph = rng.uniform(0, 6.28, 12)
    w0 = rng.uniform(3.0, 9.0, 12)
    for t in range(ticks):
        if t in script:
            Parameters.control_mode = script[t]
        tt = 0.01 * t
        q = stand + 0.1 * np.sin(w0 * tt + ph)
        qd = 0.1 * w0 * np.cos(w0 * tt + ph)
        roll, pitch = 0.05 * np.sin(3.1 * tt + 0.3), 0.04 * np.sin(2.3 * tt + 1.1)
        if 110 <= t < 150:
            roll = 0.8            # > 40 deg: locomotionSafe() fails
        if 150 <= t < 250:
            roll = 3.0            # upside down: rBody[2,2] < 0
        if 395 <= t:
            pitch = -0.75         # |pitch| > 40 deg
        body = np.concatenate([[0.0, 0.0, 0.35], quat_xyzw(roll, pitch, 0.2 * np.sin(0.9 * tt)),
                               [0.4 + 0.1 * np.sin(2 * tt), 0.05 * np.cos(3 * tt), 0.02 * np.sin(5 * tt)],
                               [0.2 * np.sin(4 * tt), 0.15 * np.cos(3.3 * tt), 0.3 * np.sin(1.7 * tt)]]).astype(np.float32)
        dof = np.stack([q, qd], 1).astype(np.float32)
        cmd = np.asarray([0.5, 0.05, 0.3], dtype=np.float32)
        with contextlib.redirect_stdout(sink):
            tau = runner.run(dof, body, cmd)
        unsafe = bool(Parameters.locomotionUnsafe)
        if unsafe:   # RL_MPC_Locomotion.py:103-105: the main loop requests recovery
            Parameters.control_mode = FSM_StateName.RECOVERY_STAND
            Parameters.locomotionUnsafe = False
        log["dof"].append(dof); log["body"].append(body); log["cmd"].append(cmd)
        log["torques"].append(np.asarray(tau, np.float64))
        log["state"].append(runner._controlFSM.currentState.stateName.value)
        log["op_mode"].append(runner._controlFSM.operatingMode.value)
        log["request"].append(Parameters.control_mode.value)
        log["unsafe"].append(unsafe)
    out = {k: np.asarray(v) for k, v in log.items()}
    out["script_ticks"] = np.asarray(sorted(script))
    out["script_modes"] = np.asarray([script[k].value for k in sorted(script)])
    np.savez_compressed(os.path.join(HERE, f"fsm_golden_{name}.npz"), **out)
    st = out["state"]
    print(name, "ticks", ticks, "states visited", sorted(set(st.tolist())), "unsafe ticks", int(out["unsafe"].sum()),
          "max |tau|", float(np.abs(out["torques"]).max()))
    Parameters.control_mode = FSM_StateName.RECOVERY_STAND


if __name__ == "__main__":
    main()
