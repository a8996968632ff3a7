"""Golden Locomotion-tick fixtures from the UNMODIFIED reference Python.

Drives the reference's own objects (RobotRunnerMin wiring, reference/MPC_Controller/
robot_runner/RobotRunnerMin.py:54-76) on a simulator-free kinematic replay, with the oracle's
`mpc_osqp` pybind module (oracle/mpc_osqp*.so: reference OSQP, COLD mode = fresh workspace
per solve) standing in for the native module (`--refmod`: the reference's own mpc_osqp.cc compiled
unmodified against oracle/eigen_standin, OSQP branch).  Logs, per tick, the
inputs, the StateEstimator.update() outputs, the 13 QP arguments when the MPC fires, and the
resulting leg torques.  Run in the build container only (needs /root/reference).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
os.environ["MPC_ORACLE_MODE"] = "warm" if "--warm" in sys.argv else "cold"
sys.path.insert(0, os.path.join(ROOT, "oracle"))   # provides `import mpc_osqp`
if "--refmod" in sys.argv:   # ... or the reference's OWN mpc_osqp.cc compiled unmodified (oracle/Makefile, eigen_standin):
    sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref", "refmod"))   # persistent workspace = the warm runs
sys.path.insert(0, "/root/reference")

import contextlib, io  # noqa: E402
with contextlib.redirect_stdout(io.StringIO()):
    from MPC_Controller.Parameters import Parameters
    from MPC_Controller.common.Quadruped import RobotType
    from MPC_Controller.robot_runner.RobotRunnerMin import RobotRunnerMin
    from MPC_Controller.utils import GaitType

Parameters.bridge_MPC_to_RL = True     # array inputs, as the RL env uses (aliengo.py:247-256)


def quat_xyzw(roll, pitch, yaw):
    cy, sy, cp, sp, cr, sr = np.cos(yaw / 2), np.sin(yaw / 2), np.cos(pitch / 2), np.sin(pitch / 2), np.cos(roll / 2), np.sin(roll / 2)
    w = cr * cp * cy + sr * sp * sy
    x = sr * cp * cy - cr * sp * sy
    y = cr * sp * cy + sr * cp * sy
    z = cr * cp * sy - sr * sp * cy
    return np.array([x, y, z, w])


def run_case(name, robot, gait, flat, ticks, seed, with_weights, compact=False):
    rng = np.random.default_rng(seed)
    Parameters.cmpc_gait = gait
    Parameters.flat_ground = flat
    with contextlib.redirect_stdout(io.StringIO()):
        runner = RobotRunnerMin()
        runner.init(robot)
    log = {k: [] for k in ("dof", "body", "cmd", "vBody", "omegaBody", "rpyBody", "groundR", "z_pre", "z_post",
                           "normal", "contact_phase", "torques", "f_ff", "counter", "solved", "qp_args", "qp_out")}
    calls = []
    if "--refmod" in sys.argv:   # the reference's Python asks for mpc.QPOASES (ConvexMPCLocomotion.py:108), whose
        # source is absent; the north-star path is the OSQP branch of the same file: same constructor arguments
        import mpc_osqp as _mpc
        q = runner.data._quadruped
        runner.cMPC._cpp_mpc = _mpc.ConvexMpc(q._bodyMass, list(q._bodyInertia), 4, runner.cMPC.horizonLength,
                                              runner.cMPC.dtMPC, Parameters.cmpc_alpha, _mpc.OSQP)
    real = runner.cMPC._cpp_mpc.compute_contact_forces

    class Proxy:
        def compute_contact_forces(self, *a):
            out = real(*a)
            calls.append(([np.asarray(x, dtype=np.float64).reshape(-1) for x in a], np.asarray(out, dtype=np.float64)))
            return out

        def reset_solver(self):
            pass
    runner.cMPC._cpp_mpc = Proxy()
    stand = np.tile([0.0, 0.8, -1.6], 4)
    ph = rng.uniform(0, 6.28, 12)
    w0 = rng.uniform(3.0, 9.0, 12)
    for t in range(ticks):
        tt = 0.01 * t
        q = stand + 0.12 * np.sin(w0 * tt + ph)
        qd = 0.12 * w0 * np.cos(w0 * tt + ph)
        dof = np.stack([q, qd], 1).astype(np.float32)
        rpy = np.array([0.06 * np.sin(3.1 * tt + 0.3), 0.05 * np.sin(2.3 * tt + 1.1), 0.2 * np.sin(0.9 * tt)])
        body = np.concatenate([[0.0, 0.0, 0.35], quat_xyzw(*rpy), [0.5 + 0.1 * np.sin(2 * tt), 0.05 * np.cos(3 * tt), 0.02 * np.sin(5 * tt)],
                               [0.2 * np.sin(4 * tt), 0.15 * np.cos(3.3 * tt), 0.3 * np.sin(1.7 * tt)]]).astype(np.float32)
        cmd = [0.5, 0.05, 0.3]
        if with_weights:
            cmd = cmd + list(rng.uniform(-1, 1, 12) * np.array(Parameters.MPC_param_scale) + np.array(Parameters.MPC_param_const)) + [0.0]
        cmd = np.asarray(cmd, dtype=np.float32)
        # --- RobotRunnerMin.run, step by step on the reference's own objects
        runner._desiredStateCommand.updateCommand(cmd)
        runner._legController.updateData(dof)
        runner._legController.zeroCommand()
        runner._stateEstimator.update(body)
        se = runner._stateEstimator.getResult()
        log["vBody"].append(np.asarray(se.vBody, np.float64).reshape(3))
        log["omegaBody"].append(np.asarray(se.omegaBody, np.float64).reshape(3))
        log["rpyBody"].append(np.asarray(se.rpyBody, np.float64).reshape(3))
        log["groundR"].append(np.asarray(runner._stateEstimator.ground_R_body_frame, np.float64).reshape(9))
        log["z_pre"].append(float(se.position[2]))
        n_before = len(calls)
        runner.cMPC.run(runner.data)
        tau = runner._legController.updateCommand()
        log["dof"].append(dof); log["body"].append(body)
        log["cmd"].append(np.concatenate([cmd, np.zeros(16 - len(cmd), np.float32)]))
        log["z_post"].append(float(se.position[2]))
        log["normal"].append(np.asarray(se.ground_normal_yaw, np.float64).reshape(3))
        log["contact_phase"].append(np.asarray(runner._stateEstimator._contactPhase, np.float64).reshape(4))
        log["torques"].append(np.asarray(tau, np.float64))
        log["f_ff"].append(np.asarray(runner.cMPC.f_ff, np.float64).reshape(12))
        log["counter"].append(runner.cMPC.iterationCounter)
        solved = len(calls) > n_before
        log["solved"].append(solved)
        if solved:
            a, o = calls[-1]
            log["qp_args"].append(np.concatenate(a)); log["qp_out"].append(o)
        else:
            log["qp_args"].append(np.zeros(96)); log["qp_out"].append(np.zeros(120))
    out = {k: np.asarray(v) for k, v in log.items()}
    out["meta"] = np.array([int(robot.value), int(gait.value), int(flat), int(with_weights)])
    if compact == "return":
        return out
    if compact:   # long replays: float32 storage (every logged quantity is float32 / float16 in the reference
        # except the solver's float64 output, whose float32 rounding is 4e-6 N), QP data only on MPC ticks
        sel = out["solved"].astype(bool)
        out = {"dof": out["dof"].astype(np.float32), "body": out["body"].astype(np.float32),
               "cmd": out["cmd"].astype(np.float32), "torques": out["torques"].astype(np.float32),
               "counter": out["counter"].astype(np.int32), "solved": sel, "z_post": out["z_post"].astype(np.float32),
               "contact_phase": out["contact_phase"].astype(np.float32), "f_ff": out["f_ff"].astype(np.float32),
               "qp_args": out["qp_args"][sel].astype(np.float32), "qp_out": out["qp_out"][sel].astype(np.float32),
               "meta": out["meta"]}
    np.savez_compressed(os.path.join(HERE, f"tick_golden_{name}.npz"), **out)
    print(name, "ticks", ticks, "qp solves", int(out["solved"].sum()), "max |tau|", np.abs(out["torques"]).max())


if __name__ == "__main__":
    if "--check" in sys.argv:   # with --refmod: the committed WARM fixtures (recorded with the oracle module) against
        # a replay through the unmodified reference Python AND the reference's own C++ source
        g = np.load(os.path.join(HERE, "tick_golden_aliengo_trot_warm.npz"))
        r = run_case("aliengo_trot_warm", RobotType.ALIENGO, GaitType.TROT, False, 60, 1, True, compact="return")
        worst = {k: float(np.abs(np.asarray(r[k], np.float64) - g[k]).max()) for k in ("torques", "qp_out", "qp_args", "f_ff")}
        print("aliengo_trot_warm", worst)
        assert worst["qp_args"] < 1e-6 and worst["qp_out"] < 1e-4 and worst["torques"] < 1e-5, worst
        if "--config1" in sys.argv:
            g = np.load(os.path.join(HERE, "tick_golden_config1_aliengo_trot_1000.npz"))
            r = run_case("config1", RobotType.ALIENGO, GaitType.TROT, False, 1000, 11, True, compact="return")
            sel = r["solved"].astype(bool)
            worst = {"torques": float(np.abs(r["torques"] - g["torques_warm"]).max()),
                     "qp_out": float(np.abs(r["qp_out"][sel] - g["qp_out_warm"]).max()),
                     "qp_args": float(np.abs(r["qp_args"][sel] - g["qp_args"]).max())}
            e = np.abs(r["qp_out"][sel] - g["qp_out_warm"]).max(1)
            print("config1 (1000 ticks, warm)", worst, "qp_out error percentiles 50/99:", np.percentile(e, [50, 99]))
            # the warm-started trajectory carries OSQP's own rounding sensitivity from solve to solve (1e-4..5e-4 N
            # per near-singular polish, tests/test_oracle.py::test_reference_self_noise_*); fixture stored as float32
            assert worst["qp_args"] < 1e-5 and worst["qp_out"] < 5e-3 and worst["torques"] < 1e-3, worst
        sys.exit(0)
    if "--config1" in sys.argv:   # BASELINE config 1: 1 Aliengo, trot, h = 10, 1000 ticks of RobotRunnerMin, twice:
        # with a fresh OSQP workspace per solve (cold: every solve is an independent function of its 13
        # arguments) and with the reference's persistent workspace (warm: what mpc_osqp.cc:757-780 does)
        runs = {}
        for mode in ("cold", "warm"):
            os.environ["MPC_ORACLE_MODE"] = mode       # read by the oracle module's ConvexMpc constructor
            runs[mode] = run_case("config1", RobotType.ALIENGO, GaitType.TROT, False, 1000, 11, True, compact="return")
        c, w = runs["cold"], runs["warm"]
        sel = c["solved"].astype(bool)
        assert np.array_equal(c["qp_args"], w["qp_args"]) and np.array_equal(sel, w["solved"].astype(bool))
        f32 = lambda a: np.asarray(a, dtype=np.float32)
        out = {"dof": f32(c["dof"]), "body": f32(c["body"]), "cmd": f32(c["cmd"]), "counter": c["counter"].astype(np.int32),
               "solved": sel, "z_post": f32(c["z_post"]), "contact_phase": f32(c["contact_phase"]),
               "qp_args": f32(c["qp_args"][sel]), "meta": c["meta"],
               "torques_cold": f32(c["torques"]), "qp_out_cold": f32(c["qp_out"][sel]),
               "torques_warm": f32(w["torques"]), "qp_out_warm": f32(w["qp_out"][sel])}
        np.savez_compressed(os.path.join(HERE, "tick_golden_config1_aliengo_trot_1000.npz"), **out)
        print("config1: 1000 ticks, 500 QP solves, cold + warm outputs")
        sys.exit(0)
    if "--warm" in sys.argv:   # the reference's persistent OSQP workspace (warm start every tick)
        run_case("aliengo_trot_warm", RobotType.ALIENGO, GaitType.TROT, False, 60, 1, True)
        sys.exit(0)
    run_case("aliengo_trot", RobotType.ALIENGO, GaitType.TROT, False, 60, 1, True)
    run_case("aliengo_walk_flat", RobotType.ALIENGO, GaitType.WALK, True, 60, 2, False)
    run_case("a1_bound", RobotType.A1, GaitType.BOUND, False, 60, 3, True)
    run_case("go1_trot_flat", RobotType.GO1, GaitType.TROT, True, 60, 4, False)
