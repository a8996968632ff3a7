"""The oracle pinned to the reference's OWN source.

`oracle/_ref/refmod/mpc_osqp*.so` is the reference's MPC_Controller/convex_MPC/mpc_osqp.cc compiled UNMODIFIED
(oracle/Makefile), linked against the reference's unmodified OSQP, with `oracle/eigen_standin/` standing in
for the three git submodules the file includes and /root/reference does not contain (Eigen: the dense subset
it uses, restated; qpOASES: declarations only; pybind11/eigen.h: empty).  Its `ConvexMpc.compute_contact_forces`
is therefore the reference's QP builder, OSQP glue, warm-start path and result handling line for line - only
the dense products / 3x3 inverse / quaternion formulas / matrix exponential underneath are ours.

These tests hold the restatement (`oracle/mpc_oracle.cc`, the thing every GPU parity test is checked
against) and the committed golden vectors to that build.  CPU only; skipped where the module was not built
(it needs /root/reference at build time; on the GPU box the prebuilt file travels with oracle/_ref/)."""
import glob
import importlib.util
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="module")
def refmod():
    so = glob.glob(os.path.join(ROOT, "oracle", "_ref", "refmod", "mpc_osqp*.so"))
    if not so:
        pytest.skip("reference-source module not built (make -C oracle needs /root/reference)")
    spec = importlib.util.spec_from_file_location("mpc_osqp", so[0])
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    assert m.TEST == 42                    # mpc_osqp.cc:979
    return m


def _args(rec, h):
    """The 13 positional arguments of compute_contact_forces (mpc_osqp.cc:578-591) from a problem record."""
    from rl_mpc_locomotion_b200 import records as R
    f = lambda o, n: [float(x) for x in rec[o:o + n]]
    return (f(R.REC_WEIGHTS, 13), f(R.REC_COM_POS, 3), f(R.REC_COM_VEL, 3), f(R.REC_RPY, 3), f(R.REC_NORMAL, 3),
            f(R.REC_ANGVEL, 3), f(R.REC_CONTACT, 4 * h), f(R.REC_FOOTPOS, 12), f(R.REC_MU, 4), f(R.REC_DES_POS, 3),
            f(R.REC_DES_VEL, 3), f(R.REC_DES_RPY, 3), f(R.REC_DES_ANGVEL, 3))


def _new(refmod, rp, h):
    return refmod.ConvexMpc(float(rp[0]), [float(x) for x in rp[1:10]], 4, h, 0.02, 1e-5, refmod.QPSolverName.OSQP)


@pytest.mark.parametrize("h,n,gaits,robots,seed", [(10, 96, (0,), (0,), 0), (10, 96, (0, 6, 1), (0,), 2),
                                                    (16, 48, (0, 6, 1), (0, 1, 2), 4)])
def test_restatement_matches_the_reference_source_build(refmod, oracle_mod, h, n, gaits, robots, seed):
    """Fresh `ConvexMpc` per QP (osqp_setup path, mpc_osqp.cc:757-759) on the seeded bench distributions:
    full-horizon forces of the restatement vs the reference's own file."""
    from rl_mpc_locomotion_b200 import synth
    rec, rt, params, _ = synth.make_batch(n, h, seed=seed, gaits=gaits, robot_types=robots)
    _, want, st = oracle_mod.solve_records(rec, h, params, rt, mode=oracle_mod.COLD, full=True)
    got = np.stack([np.asarray(_new(refmod, params[rt[i]], h).compute_contact_forces(*_args(rec[i], h))) for i in range(n)])
    assert (st == 1).all() and got.shape == want.shape
    err = np.abs(got - want).max(1)
    print(f"h={h}: |restatement - reference source| max {err.max():.3g} N, median {np.median(err):.3g} N")
    assert err.max() < 1e-6      # measured 3e-8: different association of the dense sums, same OSQP trajectory


@pytest.mark.parametrize("h", [10, 16])
def test_golden_vectors_match_the_reference_source_build(refmod, h):
    """The committed golden forces (what the warp emulation and the GPU kernel are held to) are what the
    reference's own file returns on the golden inputs."""
    g = np.load(os.path.join(GOLD, f"qp_golden_h{h}.npz"))
    rec, rt, params = g["records"], g["robot_type"], g["robot_params"]
    got = np.stack([np.asarray(_new(refmod, params[rt[i]], h).compute_contact_forces(*_args(rec[i], h)))
                    for i in range(len(rec))])
    err = np.abs(got - g["forces"]).max(1)
    print(f"h={h}: |golden - reference source| max {err.max():.3g} N, 99th percentile {np.percentile(err, 99):.3g} N")
    # 1e-8 N except on the rare QPs whose polished KKT system is nearly singular: there the reference's own
    # answer moves by 1e-5..5e-4 N with the rounding of its inputs (test_oracle.py::test_reference_self_noise_*)
    assert err.max() < 2e-4 and np.percentile(err, 90) < 1e-6


@pytest.mark.parametrize("name,rtype", [("aliengo_trot", 0), ("a1_bound", 1), ("aliengo_walk_flat", 0), ("go1_trot_flat", 2)])
def test_warm_workspace_path_matches_the_reference_source_build(refmod, oracle_mod, name, rtype):
    """ONE persistent `ConvexMpc` over 30 consecutive MPC solves of a replayed gait: the reference's
    osqp_update_P_A / update_lin_cost / update_bounds / warm-started osqp_solve path (mpc_osqp.cc:760-780)
    vs the restatement's WARM mode, solve for solve."""
    from rl_mpc_locomotion_b200.records import unpack_record
    from rl_mpc_locomotion_b200.robots import params_table
    g = np.load(os.path.join(GOLD, f"tick_golden_{name}.npz"))
    a = g["qp_args"][np.nonzero(g["solved"])[0]]
    rec = np.concatenate([a[:, :28], a[:, 68:96], a[:, 28:68]], 1).astype(np.float32)
    rp = params_table()[rtype].astype(np.float64)
    ref, ora = _new(refmod, rp, 10), oracle_mod.OracleMpc(rp[0], rp[1:10], horizon=10)
    worst = 0.0
    for t in range(len(rec)):
        st, want = ora.solve(oracle_mod.WARM, **unpack_record(rec[t], 10))
        got = np.asarray(ref.compute_contact_forces(*_args(rec[t], 10)))
        assert st == 1 and got.size == want.size
        worst = max(worst, np.abs(got - want).max())
    assert worst < 1e-5, worst
    # reset_solver() only raises initial_run_ (mpc_osqp.cc:576): the workspace, hence the next result, is untouched
    ref.reset_solver()
    ora2 = oracle_mod.OracleMpc(rp[0], rp[1:10], horizon=10)
    for t in range(len(rec)):
        ora2.solve(oracle_mod.WARM, **unpack_record(rec[t], 10))
    oracle_mod.lib().mpc_oracle_reset_solver(ora2._o)
    st, want = ora2.solve(oracle_mod.WARM, **unpack_record(rec[0], 10))
    got = np.asarray(ref.compute_contact_forces(*_args(rec[0], 10)))
    assert np.abs(got - want).max() < 1e-5


def test_non_finite_input_comes_back_as_nan_from_the_reference(refmod):
    """What the reference really does with a NaN state: OSQP's inf-norms skip NaN (comparisons with NaN are
    false), the run ends OSQP_SOLVED and the caller gets 3*4*h NaN forces.  The B200 path deviates on
    purpose: it reports MPC_STATUS_NON_FINITE and returns zeros / an empty drop-in result."""
    from rl_mpc_locomotion_b200 import synth
    rec, rt, params, _ = synth.make_batch(1, 10, seed=0)
    r = rec[0].copy()
    r[16] = np.nan                           # com_velocity[0]
    out = np.asarray(_new(refmod, params[rt[0]], 10).compute_contact_forces(*_args(r, 10)))
    assert out.size == 120 and np.isnan(out).all()


def test_reference_python_and_reference_cpp_reproduce_the_tick_fixtures(refmod):
    """BASELINE config 1 end to end on the reference's own code: the UNMODIFIED reference Python
    (`RobotRunnerMin` wiring on a kinematic replay) calling the UNMODIFIED reference mpc_osqp.cc (OSQP branch,
    persistent workspace) reproduces the committed warm fixtures that the CUDA tick is tested against - 60 ticks
    of Aliengo/trot and the 1000-tick config-1 run.  Needs /root/reference (the generator imports it)."""
    import subprocess
    import sys
    if not os.path.isdir("/root/reference/MPC_Controller"):
        pytest.skip("needs /root/reference")
    out = subprocess.run([sys.executable, os.path.join(GOLD, "make_tick_golden.py"), "--refmod", "--check", "--config1"],
                         capture_output=True, text=True, timeout=600)
    print(out.stdout[-600:])
    assert out.returncode == 0, out.stderr[-2000:]
