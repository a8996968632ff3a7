"""Kernel LOGIC on the CPU: the very warp program the sm_100a kernel compiles, built with
-DMPC_EMULATE (tests/warp_emu, 32 lanes executed sequentially), against the oracle (real
reference OSQP).  The GPU parity tests (-m gpu) repeat this on the device."""
import os

import numpy as np
import pytest

import emu
from rl_mpc_locomotion_b200 import synth
from rl_mpc_locomotion_b200.robots import RobotType

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("h", [10, 16])
def test_emulated_kernel_matches_golden_reference_output(h):
    g = np.load(os.path.join(GOLD, f"qp_golden_h{h}.npz"))
    out = emu.solve(g["records"], h, g["robot_params"], g["robot_type"])
    assert np.array_equal(out["status"], g["status"])
    assert np.array_equal(out["info"][:, 0].astype(int), g["iters"])
    assert np.array_equal(out["info"][:, 3].astype(int), g["polish"])
    err = np.abs(out["grf_full"] - g["forces"]).max()
    assert err < 1e-3, err


def test_emulated_builder_and_scaling(oracle_mod):
    from oracle import np_osqp
    from rl_mpc_locomotion_b200.records import unpack_record
    rec, rt, params, _ = synth.make_batch(3, 10, seed=12, robot_types=tuple(RobotType))
    out = emu.solve(rec, 10, params, rt, want_dbg=True)
    for r in range(3):
        rp = params[rt[r]].astype(np.float64)
        o = oracle_mod.OracleMpc(rp[0], rp[1:10], horizon=10)
        qp = o.build(**unpack_record(rec[r], 10))
        assert np.abs(out["P"][r] - qp["P"]).max() < 1e-12 * np.abs(qp["P"]).max()
        assert np.abs(out["q"][r] - qp["q"]).max() < 1e-12 * np.abs(qp["q"]).max()
        _, _, _, sc = np_osqp.scale_data(qp["P"], qp["q"], qp["C"])
        assert np.allclose(out["D"][r], sc.D, rtol=1e-11) and np.allclose(out["E"][r], sc.E, rtol=1e-11)
        assert abs(out["c"][r] / sc.c - 1) < 1e-11


def test_emulated_kernel_vs_oracle_mixed_batch(oracle_mod):
    rec, rt, params, _ = synth.make_batch(160, 10, seed=21, gaits=(0, 6, 1), robot_types=tuple(RobotType))
    out = emu.solve(rec, 10, params, rt)
    f0, ff, st = oracle_mod.solve_records(rec, 10, params, rt, mode=oracle_mod.COLD, full=True)
    assert np.array_equal(out["status"], st)
    per = np.abs(out["grf_full"] - ff).max(axis=1)
    assert np.abs(out["grf0"] - f0).max() < 1e-3
    assert np.quantile(per, 0.99) < 1e-3 and per.max() < 1e-2


def test_fixed_iteration_mode_converges_to_true_minimiser(oracle_mod):
    """MPC_MODE_FIXED (BASELINE config 5): more iterations -> closer to the true minimiser."""
    g = np.load(os.path.join(GOLD, "qp_golden_h10.npz"))
    rec, rt, params = g["records"][:24], g["robot_type"][:24], g["robot_params"]
    errs = []
    for iters in (25, 100, 400):
        prm = emu.default_params(fixed_iters=iters, adaptive_rho=0, polish=0)
        out = emu.solve(rec, 10, params, rt, params=prm)
        errs.append(np.abs(out["grf_full"] - g["true_minimiser"][:24]).max(axis=1).mean())
    assert errs[2] < errs[1] < errs[0]
