"""CPU tests of the host side: record layout vs the C header, gait tables vs the reference's
own Gait.py (golden fixture), the C-ABI library's exports, and loud failure without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from rl_mpc_locomotion_b200 import _lib, gait, records, synth
from rl_mpc_locomotion_b200.robots import ROBOTS, RobotType, params_table

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "mpc_b200.h")
GOLD = os.path.join(ROOT, "tests", "golden")


def test_record_offsets_match_header():
    txt = open(HEADER).read()
    for name, val in re.findall(r"#define\s+MPC_REC_(\w+)\s+(\d+)", txt):
        assert getattr(records, f"REC_{name}") == int(val), name
    assert records.rec_stride(10) == 96 and records.rec_stride(16) == 120
    assert int(re.search(r"#define\s+MPC_ROBOT_PARAMS_STRIDE\s+(\d+)", txt).group(1)) == records.ROBOT_PARAMS_STRIDE


def test_pack_unpack_roundtrip():
    rec, rt, params, _ = synth.make_batch(5, 10, seed=9)
    kw = records.unpack_record(rec[2], 10)
    table = kw.pop("foot_contact_states")
    again = records.pack_records(10, foot_contact_states=table, **kw)
    assert np.array_equal(again[0], rec[2])


def test_gait_tables_match_reference_gait_py():
    g = np.load(os.path.join(GOLD, "gait_golden.npz"))
    for gid in gait.GAITS:
        for tick in range(40):
            it, phase = gait.gait_iteration(tick, 2)
            assert np.array_equal(gait.mpc_table(gid, it, 10), g[f"table_{gid}"][tick]), (gid, tick)
            assert np.allclose(gait.contact_state(gid, phase), g[f"contact_{gid}"][tick], atol=1e-6)
            assert np.allclose(gait.swing_state(gid, phase), g[f"swing_{gid}"][tick], atol=1e-6)
    # horizon 16 cycles the 10-segment gait
    t16 = gait.mpc_table(0, 0.5, 16)
    assert np.array_equal(t16[:40], gait.mpc_table(0, 0.5, 10)) and np.array_equal(t16[40:64], t16[0:24])


def test_robot_table_matches_reference_constants():
    a = ROBOTS[RobotType.ALIENGO]
    assert abs(a.mass - 18.082) < 1e-12 and a.body_height == 0.35 and a.friction == 0.4
    assert params_table().shape == (3, 12)
    assert np.allclose(ROBOTS[RobotType.A1].inertia_diag, (0.17, 0.57, 0.64))
    assert np.allclose(ROBOTS[RobotType.GO1].hip_location(3), (-0.1881, -0.04675, 0))


def test_library_exports_every_declared_symbol():
    txt = open(HEADER).read()
    declared = set(re.findall(r"\b(mpc_b200_\w+)\s*\(", txt))
    assert declared == set(_lib.EXPORTS)
    lib = _lib.lib()
    for sym in declared:
        assert hasattr(lib, sym), sym
    assert b"sm_100a" in lib.mpc_b200_version()
    s = _lib.default_settings()
    assert (s.horizon, s.max_iter, s.check_every, s.scaling_iters, s.polish) == (10, 4000, 25, 10, 1)
    assert (s.rho, s.sigma, s.relax, s.eps_abs, s.delta) == (0.1, 1e-6, 1.6, 1e-3, 1e-6)
    info = _lib.kernel_info(10)
    assert info["smem_bytes"] < 227 * 1024 and info["robots_per_cta"] == 1


def test_no_cpu_fallback():
    """Without a CUDA device the product must fail loudly (never route through the oracle)."""
    if _lib.lib().mpc_b200_device_count() > 0:
        pytest.skip("GPU present")
    from rl_mpc_locomotion_b200.batched import solve_host
    rec, rt, params, _ = synth.make_batch(2, 10, seed=0)
    with pytest.raises(_lib.MpcB200Error, match="no CUDA device"):
        solve_host(rec, 10, robot_params=params, robot_type=rt)
    with pytest.raises(_lib.MpcB200Error):
        _lib.check(_lib.lib().mpc_b200_solve_batch(C.byref(_lib.default_settings(horizon=12)), None, None, None,
                                                   1, 0, None, None, None, None, None, None, None, None), "bad horizon")


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "rl_mpc_locomotion_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert "oracle" not in src.replace("the oracle", "").replace("CPU oracle", "") or f in ("synth.py",), f


def test_host_entry_rejects_bad_robot_type_before_touching_the_gpu():
    """Argument validation of the C-ABI happens on the host, before the device probe: a robot_type
    outside [0, n_types) or n_types < 1 is MPC_B200_EINVAL (device-pointer entry points flag the
    robot with MPC_STATUS_BAD_INDEX instead, tests/test_gpu_parity.py)."""
    from rl_mpc_locomotion_b200.batched import solve_host
    rec, rt, params, _ = synth.make_batch(3, 10, seed=0)
    bad = rt.copy()
    bad[1] = 7
    with pytest.raises(_lib.MpcB200Error, match=r"robot_type\[1\] = 7 outside"):
        solve_host(rec, 10, robot_params=params, robot_type=bad)
    tp = _lib.tick_params(cmd_stride=5)
    s = _lib.default_settings()
    one = C.c_void_p(8)   # never dereferenced: validation fails first
    rc = _lib.lib().mpc_b200_tick(C.byref(s), C.byref(tp), one, one, None, one, 0, None, None, one, one, None, 3, 4,
                                  None, one, one, None, None, None, None, None, None)
    assert rc == -1 and b"cmd_stride" in _lib.lib().mpc_b200_last_error()


def test_no_reference_gait_can_overflow_the_tick_kernels_polish_buffer():
    """The fused tick has no polish-overflow pass (include/mpc_b200.h, MPC_STATUS_POLISH_OVERFLOW):
    its polish buffer holds 6h free coordinates = 2h stance leg-steps with all three force
    components free.  Every gait of the reference (ConvexMPCLocomotion.py:30-56), at every solve
    phase, for both built horizons, must stay within that."""
    for h in (10, 16):
        for gid in gait.GAITS:
            for k in range(10):
                table = gait.mpc_table(gid, k + 0.5, h)
                assert table.sum() <= 2 * h, (h, gid, k, table.sum())


def test_integration_doc_ctypes_stub_matches_the_abi():
    """INTEGRATION.md §2 shows the ctypes binding a maintainer would write: its settings structure must be the
    one the package binds (same fields, order and types) and its call must pass every argument of the header."""
    import ctypes as C
    import re
    from rl_mpc_locomotion_b200 import _lib
    doc = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "INTEGRATION.md")).read()
    block = re.search(r"class MpcB200Settings\(C\.Structure\):.*?\n\n", doc, re.S).group(0)
    ns = {"C": C}
    exec(block, ns)
    got = [(n, t) for n, t in ns["MpcB200Settings"]._fields_]
    want = [(n, t) for n, t in _lib.Settings._fields_]
    assert [n for n, _ in got] == [n for n, _ in want]
    assert all(C.sizeof(a) == C.sizeof(b) for (_, a), (_, b) in zip(got, want))
    assert C.sizeof(ns["MpcB200Settings"]) == C.sizeof(_lib.Settings)
    call = re.search(r"lib\.mpc_b200_solve_batch\((.*?)\)\nassert", doc, re.S).group(1)
    call = re.sub(r"#[^\n]*", "", call)                           # drop the trailing comments
    depth, n_args = 0, 1
    for ch in call:
        depth += ch in "([" 
        depth -= ch in ")]"
        n_args += (ch == "," and depth == 0)
    assert n_args == 14                                            # include/mpc_b200.h: mpc_b200_solve_batch
