"""Offset/duration gaits (host mirror, vectorised over robots).

Mirrors reference/MPC_Controller/convex_MPC/Gait.py:3-93 (OffsetDurationGait) and the gait
set constructed in ConvexMPCLocomotion.__init__ (ConvexMPCLocomotion.py:30-56), including
the reference's *fractional* iteration (true division in ``setIterations``, Gait.py:27 —
SURVEY App. B Q4).  The device tick kernel evaluates the same formulas; this module is the
host-side table source (synthetic inputs, the ConvexMpc compat path, tests).
"""
from __future__ import annotations

import numpy as np

N_SEGMENT = 10
# gait id -> (offsets, durations) in MPC segments; ids follow ConvexMPCLocomotion.run
# (ConvexMPCLocomotion.py:227-241): 0 trot, 1 bound, 2 pronk, 3 pace, 5 gallop, 6 walk,
# 7 trot-run; any other id falls back to trot exactly as the reference does.
GAITS = {
    0: ((0, 5, 5, 0), (5, 5, 5, 5)),
    1: ((5, 5, 0, 0), (4, 4, 4, 4)),
    2: ((0, 0, 0, 0), (4, 4, 4, 4)),
    3: ((5, 0, 5, 0), (5, 5, 5, 5)),
    5: ((0, 2, 7, 9), (4, 4, 4, 4)),
    6: ((0, 3, 5, 8), (5, 5, 5, 5)),
    7: ((0, 5, 5, 0), (4, 4, 4, 4)),
}
GAIT_NAMES = {0: "trot", 1: "bound", 2: "pronk", 3: "pace", 5: "gallop", 6: "walk", 7: "trotrun"}


def gait_arrays(gait_id) -> tuple:
    """offsets[...,4], durations[...,4] (float64) for scalar or array gait ids."""
    gid = np.asarray(gait_id)
    off = np.empty(gid.shape + (4,))
    dur = np.empty(gid.shape + (4,))
    off[...] = GAITS[0][0]
    dur[...] = GAITS[0][1]
    for k, (o, d) in GAITS.items():
        m = gid == k
        off[m] = o
        dur[m] = d
    return off, dur


def gait_iteration(iteration_counter, iterations_per_mpc: int):
    """``__iteration`` and ``__phase`` of Gait.setIterations (Gait.py:26-28)."""
    c = np.asarray(iteration_counter, dtype=np.float64)
    it = (c / iterations_per_mpc) % N_SEGMENT
    span = iterations_per_mpc * N_SEGMENT
    phase = np.mod(np.asarray(iteration_counter), span).astype(np.float64) / float(span)
    return it, phase


def mpc_table(gait_id, iteration, horizon: int = 10) -> np.ndarray:
    """Horizon contact table, Gait.getMpcTable (Gait.py:69-84).

    ``iteration`` is the (possibly fractional) ``__iteration``.  Returns
    ``[..., 4*horizon]`` float32 with index ``i*4+leg``.  For horizon > 10 the 10-segment
    gait is cycled (``(i+iteration+1) % 10`` for ``i < horizon`` — SURVEY §7.7).
    """
    off, dur = gait_arrays(gait_id)
    it = np.asarray(iteration, dtype=np.float64)
    i = np.arange(horizon, dtype=np.float64)
    cur = np.mod(i + it[..., None] + 1.0, N_SEGMENT)             # [..., h]
    progress = cur[..., :, None] - off[..., None, :]             # [..., h, 4]
    progress = np.where(progress < 0, progress + N_SEGMENT, progress)
    table = (progress < dur[..., None, :]).astype(np.float32)
    return table.reshape(table.shape[:-2] + (4 * horizon,))


def _f32_phase_arrays(gait_id, phase):
    off, dur = gait_arrays(gait_id)
    # the reference keeps offsets/durations as float32 arrays (ConvexMPCLocomotion.py:30-56) and
    # mixes them with the Python-float phase, so every operation below happens in float32
    offf = off.astype(np.float32) / np.float32(N_SEGMENT)
    durf = dur.astype(np.float32) / np.float32(N_SEGMENT)
    ph = np.asarray(phase, dtype=np.float64).astype(np.float32)[..., None]
    return offf, durf, ph


def contact_state(gait_id, phase) -> np.ndarray:
    """Gait.getContactState (Gait.py:30-44): stance progress in [0,1], 0 while in swing.
    float32 arithmetic in the reference's operation order (phase boundaries depend on it)."""
    offf, durf, ph = _f32_phase_arrays(gait_id, phase)
    p = ph - offf
    p = np.where(p < 0, p + np.float32(1.0), p)
    return np.where(p > durf, np.float32(0.0), p / durf)


def swing_state(gait_id, phase) -> np.ndarray:
    """Gait.getSwingState (Gait.py:46-67): swing progress in [0,1], 0 while in stance."""
    offf, durf, ph = _f32_phase_arrays(gait_id, phase)
    sw_off = offf + durf
    sw_off = np.where(sw_off > 1, sw_off - np.float32(1.0), sw_off)
    sw_dur = np.float32(1.0) - durf
    p = ph - sw_off
    p = np.where(p < 0, p + np.float32(1.0), p)
    safe = np.where(sw_dur == 0, np.float32(1.0), sw_dur)
    return np.where((p > sw_dur) | (sw_dur == 0), np.float32(0.0), p / safe)
