"""rl_mpc_locomotion_b200 — B200-native batched convex-MPC stance-force solver.

Drop-in for the native QP path of silvery107/rl-mpc-locomotion (`mpc_osqp`), see DESIGN.md.
"""
__version__ = "0.1.0"
