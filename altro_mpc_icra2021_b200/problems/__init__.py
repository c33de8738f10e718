"""Benchmark problem families of the reference (random linear, rocket landing, quadruped, grasp,
flexible satellite) as batched Problem builders, plus the MPC shift-loop helpers."""
from . import flexsat, grasp, mpc, quadruped, random_linear, rocket  # noqa: F401
