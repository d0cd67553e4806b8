"""mpc_collisionavoidance_b200 -- batched NMPC solve engine for NVIDIA B200 (sm_100a).

Drop-in for the `AcadosOcpSolver.{set, get, solve}` surface used by the nmpc_ca collision-avoidance scripts of
ivanacollg/MPC_CollisionAvoidance, for thousands of independent instances at once.  CUDA-only: importing the
description classes works anywhere; constructing a solver needs libusvmpc.so and a CUDA device.
"""
from .ocp import AcadosModel, AcadosOcp, AcadosOcpConstraints, AcadosOcpCost, AcadosOcpDims, AcadosOcpOptions  # noqa: F401


def __getattr__(name):
    if name in ("BatchedAcadosOcpSolver", "AcadosOcpSolver"):
        from .solver import BatchedAcadosOcpSolver
        return BatchedAcadosOcpSolver
    raise AttributeError(name)
