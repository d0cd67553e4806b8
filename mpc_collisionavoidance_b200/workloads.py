"""Synthetic NMPC batches for the USV collision-avoidance benchmark.

Seeded and numpy-only so that tests, bench.py and the CPU baseline all see identical inputs.
The ranges are the survey's; nothing here touches the GPU.
"""
from dataclasses import dataclass

import numpy as np

U_REF = 0.9
DT = 0.05

# Position reference: X_ref = X_0 + XREF_FACTOR * u_ref * dt * N.  The survey's 1.5 asks for more distance than the
# horizon can cover at u_ref; at N = 100 that drives the surge speed across the model's non-smooth drag switch
# (u > 1.25, usv_model.py: if_else) and the reference's full-step SQP then cycles on every scene (0/64 converged).  With
# 0.8 the reference converges on 81 % of the N = 100 scenes (52/64, same seed); SURVEY.md section 8(d) lets the builder
# tune the ranges and asks for them to be recorded: this is the only range that differs from the survey's, config 4 only.
XREF_FACTOR = {4: 0.8}

# BASELINE.json configs: (N, K, batch, num_steps)
CONFIGS = {
    1: dict(N=20, K=3, B=1, num_steps=1),
    2: dict(N=40, K=5, B=4096, num_steps=4),
    3: dict(N=40, K=20, B=4096, num_steps=4),
    4: dict(N=100, K=5, B=16384, num_steps=4),
    5: dict(N=40, K=5, B=131072, num_steps=4),
}


@dataclass
class Batch:
    x0: np.ndarray      # [B, 6]
    p: np.ndarray       # [B, 2K]   obstacle centres (ox_1, oy_1, ...), constant over the horizon
    lh: np.ndarray      # [B, K]    obstacle radii = lower bound of the distance constraint
    yref: np.ndarray    # [B, 8]
    yref_e: np.ndarray  # [B, 6]


def make_batch(config_id: int, B: int = None, N: int = None, K: int = None, seed: int = None) -> Batch:
    cfg = CONFIGS[config_id]
    B = cfg["B"] if B is None else B
    N = cfg["N"] if N is None else N
    K = cfg["K"] if K is None else K
    rng = np.random.default_rng(1234 + config_id if seed is None else seed)
    if config_id == 5:
        # Monte-Carlo disturbance scenarios: 128 base scenes x draws of x0 += N(0, diag(...)^2)
        nbase = min(128, B)
        base = _scenes(rng, nbase, N, K, XREF_FACTOR.get(config_id, 1.5))
        reps = -(-B // nbase)
        idx = np.tile(np.arange(nbase), reps)[:B]
        sig = np.array([0.05, 0.05, 0.02, 0.05, 0.01, 0.02])
        x0 = base.x0[idx] + rng.standard_normal((B, 6)) * sig
        return Batch(x0, base.p[idx].copy(), base.lh[idx].copy(), base.yref[idx].copy(), base.yref_e[idx].copy())
    return _scenes(rng, B, N, K, XREF_FACTOR.get(config_id, 1.5))


def _scenes(rng, B, N, K, xref_factor=1.5) -> Batch:
    x0 = np.stack([np.zeros(B), rng.uniform(-1, 1, B), rng.uniform(-0.3, 0.3, B), rng.uniform(0.4, 1.0, B),
                   rng.uniform(-0.02, 0.02, B), rng.uniform(-0.05, 0.05, B)], axis=1)
    span = U_REF * DT * N
    yref = np.zeros((B, 8))
    yref[:, 0] = x0[:, 0] + span * xref_factor
    yref[:, 3] = U_REF
    p = np.zeros((B, 2 * K))
    lh = np.zeros((B, K))
    for k in range(K):
        todo = np.ones(B, dtype=bool)
        while todo.any():
            n = int(todo.sum())
            ox = rng.uniform(1.0, 1.0 + span * 1.2, n)
            oy = rng.uniform(-2.5, 2.5, n)
            r = rng.uniform(0.3, 0.9, n)
            ok = np.hypot(ox - x0[todo, 0], oy - x0[todo, 1]) >= r + 0.3
            ids = np.flatnonzero(todo)[ok]
            p[ids, 2 * k], p[ids, 2 * k + 1], lh[ids, k] = ox[ok], oy[ok], r[ok]
            todo[ids] = False
    return Batch(x0, p, lh, yref, yref[:, :6].copy())


@dataclass
class GuidanceBatch:
    x0: np.ndarray      # [B, 8]   u, v, ye, chie, psied, xned, yned, psi
    p: np.ndarray       # [B, 16]  obstacle centres in NED
    lh: np.ndarray      # [B, 8]   obstacle radii (lower bound of the distance rows)
    yref: np.ndarray    # [B, 9]
    yref_e: np.ndarray  # [B, 8]


def make_guidance_batch(B: int, seed: int = 77) -> GuidanceBatch:
    """Synthetic scenes for the deployed collision-avoidance OCP (usv_guidance_ca1): the vessel follows the path
    (4,-5) -> (4,25) of the script (main.py:73-75,101-104) heading north; the obstacles of the script's list, jittered,
    the unused slots parked at (100,100) like the node does (nmpc_guidance_ca1.cpp:330-363)."""
    rng = np.random.default_rng(seed)
    x0 = np.stack([rng.uniform(0.4, 1.0, B), rng.uniform(-0.05, 0.05, B), rng.uniform(-0.5, 0.5, B), rng.uniform(-0.3, 0.3, B),
                   rng.uniform(-0.3, 0.3, B), 4.0 + rng.uniform(-0.5, 0.5, B), rng.uniform(-6.0, 6.0, B),
                   np.pi / 2 + rng.uniform(-0.3, 0.3, B)], axis=1)
    base = np.array([[4, 4], [4, 7], [4, 12], [4, 20]], dtype=float)
    p = np.full((B, 8, 2), 100.0)
    p[:, :4] = base[None] + rng.uniform(-1.0, 1.0, (B, 4, 2))
    lh = np.full((B, 8), 1.5)
    lh[:, :4] = rng.uniform(1.0, 2.0, (B, 4))
    return GuidanceBatch(x0, p.reshape(B, 16), lh, np.zeros((B, 9)), np.zeros((B, 8)))


def guidance_ca1_ocp(N: int = 100, Tf: float = 5.0, nlp_solver_type: str = "SQP_RTI", soft: bool = True):
    """The OCP of the deployed CA node as an AcadosOcp-style description, numbers of
    nmpc_ca/scripts/usv_guidance_ca1/acados_settings.py:70-208 (soft = False: the same rows as hard constraints)."""
    from .ocp import AcadosOcp
    nx, nu, K = 8, 1, 8
    ocp = AcadosOcp()
    ocp.model.name = "usv_model_guidance_ca1"
    ocp.dims.N = N
    Q = np.diag([0, 0, 0.05, 0.01, 0, 0, 0, 0.0])
    ocp.cost.W = np.block([[Q, np.zeros((nx, nu))], [np.zeros((nu, nx)), np.array([[0.2]])]])
    ocp.cost.W_e = np.diag([0, 0, 0.1, 0.05, 0, 0, 0, 0.0])
    Vx = np.zeros((nx + nu, nx)); Vx[:nx, :nx] = np.eye(nx)
    Vu = np.zeros((nx + nu, nu)); Vu[nx:, :] = np.eye(nu)
    ocp.cost.Vx, ocp.cost.Vu, ocp.cost.Vx_e = Vx, Vu, np.eye(nx)
    ocp.cost.yref, ocp.cost.yref_e = np.zeros(nx + nu), np.zeros(nx)
    ocp.constraints.lbu, ocp.constraints.ubu, ocp.constraints.idxbu = np.array([-0.5]), np.array([0.5]), np.array([0])
    ocp.constraints.lh, ocp.constraints.uh = np.full(K, 1.5), np.full(K, 1e6)
    if soft:
        ocp.cost.zl, ocp.cost.zu, ocp.cost.Zl, ocp.cost.Zu = np.ones(K), np.ones(K), np.zeros(K), np.zeros(K)
        ocp.constraints.lsh, ocp.constraints.ush, ocp.constraints.idxsh = np.full(K, -0.2), np.zeros(K), np.arange(K)
    ocp.constraints.x0 = np.zeros(nx)
    ocp.parameter_values = np.full(2 * K, 100.0)
    o = ocp.solver_options
    o.tf = Tf
    o.qp_solver, o.hessian_approx, o.integrator_type = "PARTIAL_CONDENSING_HPIPM", "GAUSS_NEWTON", "ERK"
    o.nlp_solver_type = nlp_solver_type
    o.sim_method_num_stages, o.sim_method_num_steps = 4, 1
    return ocp


def benchmark_ocp(config_id: int, nlp_solver_type: str = "SQP"):
    """The benchmark OCP of SURVEY.md section 8d as an AcadosOcp-style description (same attribute names as the
    nmpc_ca `acados_settings.py` scripts): 3-DOF USV, LINEAR_LS tracking cost, thrust and velocity boxes, K obstacle
    distance rows with uh = 1e6, ERK4, Gauss-Newton, partial-condensing HPIPM with cond_N = N."""
    from .ocp import AcadosOcp
    cfg = CONFIGS[config_id]
    N, K = cfg["N"], cfg["K"]
    nx, nu = 6, 2
    ocp = AcadosOcp()
    ocp.model.name = "usv3"
    ocp.dims.N = N
    Q = np.diag([1, 1, 0.1, 10, 0.1, 0.1])
    R = np.diag([1e-3, 1e-3])
    ocp.cost.W = np.block([[Q, np.zeros((nx, nu))], [np.zeros((nu, nx)), R]])
    ocp.cost.W_e = 5 * Q
    Vx = np.zeros((nx + nu, nx)); Vx[:nx, :nx] = np.eye(nx)
    Vu = np.zeros((nx + nu, nu)); Vu[nx:, :] = np.eye(nu)
    ocp.cost.Vx, ocp.cost.Vu, ocp.cost.Vx_e = Vx, Vu, np.eye(nx)
    ocp.cost.yref, ocp.cost.yref_e = np.zeros(nx + nu), np.zeros(nx)
    ocp.constraints.lbu, ocp.constraints.ubu = np.array([-30.0, -30.0]), np.array([35.0, 35.0])
    ocp.constraints.idxbu = np.array([0, 1])
    ocp.constraints.lbx, ocp.constraints.ubx = np.array([-1.5, -1.5, -1.0]), np.array([1.5, 1.5, 1.0])
    ocp.constraints.idxbx = np.array([3, 4, 5])
    ocp.constraints.lh, ocp.constraints.uh = np.full(K, 0.5), np.full(K, 1e6)
    ocp.constraints.x0 = np.array([0, 0, 0, 0.7, 0, 0.0])
    ocp.parameter_values = np.zeros(2 * K)
    o = ocp.solver_options
    o.tf = DT * N
    o.qp_solver, o.hessian_approx, o.integrator_type = "PARTIAL_CONDENSING_HPIPM", "GAUSS_NEWTON", "ERK"
    o.nlp_solver_type = nlp_solver_type
    o.sim_method_num_stages, o.sim_method_num_steps = 4, cfg["num_steps"]
    o.nlp_solver_max_iter, o.qp_solver_iter_max = 100, 50
    return ocp
