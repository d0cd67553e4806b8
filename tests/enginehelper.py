"""Drive the CUDA engine through its public Python surface (-> C ABI) from the same problem description the
oracle and the reference harness use (refharness.RefProblem)."""
import numpy as np

from mpc_collisionavoidance_b200 import AcadosOcp
from refharness import RefProblem


def ocp_from_problem(P: RefProblem) -> AcadosOcp:
    ocp = AcadosOcp()
    ocp.model.name = "pendulum" if P.model == 1 else "usv3"
    ocp.dims.N = P.N
    ocp.cost.W, ocp.cost.W_e = np.array(P.W), np.array(P.We)
    ocp.constraints.lbu, ocp.constraints.ubu = P.lbu, P.ubu
    ocp.constraints.idxbu = np.arange(len(P.lbu))
    ocp.constraints.lbx, ocp.constraints.ubx, ocp.constraints.idxbx = P.lbx, P.ubx, P.idxbx
    ocp.constraints.lh = np.zeros(P.K)
    ocp.constraints.uh = np.full(P.K, P.uh)
    ocp.parameter_values = np.zeros(2 * P.K)
    o = ocp.solver_options
    o.tf = P.dt * P.N
    o.nlp_solver_type = "SQP" if P.icfg[5] == 0 else "SQP_RTI"
    o.sim_method_num_steps, o.sim_method_num_stages = int(P.icfg[3]), int(P.icfg[4])
    o.nlp_solver_max_iter, o.qp_solver_iter_max = int(P.icfg[6]), int(P.icfg[7])
    o.nlp_solver_tol_stat, o.nlp_solver_tol_eq, o.nlp_solver_tol_ineq, o.nlp_solver_tol_comp = [float(v) for v in P.dcfg[1:5]]
    return ocp


def engine_solve(P: RefProblem, x0, p, lh, yref, yref_e, xinit=None, uinit=None, piinit=None, per_stage_calls=False,
                 solver=None):
    """Same inputs/outputs as oracleport.solve_batch, computed on cuda:0 by the engine."""
    from mpc_collisionavoidance_b200 import BatchedAcadosOcpSolver
    x0 = np.atleast_2d(np.asarray(x0, dtype=np.float64))
    B, N, K = x0.shape[0], P.N, P.K
    s = solver or BatchedAcadosOcpSolver(ocp_from_problem(P), batch=B)
    s.set(0, "lbx", x0)
    s.set(0, "ubx", x0)
    yref = np.asarray(yref, dtype=np.float64)
    if per_stage_calls:
        # exactly the call sequence of the reference scripts (usv_guidance_ca1/main.py:116-131): 3N+4 setter calls
        for j in range(N):
            s.set(j, "yref", yref if yref.ndim == 2 else yref[:, j])
            if K:
                s.set(j, "p", p if np.ndim(p) == 2 else p[:, j])
                s.constraints_set(j, "lh", lh if np.ndim(lh) == 2 else lh[:, j])
        if K:
            s.set(N, "p", p if np.ndim(p) == 2 else p[:, N])
    else:
        s.set("every" if yref.ndim == 2 else "all", "yref", yref)
        if K:
            s.set("every" if np.ndim(p) == 2 else "all", "p", np.asarray(p, dtype=np.float64))
            s.constraints_set("every" if np.ndim(lh) == 2 else "all", "lh", np.asarray(lh, dtype=np.float64))
    s.set(N, "yref", np.asarray(yref_e, dtype=np.float64))
    if xinit is None:
        s.options_set("cold_start", 1)
    else:
        s.options_set("cold_start", 0)
        s.set("all", "x", np.asarray(xinit, dtype=np.float64))
        s.set("all", "u", np.asarray(uinit, dtype=np.float64))
        s.set("all", "pi", np.asarray(piinit, dtype=np.float64))
        for k in range(N + 1):
            for f in ("lam", "t"):
                d = s._dims(k, f)
                if d:
                    s.set(k, f, np.zeros((B, d)))
    status = s.solve()
    st = s.stats_table()
    out = dict(x=s.get_all("x"), u=s.get_all("u"), pi=s.get_all("pi"), status=np.asarray(status), sqp_iter=st[:, 1].astype(int),
               qp_iter=st[:, 2].astype(int), res=st[:, 3:7], solve_calls=st[:, 8].astype(int), solver=s)
    out["lam"] = [s.get(k, "lam") for k in range(N + 1)]
    out["t"] = [s.get(k, "t") for k in range(N + 1)]
    return out
