"""Batched closed-loop simulation with the real-time-iteration scheme -- the loop of the reference's scripts
(nmpc_ca/scripts/usv_guidance_ca1/main.py:116-175): set x0, one SQP_RTI step warm-started from the previous
iterate, apply u_0, take the model prediction x_1 (plus an optional disturbance) as the next x0.  Everything stays
on the GPU between steps: the state feedback is a device-to-device copy through the solver's own get/set."""
import numpy as np


def simulate_closed_loop(solver, x0, n_steps, disturbance=None, on_step=None, split_phases=False):
    """Run `n_steps` control steps for all B instances of `solver` (a BatchedAcadosOcpSolver in SQP_RTI mode whose
    references, parameters and bounds are already set).  x0: [B, nx] numpy or torch.  disturbance(step, x_next) may
    return the perturbed next state (torch tensor on the solver's device).  Returns (X [B, n_steps+1, nx],
    U [B, n_steps, nu], status [B, n_steps]) as numpy arrays.

    split_phases: the real-time-iteration schedule (ocp_nlp_sqp_rti.c:459-488, options_set("rti_phase", 1 | 2)): the
    preparation phase (linearisation at the current iterate) runs BEFORE the measurement arrives, the feedback phase
    (QP with the new x0, variable update) after it -- only the feedback phase is on the measurement-to-control path.
    The trajectories are the same as with the single-call step (rti_phase 0)."""
    torch = solver._torch
    dev = f"cuda:{solver.device}"
    x = torch.as_tensor(np.asarray(x0) if not torch.is_tensor(x0) else x0, dtype=torch.float64, device=dev).contiguous()
    B, nx, nu = solver.B, solver.nx, solver.nu
    X = torch.empty((B, n_steps + 1, nx), dtype=torch.float64, device=dev)
    U = torch.empty((B, n_steps, nu), dtype=torch.float64, device=dev)
    S = torch.empty((B, n_steps), dtype=torch.float64, device=dev)
    X[:, 0] = x
    solver.options_set("cold_start", 0)     # warm start from the previous iterate, no shifting (as the scripts do)
    if split_phases:
        solver.options_set("rti_phase", 1)
        solver.solve_async()                # first preparation
    for i in range(n_steps):
        solver.set(0, "lbx", x)
        solver.set(0, "ubx", x)
        if split_phases:
            solver.options_set("rti_phase", 2)
        solver.solve_async()
        U[:, i] = solver.get(0, "u", device=True)
        x = solver.get(1, "x", device=True)
        S[:, i] = solver.stats_table(device=True)[:, 0]
        if disturbance is not None:
            x = disturbance(i, x)
        x = x.contiguous()
        X[:, i + 1] = x
        if on_step is not None:
            on_step(i, solver)
        if split_phases and i + 1 < n_steps:
            solver.options_set("rti_phase", 1)
            solver.solve_async()            # preparation for the next step while the plant moves
    if split_phases:
        solver.options_set("rti_phase", 0)
    torch.cuda.synchronize()
    return X.cpu().numpy(), U.cpu().numpy(), S.cpu().numpy().astype(np.int64)
