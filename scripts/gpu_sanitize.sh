#!/bin/bash
# compute-sanitizer on a small batch: memcheck, racecheck (shared-memory hazards between lanes), synccheck
set -x
mkdir -p gpurun_out
cat > /tmp/san.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np
import refharness as rh, oracleport as op
from enginehelper import engine_solve
from mpc_collisionavoidance_b200.workloads import make_batch
b = make_batch(2, B=6, seed=5)
P = rh.RefProblem(N=40, K=5, num_steps=4, max_iter=12)
r = engine_solve(P, b.x0, b.p, b.lh, b.yref, b.yref_e)
a = op.solve_batch(P, b.x0, b.p, b.lh, b.yref, b.yref_e, nthreads=4)
print("status", r["status"], a["status"], "max|dx|", np.abs(r["x"] - a["x"]).max())
if os.environ.get("SAN_MORE"):
    # the fp32 factorisation (two warps, named barrier) and the soft-constraint instantiation of the kernel
    from mpc_collisionavoidance_b200 import BatchedAcadosOcpSolver
    from enginehelper import ocp_from_problem
    s = BatchedAcadosOcpSolver(ocp_from_problem(P), batch=6)
    s.options_set("riccati_precision", 32)
    r32 = engine_solve(P, b.x0, b.p, b.lh, b.yref, b.yref_e, solver=s)
    print("fp32 status", r32["status"], "max|dx|", np.abs(r32["x"] - a["x"]).max())
    from mpc_collisionavoidance_b200.workloads import guidance_ca1_ocp, make_guidance_batch
    g = make_guidance_batch(2)
    sg = BatchedAcadosOcpSolver(guidance_ca1_ocp(), batch=2)
    sg.options_set("cold_start", 1)
    sg.set(0, "lbx", g.x0); sg.set(0, "ubx", g.x0); sg.set("every", "p", g.p); sg.constraints_set("every", "lh", g.lh)
    sg.set("every", "yref", g.yref); sg.set(100, "yref", g.yref_e)
    print("guidance rti status", sg.solve(), "u0", sg.get(0, "u").ravel())
PY
for tool in ${TOOLS:-memcheck racecheck synccheck}; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python /tmp/san.py > gpurun_out/sanitizer_$tool.log 2>&1
  tail -6 gpurun_out/sanitizer_$tool.log
done
