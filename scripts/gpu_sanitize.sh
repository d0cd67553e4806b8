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
PY
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python /tmp/san.py > gpurun_out/sanitizer_$tool.log 2>&1
  tail -6 gpurun_out/sanitizer_$tool.log
done
