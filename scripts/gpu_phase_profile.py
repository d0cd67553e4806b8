"""Phase clocks of single instances (diagnostic).  Needs a library built with -DUSVMPC_PROFILE=<qp_iter threshold>
(USVMPC_LIB points at it): instances whose total QP iteration count reaches the threshold print their per-phase
clock64() totals.  Solves the benchmark batch (clocks under load), then the slowest instances alone (one warp on an
otherwise idle GPU = the latency that bounds the batch time)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from mpc_collisionavoidance_b200 import BatchedAcadosOcpSolver  # noqa: E402
from mpc_collisionavoidance_b200.workloads import CONFIGS, benchmark_ocp, make_batch  # noqa: E402


def solve(cfg_id, b, sel):
    N = CONFIGS[cfg_id]["N"]
    B = len(sel)
    s = BatchedAcadosOcpSolver(benchmark_ocp(cfg_id), batch=B)
    s.options_set("cold_start", 1)
    x0 = b.x0[sel]
    s.set(0, "lbx", x0); s.set(0, "ubx", x0)
    s.set("every", "p", b.p[sel]); s.constraints_set("every", "lh", b.lh[sel])
    s.set("every", "yref", b.yref[sel]); s.set(N, "yref", b.yref_e[sel])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    s.solve()
    torch.cuda.synchronize()
    return s.stats_table(), (time.perf_counter() - t0) * 1e3


if __name__ == "__main__":
    cfg_id = 2
    b = make_batch(cfg_id, seed=1234 + cfg_id)   # rank-0 batch of bench.py
    allsel = np.arange(len(b.x0))
    solve(cfg_id, b, allsel)
    st, ms = solve(cfg_id, b, allsel)
    print(f"full batch {ms:.1f} ms", flush=True)
    order = np.argsort(-st[:, 2])[:3]
    print("slowest", order, st[order, 2], flush=True)
    for i in order:
        st1, ms1 = solve(cfg_id, b, np.array([i]))
        print(f"alone inst {i}: status {st1[0, 0]:.0f} sqp {st1[0, 1]:.0f} qp {st1[0, 2]:.0f}  {ms1:.1f} ms", flush=True)
    # the 148 slowest together: one straggler per SM (the tail regime of the full batch)
    sel = np.argsort(-st[:, 2])[:148]
    st2, ms2 = solve(cfg_id, b, sel)
    print(f"148 slowest together: {ms2:.1f} ms", flush=True)
