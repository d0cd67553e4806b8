"""Solve the slowest instance of the benchmark batch alone (B = 1): the latency that bounds the batch time.  Run under ncu
for per-line stall attribution of a single warp."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from gpu_phase_profile import solve  # noqa: E402
from mpc_collisionavoidance_b200.workloads import make_batch  # noqa: E402

if __name__ == "__main__":
    inst = int(sys.argv[1]) if len(sys.argv) > 1 else 1089
    b = make_batch(2, seed=1236)
    st, ms = solve(2, b, np.array([inst]))
    print(f"alone inst {inst}: status {st[0, 0]:.0f} sqp {st[0, 1]:.0f} qp {st[0, 2]:.0f}  {ms:.1f} ms")
