"""Build register-allocation variants of libusvmpc.so (different __launch_bounds__ min-CTAs) for occupancy sweeps."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from mpc_collisionavoidance_b200 import build as b

for m in (int(a) for a in sys.argv[1:]):
    out = os.path.join(b.HERE, f"libusvmpc_c{m}.so")
    b.build(force=True, min_ctas=m, out=out)
    print(out)
