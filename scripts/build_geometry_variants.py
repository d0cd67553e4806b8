"""Build launch-geometry variants of libusvmpc.so: name=WPCxMINCTAS (warps per CTA x resident CTAs per SM the register
allocation allows), e.g. `4x2 1x16 2x8`.  Outputs libusvmpc_w<W>c<C>.so next to the product library."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from mpc_collisionavoidance_b200 import build as b

for a in sys.argv[1:]:
    w, c = (int(v) for v in a.split("x"))
    out = os.path.join(b.HERE, f"libusvmpc_w{w}c{c}.so")
    b.build(force=True, min_ctas=c, out=out, defines=(f"USVMPC_WPC={w}",), verbose=True)
    print(out)
