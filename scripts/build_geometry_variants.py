"""Build launch-geometry variants of libusvmpc.so: name=THREADSxMINCTAS[xCHAINWARPS] (threads per block x resident
blocks per SM the register allocation allows x warps sharing a factorisation step), e.g. `256x2 128x3 128x4x1`.
Outputs libusvmpc_t<T>c<C>w<W>.so next to the product library."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from mpc_collisionavoidance_b200 import build as b

for a in sys.argv[1:]:
    v = [int(x) for x in a.split("x")]
    t, c, w = v[0], v[1], (v[2] if len(v) > 2 else 2)
    out = os.path.join(b.HERE, f"libusvmpc_t{t}c{c}w{w}.so")
    b.build(force=True, min_ctas=c, out=out, defines=(f"USVMPC_THREADS={t}", f"USVMPC_CHAIN_WARPS={w}"))
    print(out)
