"""Build profile variants: python scripts/build_prof.py <name>:<min_ctas>[:DEF=VAL,...] ... -> libusvmpc_<name>.so with
-DUSVMPC_PROFILE=1700"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from mpc_collisionavoidance_b200 import build as b

for a in sys.argv[1:]:
    parts = a.split(":")
    name, c = parts[0], int(parts[1])
    defs = tuple(parts[2].split(",")) if len(parts) > 2 and parts[2] else ()
    prof = () if name.startswith("np") else ("USVMPC_PROFILE=1700",)
    out = os.path.join(b.HERE, f"libusvmpc_{name}.so")
    b.build(force=True, min_ctas=c, out=out, defines=prof + defs)
    print(out)
