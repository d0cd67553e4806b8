"""Resource usage and instruction mix of the solve kernels in libusvmpc.so -> profiles/r2_sass_digest.txt
(cuobjdump -res-usage / -sass on the built library; no GPU needed)."""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "mpc_collisionavoidance_b200", "libusvmpc.so")
KEYS = ["DMMA", "DFMA", "DMUL", "DADD", "DSETP", "MUFU", "LDS", "STS", "LDG", "STG", "LDL", "STL", "LDC", "LDCU", "REDUX", "SHFL", "BAR",
        "WARPSYNC", "IMAD", "ISETP", "BRA", "CALL", "ATOMG", "UBLKCP", "UTMALDG", "UTCHMMA", "UTCQMMA"]


def main():
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    out = ["# libusvmpc.so (sm_100a): resources and instruction mix of every kernel; static counts of SASS mnemonics",
           "# (UBLKCP / UTMALDG = TMA, UTC*MMA = tcgen05: none on this path, by design -- DESIGN.md section 2.2)", ""]
    names = re.findall(r"Function (\S+):\n\s*(REG:\d+.*)", res)
    for n, r in names:
        dem = subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
        out.append(f"{dem[:110]}\n    {r}")
    out.append("")
    cur, counts = None, collections.OrderedDict()
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            counts[cur] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur:
            counts[cur][m.group(1)] += 1
            counts[cur]["total"] += 1
    for k, c in counts.items():
        out.append(f"{k[:110]}\n    total {c['total']}  " + "  ".join(f"{m} {c[m]}" for m in KEYS if c[m]))
    open(os.path.join(ROOT, "profiles", "r2_sass_digest.txt"), "w").write("\n".join(out) + "\n")
    print("\n".join(out))


if __name__ == "__main__":
    main()
