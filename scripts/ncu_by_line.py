"""Aggregate an ncu source page (`ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > X.csv`) by source
line and by enclosing function of csrc/cta_kernel.cuh: warp-stall samples (= where resident warps spend their time)
and warp instructions executed.   python scripts/ncu_by_line.py X.csv [top_lines]"""
import csv
import os
import re
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def functions(path):
    """(first line, name) of every function-like definition in the file"""
    out = []
    pat = re.compile(r"^\s*(?:template\s*<[^>]*>\s*)?(?:MDEVNI|MDEV|DEV|__global__|static|inline)\b.*?\b([A-Za-z_][A-Za-z_0-9]*)\s*\(")
    for i, line in enumerate(open(path), 1):
        m = pat.match(line)
        if m and not line.strip().endswith(";"):
            out.append((i, m.group(1)))
    return out


def main():
    rows = csv.reader(open(sys.argv[1], newline=""))
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    cur_file, col = None, None
    by_line = defaultdict(lambda: [0, 0, 0])   # samples, warp instructions, barrier-stall samples
    cur_line = None
    for r in rows:
        if not r:
            continue
        if r[0] in ("File Name", "File Path"):
            cur_file = os.path.basename(r[1]); continue
        if r[0] == "Function Name" or r[0] == "Kernel Name":
            continue
        if r[0] == "Line No":
            col = {h: i for i, h in enumerate(r)}
            ncol = len(r)
            continue
        if cur_file is None or col is None:
            continue
        if r[0] == "" or len(r) != ncol or r[col["Address"]] != "-":
            continue            # SASS rows: the source-line row above them already carries their sums
        cur_line = int(r[0])
        try:
            smp = int(r[col["# Samples"]] or 0); ins = int(r[col["Instructions Executed"]] or 0)
            bar = int(r[col["stall_barrier"]] or 0)
        except ValueError:
            continue
        e = by_line[(cur_file, cur_line)]
        e[0] += smp; e[1] += ins; e[2] += bar
    tot_s = sum(v[0] for v in by_line.values()); tot_i = sum(v[1] for v in by_line.values())
    print(f"total samples {tot_s}, warp instructions {tot_i}")
    fn = {}
    for f in set(k[0] for k in by_line):
        p = os.path.join(ROOT, "mpc_collisionavoidance_b200", "csrc", f)
        fn[f] = functions(p) if os.path.exists(p) else []
    by_fn = defaultdict(lambda: [0, 0, 0])
    for (f, ln), v in by_line.items():
        name = "?"
        for start, nm in fn.get(f, []):
            if start <= ln:
                name = nm
        e = by_fn[(f, name)]
        for i in range(3):
            e[i] += v[i]
    print("\n== by function: samples %, non-barrier samples %, warp instructions %")
    nb_tot = sum(v[0] - v[2] for v in by_fn.values())
    for (f, name), v in sorted(by_fn.items(), key=lambda kv: -kv[1][0]):
        if v[0] * 200 < tot_s:
            continue
        print(f"  {f:18s} {name:22s} {100 * v[0] / tot_s:6.2f}  {100 * (v[0] - v[2]) / nb_tot:6.2f}  {100 * v[1] / tot_i:6.2f}")
    print(f"\n== top {top} lines by non-barrier samples: file:line samples% nonbarrier% instr%")
    for (f, ln), v in sorted(by_line.items(), key=lambda kv: -(kv[1][0] - kv[1][2]))[:top]:
        print(f"  {f}:{ln:<5d} {100 * v[0] / tot_s:6.2f} {100 * (v[0] - v[2]) / nb_tot:6.2f} {100 * v[1] / tot_i:6.2f}")


if __name__ == "__main__":
    main()
