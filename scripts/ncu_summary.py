"""Digest `ncu --set full` captures of the solve kernel into the files bench.py and the judge read:

    python scripts/ncu_summary.py gpurun_out/r2_cfg2_b4096.ncu-rep:2:4096 gpurun_out/r2_cfg2_b32768.ncu-rep:2:32768 ...

For every REPORT:CONFIG:BATCH it writes profiles/<name>_summary.txt (the metrics named in
/opt/skills/guides/B200_PROFILING.md for the solve kernel's launch) and adds an entry to profiles/r2_ncu_traffic.json
(`dram__bytes_read.sum + dram__bytes_write.sum` per launch), which bench.py reports as roofline.traffic."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed.sum", "smsp__inst_executed.sum",
        "sm__inst_executed.avg.per_cycle_elapsed", "smsp__issue_active.avg.pct", "smsp__issue_active.avg.per_cycle_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__shared_mem_per_block_dynamic",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.sum", "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_fma.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__cycles_elapsed.max", "sm__cycles_active.avg"]


def raw_rows(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    head, units = rows[0], rows[1]
    return head, units, rows[2:]


def metrics_log(path, cfg, batch, traffic):
    """a `ncu --metrics ... --csv --log-file` capture (one row per metric) of the solve kernel"""
    name = os.path.splitext(os.path.basename(path))[0]
    text = open(path).read()
    text = text[text.index('"ID"'):]
    rows = list(csv.DictReader(io.StringIO(text)))
    lines = [f"# {name}: ncu --metrics (traffic / issue) --clock-control none, config {cfg}, batch {batch}"]
    val = {}
    for r in rows:
        if "nmpc_solve_kernel" not in r["Kernel Name"]:
            continue
        u = r["Metric Unit"]
        lines.append(f"  {r['Metric Name']:70s} {r['Metric Value']:>20s} {u}")
        try:
            v = float(r["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        val[r["Metric Name"]] = v * {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    dram = val["dram__bytes_read.sum"] + val["dram__bytes_write.sum"]
    lines.append(f"  -> dram bytes per launch: {dram:.0f}")
    entry = {"config": int(cfg), "batch": int(batch), "dram_bytes_per_launch": int(dram),
             "source": f"profiles/{name}_summary.txt (ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum)"}
    traffic[:] = [e for e in traffic if not (e["config"] == entry["config"] and e["batch"] == entry["batch"])] + [entry]
    open(os.path.join(ROOT, "profiles", name + "_summary.txt"), "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


def main():
    traffic_path = os.path.join(ROOT, "profiles", "r2_ncu_traffic.json")
    traffic = json.load(open(traffic_path)) if os.path.exists(traffic_path) else []
    for spec in sys.argv[1:]:
        rep, cfg, batch = spec.split(":")
        if rep.endswith(".csv"):
            metrics_log(rep, cfg, batch, traffic)
            continue
        name = os.path.splitext(os.path.basename(rep))[0]
        head, units, rows = raw_rows(rep)
        col = {h: i for i, h in enumerate(head)}
        lines = [f"# {name}: ncu --set full --clock-control none, config {cfg}, batch {batch}; one row per captured launch"]
        for r in rows:
            kname = r[col["Kernel Name"]]
            lines.append(f"kernel: {kname[:110]}")
            for k in KEEP:
                if k in col:
                    lines.append(f"  {k:90s} {r[col[k]]:>18s} {units[col[k]]}")
            if "nmpc_solve_kernel" in kname:
                def num(k):
                    v, u = float(r[col[k]].replace(",", "")), units[col[k]]
                    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
                dram = num("dram__bytes_read.sum") + num("dram__bytes_write.sum")
                entry = {"config": int(cfg), "batch": int(batch), "dram_bytes_per_launch": int(dram),
                         "source": f"profiles/{name}_summary.txt (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum)"}
                traffic = [e for e in traffic if not (e["config"] == entry["config"] and e["batch"] == entry["batch"])] + [entry]
                lines.append(f"  -> dram bytes per launch: {dram:.0f}")
        open(os.path.join(ROOT, "profiles", name + "_summary.txt"), "w").write("\n".join(lines) + "\n")
        print("\n".join(lines))
    json.dump(sorted(traffic, key=lambda e: (e["config"], e["batch"])), open(traffic_path, "w"), indent=1)


if __name__ == "__main__":
    main()
