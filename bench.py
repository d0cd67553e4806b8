#!/usr/bin/env python
"""bench.py -- NMPC solves/sec of the batched engine on BASELINE.json's configurations (default: the headline, config 2).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config 1..5] [--batch B]
                    [--riccati-precision 32|64] [--strong] [--no-cpu]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" = one full solve (SQP to the 1e-6 KKT tolerance, cold start x_k = x0, u = 0, pi = 0, max 100 iterations --
the reference's semantics) of one synthetic batch of `batch` independent instances per GPU
(weak scaling; --strong splits the config's batch over the GPUs; config 5 is 131072 instances over 8 GPUs = 16384 per GPU).
  value     instances / s over all ranks, inputs resident in HBM, CUDA events around every one of the K steps on the
            launching stream, 256 MB written between steps (L2 flush, outside the timed intervals), max over ranks;
            for N > 1 each step includes the in-place all-gather of the packed results
  converged_solves_per_s   the same counting only instances that return status 0
  e2e       the same through the public API with pinned HOST buffers: H2D of the step's inputs, solve, D2H of the
            packed trajectories + statistics inside the timed region (phases event-timed: h2d_ms, solve_ms, d2h_ms)
  roofline  algorithmic HBM bytes of the solve kernel (SURVEY.md 8d streaming model x iteration counts the kernel
            actually executed) / its CUDA-event duration, against MEASURED_PEAKS.json; `traffic` = measured DRAM bytes
            per launch of the same (config, batch) from the committed ncu capture (profiles/r2_ncu_traffic.json)
  cpu_baseline  the reference's own CPU implementation (oracle/_ref, else the C port) on a bounded sample, rank 0, N=1;
            value = sample / seconds inside the solves of the busiest thread (solver construction excluded)
`--impl reference` times that CPU implementation alone on the same workload definition (same `config` object).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from mpc_collisionavoidance_b200.workloads import CONFIGS, benchmark_ocp, make_batch  # noqa: E402

METRIC = "NMPC solves/sec (USV 3-DOF, N=40, 5 obstacles) at batch=4096"
UNIT = "solves/s"


def metric_name(cfg_id):
    if cfg_id == 2:
        return METRIC
    c = CONFIGS[cfg_id]
    return f"NMPC solves/sec (USV 3-DOF, N={c['N']}, {c['K']} obstacles) at batch={c['B']}"


def problem_for(cfg_id, nlp_type=0):
    import refharness as rh
    c = CONFIGS[cfg_id]
    return rh.RefProblem(N=c["N"], K=c["K"], num_steps=c["num_steps"], nlp_type=nlp_type)


def per_gpu_batch(cfg_id, world, strong=False, override=0):
    if override:
        return override
    B = CONFIGS[cfg_id]["B"]
    if cfg_id == 5:
        return B // 8 if not strong else B // world       # 131072 instances sharded over 8 GPUs: 16384 per GPU
    return B // world if strong else B


def bench_config(cfg_id, B, world, strong):
    """the `config` object of the JSON line: identical in the engine arm and the reference arm"""
    c = CONFIGS[cfg_id]
    return {"workload": (f"configs[{cfg_id - 1}]: batch={B} independent instances per GPU, USV 3-DOF nx=6 nu=2, N={c['N']}, "
                         f"{c['K']} obstacles, fp64, full SQP tol 1e-6 max_iter 100, ERK4 x{c['num_steps']} steps, cold start"),
            "global_batch": world * B,
            "parallelism": (f"batch sharded over {world} GPU(s), one all-gather of results" if world > 1 else "1 GPU"),
            "scaling": "strong" if strong else "weak"}


# ------------------------------------------------------------------------------------------------ CPU arm
class quiet_stdout:
    """the reference C stack printf()s on every non-converged solve; keep that out of the JSON line on stdout"""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        self.null = os.open(os.devnull, os.O_WRONLY)
        os.dup2(self.null, 1)

    def __exit__(self, *a):
        os.dup2(self.saved, 1)
        os.close(self.null)
        os.close(self.saved)


def cpu_reference_run(cfg_id, nsample, nthreads, seed=None):
    """time the reference's CPU implementation on the first `nsample` instances of the workload.  The time is the
    harness' own clock around the solves (batch generation and acados_create() are outside): returns
    (solves/s, kind, result dict)."""
    import refharness as rh
    P = problem_for(cfg_id)
    b = make_batch(cfg_id, B=nsample, seed=seed)
    if rh.available():
        try:
            with quiet_stdout():
                r = rh.solve_batch(P, b.x0, b.p, b.lh, b.yref, b.yref_e, nthreads=nthreads)
            return nsample / r["seconds"], "reference", r
        except OSError:
            pass
    import oracleport as op
    if not op.available():
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "port", "CC=gcc"], stdout=subprocess.DEVNULL)
    r = op.solve_batch(P, b.x0, b.p, b.lh, b.yref, b.yref_e, nthreads=nthreads)
    return nsample / r["seconds"], "port", r


def cpu_baseline(cfg_id, target_seconds=15.0):
    cores = os.cpu_count() or 1
    v0, kind, _ = cpu_reference_run(cfg_id, 2 * cores, cores)       # calibrate
    n = int(max(2 * cores, min(4096, v0 * target_seconds)))
    v, kind, r = cpu_reference_run(cfg_id, n, cores)
    conv = int((r["status"] == 0).sum())
    return {"value": round(v, 2), "unit": UNIT, "cores": cores, "kind": kind,
            "converged_solves_per_s": round(v * conv / n, 2),
            "sample": f"first {n} instances of the same seeded workload, {cores} threads, one solver per thread, "
                      f"{r['seconds']:.1f} s inside the solves, {conv}/{n} converged"}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    cfg_id = args.config
    v0, kind, _ = cpu_reference_run(cfg_id, 2 * cores, cores)
    budget = 120.0 / max(1, args.steps + args.warmup)            # whole run ends within a few minutes
    B = per_gpu_batch(cfg_id, world, args.strong, args.batch)
    n = int(max(cores, min(B, v0 * min(budget, 20.0))))
    for _ in range(args.warmup):
        cpu_reference_run(cfg_id, n, cores)
    secs, conv = 0.0, 0
    for _ in range(args.steps):
        v, kind, r = cpu_reference_run(cfg_id, n, cores)
        secs += r["seconds"]
        conv += int((r["status"] == 0).sum())
    value = args.steps * n / secs
    line = {"impl": "reference", "metric": metric_name(cfg_id), "value": round(value, 2), "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(1e3 * secs / args.steps, 3),
            "higher_is_better": True, "scaling": "strong" if args.strong else "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": bench_config(cfg_id, B, world, args.strong),
            "converged_solves_per_s": round(conv / secs, 2),
            "workload_stats": {"converged_frac": round(conv / (args.steps * n), 4),
                               "note": f"CPU arm: each step solves the first {n} instances of the batch"},
            "cpu_baseline": {"value": round(value, 2), "unit": UNIT, "cores": cores, "kind": kind,
                             "sample": f"first {n} instances per step of the same seeded workload, {cores} threads, time inside the solves"},
            "e2e": {"value": round(value, 2), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self._stop, self.th = index, [], threading.Event(), None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self.th = threading.Thread(target=self._run, daemon=True)
        self.th.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self.th.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons,
                "samples": len(self.rows), "power_w_max": max(float(r[2]) for r in self.rows)}


# ------------------------------------------------------------------------------------------------ roofline
def algorithmic_bytes(cfg_id, sqp_iters, ipm_iters, B):
    """SURVEY.md section 8(d) streaming model, 8-byte words, summed over the batch with the iteration counts the kernel
    reports: per stage and IPM iteration W_ipm = Q + 4L + 2S, per stage and SQP iteration W_lin = Q + (nv+nx+rows)."""
    c = CONFIGS[cfg_id]
    N, K = c["N"], c["K"]
    nx, nu = 6, 2
    nv = nx + nu
    rows = 2 * (2 + 3 + K)
    Q = nx * nv + nx + nv + 2 * K + rows
    L = nv * (nv + 1) // 2 + nv
    S = nv + nx + 3 * rows
    w_ipm, w_lin = Q + 4 * L + 2 * S, Q + (nv + nx + rows)
    io = B * ((nx + 2 * K + K + nv) + ((N + 1) * nx + N * nu + 6))
    return 8.0 * (sqp_iters * N * w_lin + ipm_iters * (N + 1) * w_ipm + io)


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def measured_traffic(cfg_id, B):
    """DRAM bytes per launch of the solve kernel from the committed `ncu --set full` capture of this workload
    (profiles/r2_ncu_traffic.json, written by scripts/ncu_summary.py), or None"""
    p = os.path.join(ROOT, "profiles", "r2_ncu_traffic.json")
    if os.path.exists(p):
        for e in json.load(open(p)):
            if e.get("config") == cfg_id and e.get("batch") == B:
                return e
    return None


# ------------------------------------------------------------------------------------------------ GPU arm
def run_engine(args):
    import torch
    import torch.distributed as dist
    from mpc_collisionavoidance_b200 import BatchedAcadosOcpSolver
    from mpc_collisionavoidance_b200 import dist as D

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the engine has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    dev = torch.device(f"cuda:{local}")
    cfg_id = args.config
    c = CONFIGS[cfg_id]
    B = per_gpu_batch(cfg_id, world, args.strong, args.batch)
    N, K, nx, nu = c["N"], c["K"], 6, 2
    if args.strong:
        # one global batch (rank 0's seed), rank r solves its contiguous slice
        full = make_batch(cfg_id, B=world * B, seed=1234 + cfg_id)
        lo, hi = D.shard_range(world * B, rank, world)
        from mpc_collisionavoidance_b200.workloads import Batch
        batch = Batch(full.x0[lo:hi], full.p[lo:hi], full.lh[lo:hi], full.yref[lo:hi], full.yref_e[lo:hi])
    else:
        batch = make_batch(cfg_id, B=B, seed=1234 + cfg_id + 1000 * rank)   # weak scaling: every rank its own batch
    s = BatchedAcadosOcpSolver(benchmark_ocp(cfg_id), batch=B, device=local)
    s.options_set("cold_start", 1)
    if args.riccati_precision == 32:
        s.options_set("riccati_precision", 32)     # BASELINE.json config 4: fp32 factorisation, fp64 residuals + refinement
    s.sync_host_sets = False
    names = ("x0", "p", "lh", "yref", "yref_e")
    host = {k: torch.from_numpy(np.ascontiguousarray(getattr(batch, k))).pin_memory() for k in names}
    devt = {k: v.to(dev) for k, v in host.items()}
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)   # 256 MB > 126 MB L2

    def set_inputs(src):
        s.set(0, "lbx", src["x0"]); s.set(0, "ubx", src["x0"])
        s.set("every", "p", src["p"]); s.constraints_set("every", "lh", src["lh"])
        s.set("every", "yref", src["yref"]); s.set(N, "yref", src["yref_e"])

    # the one collective of the path: the solve's epilogue writes the packed result rows into this rank's
    # slice of the gathered tensor, the all-gather runs in place
    gathered, own = D.gather_buffer(torch, B, world, rank, N, nx, nu, dev)
    s.set_result_buffer(own)

    def step_device():
        s.solve_async()
        if world > 1:
            D.all_gather_in_place(dist, gathered, own)

    out_host = {"packed": torch.empty((B, own.shape[1]), dtype=torch.float64).pin_memory(),
                "stats": torch.empty((B, 16), dtype=torch.float64).pin_memory()}
    e2e_ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]

    def step_e2e():
        e2e_ev[0].record()
        set_inputs(host)                                   # H2D from pinned host memory
        e2e_ev[1].record()
        s.solve_async()
        e2e_ev[2].record()
        out_host["packed"].copy_(own, non_blocking=True)   # D2H of the result: trajectories + status / residuals
        out_host["stats"].copy_(s.stats_table(device=True), non_blocking=True)
        e2e_ev[3].record()
        torch.cuda.current_stream().synchronize()
        return [e2e_ev[i].elapsed_time(e2e_ev[i + 1]) for i in range(3)]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    set_inputs(devt)
    torch.cuda.synchronize()
    for _ in range(args.warmup):
        step_device()
    barrier()
    # ---- device-resident timing: CUDA events on the launching (current torch) stream around each of the K steps, L2
    # flushed (256 MB written) between the steps and outside the timed intervals
    l0 = s.info("launches")
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    with ClockSampler(local) as clk:
        barrier()
        for i in range(args.steps):
            flush.zero_()
            barrier()
            ev[i][0].record()
            step_device()
            ev[i][1].record()
        barrier()
    launches = int(s.info("launches") - l0)
    total_ms = sum(a.elapsed_time(b) for a, b in ev)
    st = s.stats_table()
    t_max = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_max, op=dist.ReduceOp.MAX)
    total_ms = float(t_max.item())
    value = world * B * args.steps / (total_ms * 1e-3)

    # ---- dominant kernel alone (no collective): duration for the roofline
    barrier()
    kern_ms = 0.0
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(args.steps):
        flush.zero_()
        k0.record()
        s.solve_async()
        k1.record()
        torch.cuda.synchronize()
        kern_ms += k0.elapsed_time(k1) / args.steps
    sqp_sum, ipm_sum = float(st[:, 1].sum()), float(st[:, 2].sum())
    abytes = algorithmic_bytes(cfg_id, sqp_sum, ipm_sum, B)
    peak, peak_src = measured_peak()
    achieved = abytes / (kern_ms * 1e-3) / 1e9
    traffic = measured_traffic(cfg_id, B)

    # ---- end to end through the public API with host buffers
    for _ in range(max(1, args.warmup // 2)):
        step_e2e()
    barrier()
    phases = np.zeros(3)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        phases += np.array(step_e2e())
    barrier()
    e2e_s = time.perf_counter() - t0
    t_e2e = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e_value = world * B * args.steps / float(t_e2e.item())
    h2d = sum(host[k].numel() * 8 for k in names) + host["x0"].numel() * 8   # x0 goes in twice (lbx and ubx)
    d2h = sum(v.numel() * 8 for v in out_host.values())
    ok = st[:, 0] == 0
    x_ok = bool(np.isfinite(out_host["packed"].numpy()[ok]).all())
    n_ok = torch.tensor([float(ok.sum())], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(n_ok)
    conv_total = float(n_ok.item())

    if rank == 0:
        line = {"metric": metric_name(cfg_id), "value": round(value, 1), "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": round(total_ms / args.steps, 3), "higher_is_better": True,
                "scaling": "strong" if args.strong else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": bench_config(cfg_id, B, world, args.strong),
                "converged_solves_per_s": round(conv_total * args.steps / (total_ms * 1e-3), 1),
                "riccati_precision": args.riccati_precision,
                "workload_stats": {"converged_frac": round(float(ok.mean()), 4), "mean_sqp_iter": round(sqp_sum / B, 2),
                                   "mean_qp_iter": round(ipm_sum / B, 2), "max_sqp_iter": int(st[:, 1].max()),
                                   "lq_fact_iterations": int(st[:, 7].sum()), "refinement_solves": int(st[:, 11].sum()),
                                   "fp32_factorisations": int(st[:, 15].sum()),
                                   "finite_outputs": x_ok,
                                   "l2": "256 MB written between timed steps (L2 flush, outside the timed intervals)",
                                   "hbm_resident_bytes": int(s.info("workspace_bytes")),
                                   "blocks_per_sm": int(s.info("ctas_per_sm")), "smem_bytes_per_block": int(s.info("smem_bytes_per_cta")),
                                   "l2_scratch_bytes_per_block": int(s.info("scratch_bytes_per_cta")),
                                   "queue_order": "longest-first from the previous solve's iteration counts (warm-up solves)"},
                "roofline": {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                             "frac": round(achieved / peak, 4),
                             "traffic": traffic["dram_bytes_per_launch"] if traffic else None,
                             "traffic_source": traffic["source"] if traffic else None,
                             "kernel": "nmpc_solve_kernel<Usv3>", "kernel_ms": round(kern_ms, 3),
                             "algorithmic_bytes_per_launch": abytes, "peak_source": peak_src},
                "e2e": {"value": round(e2e_value, 1), "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                        "ms_per_step": round(1e3 * float(t_e2e.item()) / args.steps, 3),
                        "h2d_ms": round(phases[0] / args.steps, 3), "solve_ms": round(phases[1] / args.steps, 3),
                        "d2h_ms": round(phases[2] / args.steps, 3)},
                "gpu_launches": launches, "clocks": clk.summary()}
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline(cfg_id)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--config", type=int, default=2, help="BASELINE.json config (1-based); 2 = the headline")
    ap.add_argument("--batch", type=int, default=0, help="override the per-GPU batch (default: the config's)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--riccati-precision", type=int, default=64, choices=[32, 64],
                    help="32: Riccati factorisation in fp32 with fp64 residuals and refinement (BASELINE.json config 4's variant)")
    ap.add_argument("--strong", action="store_true", help="strong scaling: the config's batch is split over the GPUs")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "engine":
        args.warmup = 3
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_engine(args)


if __name__ == "__main__":
    main()
