"""RTI step time of the DEPLOYED collision-avoidance solver (usv_model_guidance_ca1: nx = 8, nu = 1, N = 100, 8 soft obstacle
rows) on the engine: one instance (what the ROS node runs at 20 Hz) and batches (Monte-Carlo over scenes), warm-started
steps, CUDA-event time per step.  Diagnostic numbers for profiles/README.md, not the headline metric."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from mpc_collisionavoidance_b200 import BatchedAcadosOcpSolver  # noqa: E402
from mpc_collisionavoidance_b200.workloads import guidance_ca1_ocp, make_guidance_batch  # noqa: E402

if __name__ == "__main__":
    out = []
    for B in (1, 256, 4096):
        g = make_guidance_batch(B)
        s = BatchedAcadosOcpSolver(guidance_ca1_ocp(), batch=B)
        s.options_set("cold_start", 1)
        s.set(0, "lbx", g.x0); s.set(0, "ubx", g.x0); s.set("every", "p", g.p); s.constraints_set("every", "lh", g.lh)
        s.set("every", "yref", g.yref); s.set(100, "yref", g.yref_e)
        s.solve()
        s.options_set("cold_start", 0)
        x = torch.as_tensor(g.x0, device="cuda")
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        steps = 20
        for _ in range(3):
            s.solve_async()
        torch.cuda.synchronize()
        ev[0].record()
        for _ in range(steps):
            s.set(0, "lbx", x); s.set(0, "ubx", x)
            s.solve_async()
            x = s.get(1, "x", device=True).contiguous()
        ev[1].record()
        torch.cuda.synchronize()
        ms = ev[0].elapsed_time(ev[1]) / steps
        st = s.stats_table()
        out.append({"batch": B, "ms_per_rti_step": round(ms, 3), "rti_steps_per_s": round(B / ms * 1e3, 1), "status_0_frac": float((st[:, 0] == 0).mean()),
                    "mean_qp_iter": float(st[:, 2].mean()), "blocks_per_sm": int(s.info("ctas_per_sm")),
                    "smem_bytes_per_block": int(s.info("smem_bytes_per_cta"))})
        print(json.dumps(out[-1]), flush=True)
