"""Generate tests/golden/*.npz from the UNMODIFIED reference stack (oracle/_ref, built by oracle/Makefile
from /root/reference).  Run here (the reference does not exist on the GPU box):

    make -C oracle ref && python tests/make_golden.py

Fixtures:
  usv_cfg{1,2,3}_solve.npz  inputs + reference outputs (x, u, status, sqp_iter, qp_iter, res) of full SQP solves
  usv_cfg4_solve.npz        64 instances of config 4 (N = 100; scene ranges of workloads.XREF_FACTOR)
  usv_cfg5_solve.npz        256 Monte-Carlo draws of config 5 (128 base scenes x 2 draws of the x0 disturbance)
  usv_cfg2_lq.npz           the instances of the headline batch on which the reference takes its LQ path
  usv_guidance_ca1.npz      the deployed CA solver (nx = 8, nu = 1, N = 100, 8 soft obstacle rows): RTI step and SQP solve
  usv_unconstrained.npz     an OCP without inequality rows (HPIPM's nc = 0 path: direct solve, zero IPM iterations)
  usv_cfg1_rti.npz          one SQP_RTI step (known answer 2 of SURVEY.md appendix B)
  usv_cfg2_qp.npz           QPs captured at HPIPM's door (after x0 elimination) with HPIPM's solution and
                            iteration count, for QP-level parity of the IPM/Riccati kernels
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, ".."))

import refharness as rh  # noqa: E402
from mpc_collisionavoidance_b200.workloads import CONFIGS, make_batch  # noqa: E402

G = os.path.join(HERE, "golden")


def solve_fixture(cfg_id, B, name):
    c = CONFIGS[cfg_id]
    b = make_batch(cfg_id, B=B)
    P = rh.RefProblem(N=c["N"], K=c["K"], num_steps=c["num_steps"])
    o = rh.solve_batch(P, b.x0, b.p, b.lh, b.yref, b.yref_e, nthreads=8)
    np.savez_compressed(os.path.join(G, name), cfg_id=cfg_id, N=c["N"], K=c["K"], num_steps=c["num_steps"], x0=b.x0, p=b.p,
                        lh=b.lh, yref=b.yref, yref_e=b.yref_e, x=o["x"], u=o["u"], status=o["status"],
                        sqp_iter=o["sqp_iter"], qp_iter=o["qp_iter"], res=o["res"], lq_calls=o["lq_calls"])
    print(name, "status", np.bincount(o["status"]), "sqp_iter", o["sqp_iter"])


def guidance_fixture():
    """the deployed CA solver (usv_guidance_ca1: nx = 8, nu = 1, N = 100, 8 SOFT obstacle rows): one SQP_RTI step and a
    full SQP solve of 24 scenes, with the slack values"""
    from mpc_collisionavoidance_b200.workloads import make_guidance_batch
    b = make_guidance_batch(24)
    out = dict(x0=b.x0, p=b.p, lh=b.lh, yref=b.yref, yref_e=b.yref_e)
    for nlp_type, tag in ((1, "rti"), (0, "sqp")):
        P = rh.RefProblem(model=2, N=100, K=8, num_steps=1, nlp_type=nlp_type, nsh=8, lsh=-0.2, ush=0.0, zl=1.0, zu=1.0, uh=1e6,
                          max_iter=30)
        s = rh.RefSolver(P)
        rs = [s.solve(b.x0[i], b.p[i], b.lh[i], b.yref[i], b.yref_e[i]) for i in range(len(b.x0))]
        for k in ("x", "u", "sl", "su", "res", "lam", "t"):
            out[f"{tag}_{k}"] = np.stack([r[k] for r in rs])
        out[f"{tag}_stat"] = np.array([[r["status"], r["sqp_iter"], r["qp_iter"]] for r in rs])
        print("guidance", tag, out[f"{tag}_stat"].T)
    np.savez_compressed(os.path.join(G, "usv_guidance_ca1.npz"), **out)


def unconstrained_problem():
    W = np.diag([1, 1, 0.1, 10, 0.1, 0.1, 1e-3, 1e-3])
    none = np.array([])
    return rh.RefProblem(N=20, K=0, num_steps=2, max_iter=30, W=W, We=5 * W[:6, :6], lbu=none, ubu=none,
                         idxbx=np.array([], dtype=np.int32), lbx=none, ubx=none)


def unconstrained_fixture():
    """an OCP without any inequality row (no boxes, K = 0): after the x0 elimination HPIPM sees nc = 0 and takes its
    direct factorise-and-solve path with iter = 0 (x_ocp_qp_ipm.c:2458-2481)"""
    B = 6
    rng = np.random.default_rng(3)
    x0 = np.array([0, 0, 0.1, 0.7, 0, 0.0]) + 0.1 * rng.standard_normal((B, 6))
    yref = np.tile(np.array([6, 1, 0, 1, 0, 0, 0, 0.0]), (B, 1))
    yref[:, 1] += rng.uniform(-1, 1, B)
    p, lh = np.zeros((B, 0)), np.zeros((B, 0))
    o = rh.solve_batch(unconstrained_problem(), x0, p, lh, yref, yref[:, :6].copy())
    np.savez_compressed(os.path.join(G, "usv_unconstrained.npz"), x0=x0, yref=yref, x=o["x"], u=o["u"], status=o["status"],
                        sqp_iter=o["sqp_iter"], qp_iter=o["qp_iter"], res=o["res"])
    print("unconstrained: status", o["status"], "sqp_iter", o["sqp_iter"], "qp_iter", o["qp_iter"])


def lq_fixture():
    """instances of the headline batch on which the reference's IPM takes its LQ path (lq_fact = 1) + their neighbours"""
    c = CONFIGS[2]
    b = make_batch(2)
    P = rh.RefProblem(N=c["N"], K=c["K"], num_steps=c["num_steps"])
    o = rh.solve_batch(P, b.x0, b.p, b.lh, b.yref, b.yref_e, nthreads=8)
    hit = np.flatnonzero(o["lq_calls"] > 0)
    sel = sorted(set(int(i + dlt) for i in hit for dlt in (0, 1) if i + dlt < len(b.x0)))
    sel = np.array(sel[:32])
    np.savez_compressed(os.path.join(G, "usv_cfg2_lq.npz"), index=sel, x0=b.x0[sel], p=b.p[sel], lh=b.lh[sel], yref=b.yref[sel],
                        yref_e=b.yref_e[sel], x=o["x"][sel], u=o["u"][sel], status=o["status"][sel], sqp_iter=o["sqp_iter"][sel],
                        qp_iter=o["qp_iter"][sel], res=o["res"][sel], lq_calls=o["lq_calls"][sel])
    print("lq fixture: instances", hit, "status", o["status"][hit], "lq_calls", o["lq_calls"][hit])


def known_answer_fixture():
    x0 = np.array([0, 0, 0, 0.7, 0, 0.0]); p = np.array([2.0, 0.2, 3.5, -0.6, 5.0, 0.5]); lh = np.full(3, 0.8)
    yref = np.array([6, 0, 0, 1, 0, 0, 0, 0.0])
    out = {}
    for nlp_type, tag in ((0, "sqp"), (1, "rti")):
        s = rh.RefSolver(rh.RefProblem(N=20, K=3, num_steps=1, nlp_type=nlp_type))
        o = s.solve(x0, p, lh, yref, yref[:6])
        for k in ("x", "u", "pi", "lam", "t", "res"):
            out[f"{tag}_{k}"] = o[k]
        out[f"{tag}_stat"] = np.array([o["status"], o["sqp_iter"], o["qp_iter"]])
    np.savez_compressed(os.path.join(G, "usv_cfg1_known_answer.npz"), x0=x0, p=p, lh=lh, yref=yref, **out)
    print("known answer", out["sqp_stat"], out["rti_stat"])


def qp_fixture():
    c = CONFIGS[2]
    b = make_batch(2, B=4)
    P = rh.RefProblem(N=c["N"], K=c["K"], num_steps=c["num_steps"])
    s = rh.RefSolver(P)
    qps = {}
    n = 0
    for i in range(4):
        for want in (0, 2, 6):
            _, buf = s.solve_capture_qp(want, b.x0[i], b.p[i], b.lh[i], b.yref[i], b.yref_e[i])
            if not buf["got"]:
                continue
            for k in ("BAbt", "b", "RSQrq", "rqz", "DCt", "d", "idxb", "ux", "pi", "lam", "t", "dims"):
                qps[f"q{n}_{k}"] = buf[k]
            qps[f"q{n}_info"] = np.array([buf["iter"], buf["status"]])
            n += 1
    np.savez_compressed(os.path.join(G, "usv_cfg2_qp.npz"), n=n, **qps)
    print("qp fixtures", n, [int(qps[f"q{i}_info"][0]) for i in range(n)])


if __name__ == "__main__":
    assert rh.available(), "build oracle/_ref first: make -C oracle ref"
    known_answer_fixture()
    solve_fixture(1, 16, "usv_cfg1_solve.npz")
    solve_fixture(2, 32, "usv_cfg2_solve.npz")
    solve_fixture(3, 8, "usv_cfg3_solve.npz")
    solve_fixture(4, 64, "usv_cfg4_solve.npz")
    solve_fixture(5, 256, "usv_cfg5_solve.npz")
    qp_fixture()
    lq_fixture()
    guidance_fixture()
    unconstrained_fixture()
