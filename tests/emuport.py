"""ctypes binding of tests/emu/libusvmpc_emu.so: the product's CUDA kernel source compiled for the CPU with a
fiber emulation of one thread block (tests/emu/emu_solver.cpp).  Test infrastructure only."""
import ctypes as C
import os
import subprocess

import numpy as np

from refharness import RefProblem, _d, _i

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB = os.path.join(HERE, "emu", "libusvmpc_emu.so")
SRC = [os.path.join(HERE, "emu", "emu_solver.cpp")] + [
    os.path.join(ROOT, "mpc_collisionavoidance_b200", "csrc", f)
    for f in ("cta_kernel.cuh", "models.cuh", "cta_layout.h", "cta_compat.h")]


def build(force=False):
    if not force and os.path.exists(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(s) for s in SRC):
        return LIB
    subprocess.check_call(["g++", "-O2", "-g", "-std=c++17", "-fPIC", "-shared", "-DUSVMPC_EMULATE",
                           "-I" + os.path.join(ROOT, "mpc_collisionavoidance_b200", "csrc"), "-pthread",
                           "-Wno-unknown-pragmas", "-o", LIB, SRC[0]])
    return LIB


def solve_batch(prob: RefProblem, x0, p, lh, yref, yref_e, xinit=None, uinit=None, piinit=None, nthreads=8,
                smem_budget=227 * 1024, block_threads=256, chain_fp32=False):
    lib = C.CDLL(build())
    lib.usvemu_configure(C.c_long(smem_budget), block_threads, int(chain_fp32))
    lib.usvemu_solve_batch.restype = C.c_double
    c = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.float64)
    x0, p, lh, yref, yref_e, xinit, uinit, piinit = map(c, (x0, p, lh, yref, yref_e, xinit, uinit, piinit))
    if p is None:
        p = np.zeros((x0.shape[0], 1))
    if lh is None:
        lh = np.zeros((x0.shape[0], 1))
    B = x0.shape[0]
    N, nx, nu = prob.N, prob.nx, prob.nu
    ncz = len(prob.lbu) + nx + prob.K
    nsh = getattr(prob, "nsh", 0)
    x = np.zeros((B, N + 1, nx)); u = np.zeros((B, N, nu)); pi = np.zeros((B, N, nx))
    wl = 2 * ncz + 2 * nsh
    lam = np.zeros((B, N + 1, wl)); t = np.zeros((B, N + 1, wl)); stats = np.zeros((B, 12))
    sv = np.zeros((B, N, max(2 * nsh, 1)))
    lib.usvemu_slack_output(_d(sv) if nsh else None)
    secs = lib.usvemu_solve_batch(_i(prob.icfg), _d(prob.dcfg), _d(prob.W), _d(prob.We), _d(prob.lbu), _d(prob.ubu),
                                  _i(prob.idxbx), _d(prob.lbx), _d(prob.ubx), B, _d(x0), _d(p), int(p.ndim > 2),
                                  _d(lh), int(lh.ndim > 2), _d(yref), int(yref.ndim > 2), _d(yref_e), _d(xinit),
                                  _d(uinit), _d(piinit), _d(x), _d(u), _d(pi), _d(lam), _d(t), _d(stats), nthreads)
    return dict(x=x, u=u, pi=pi, lam=lam, t=t, status=stats[:, 0].astype(int), sqp_iter=stats[:, 1].astype(int),
                qp_iter=stats[:, 2].astype(int), res=stats[:, 3:7], lq_calls=stats[:, 7].astype(int), solve_calls=stats[:, 8].astype(int),
                itref=stats[:, 11].astype(int), fp32_facts=stats[:, 9].astype(int), seconds=secs,
                sl=sv[:, :, :nsh], su=sv[:, :, nsh:2 * nsh])
