"""ctypes binding of oracle/libusv_oracle.so (the plain-C restatement, oracle/usv_oracle.c).
Test infrastructure only: the product never imports this."""
import ctypes as C
import os
import numpy as np

from refharness import RefProblem, _d, _i  # noqa: F401  (same problem description as the reference harness)

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "..", "oracle", "libusv_oracle.so")


class usvo_problem(C.Structure):
    _fields_ = [("raw", C.c_char * 16384)]  # opaque, larger than sizeof(usvo_problem)


def available():
    return os.path.exists(LIB)


class OracleSolver:
    def __init__(self, prob: RefProblem):
        self.lib = C.CDLL(LIB)
        self.lib.usvo_solve_batch.restype = C.c_double
        self.p = prob
        self.P = usvo_problem()
        self.lib.usvo_problem_init(C.byref(self.P), _i(prob.icfg), _d(prob.dcfg), _d(prob.W), _d(prob.We), _d(prob.lbu),
                                   _d(prob.ubu), _i(prob.idxbx), _d(prob.lbx), _d(prob.ubx))

    def integrate(self, x, u):
        P = self.p
        x = np.ascontiguousarray(x, dtype=np.float64); u = np.ascontiguousarray(u, dtype=np.float64)
        xn = np.zeros(P.nx); A = np.zeros((P.nx, P.nx), order="F"); B = np.zeros((P.nx, P.nu), order="F")
        self.lib.usvo_integrate(C.byref(self.P), _d(x), _d(u), _d(xn), _d(A), _d(B))
        return xn, A, B

    def solve(self, x0, p, lh, yref, yref_e, xinit=None, uinit=None, piinit=None):
        P = self.p
        N, nx, nu = P.N, P.nx, P.nu
        c = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.float64)
        x0, p, lh, yref, yref_e, xinit, uinit, piinit = map(c, (x0, p, lh, yref, yref_e, xinit, uinit, piinit))
        if p is None:
            p = np.zeros(1)
        if lh is None:
            lh = np.zeros(1)
        x = np.zeros((N + 1, nx)); u = np.zeros((N, nu)); pi = np.zeros((N, nx))
        w = 2 * P.nbm + 2 * P.nsh
        lam = np.zeros((N + 1, w)); t = np.zeros((N + 1, w)); stats = np.zeros(9)
        sl = np.zeros((N, max(P.nsh, 1))); su = np.zeros((N, max(P.nsh, 1)))
        self.lib.usvo_solve_ex(C.byref(self.P), _d(x0), _d(p), int(p.ndim > 1), _d(lh), int(lh.ndim > 1), _d(yref),
                               int(yref.ndim > 1), _d(yref_e), _d(xinit), _d(uinit), _d(piinit), _d(x), _d(u), _d(pi),
                               _d(lam), _d(t), _d(stats), _d(sl), _d(su))
        out = dict(x=x, u=u, pi=pi, lam=lam, t=t, status=int(stats[0]), sqp_iter=int(stats[1]),
                   qp_iter=int(stats[2]), res=stats[3:7].copy(), lq_calls=int(stats[7]), solve_calls=int(stats[8]))
        if P.nsh:
            out.update(sl=sl, su=su)
        return out


def qp_solve(buf, sqp_mode=True, tol4=(1e-6,) * 4, iter_max=50):
    """Solve a QP captured by RefSolver.solve_capture_qp with the oracle's IPM."""
    lib = C.CDLL(LIB)
    dims = np.ascontiguousarray(buf["dims"], dtype=np.int32)
    N = dims.shape[0] - 1
    out = {k: np.zeros_like(buf[k]) for k in ("ux", "pi", "lam", "t")}
    info = np.zeros(8)
    tol = np.array(tol4, dtype=np.float64)
    g = lambda k: buf[k].shape[1]
    lib.usvo_qp_solve_flat(N, _i(dims), _d(buf["BAbt"]), g("BAbt"), _d(buf["b"]), g("b"), _d(buf["RSQrq"]), g("RSQrq"),
                           _d(buf["rqz"]), g("rqz"), _d(buf["DCt"]), g("DCt"), _d(buf["d"]), g("d"), _i(buf["idxb"]),
                           g("idxb"), int(sqp_mode), _d(tol), iter_max, _d(out["ux"]), g("ux"), _d(out["pi"]), g("pi"),
                           _d(out["lam"]), _d(out["t"]), g("lam"), _d(info))
    out.update(iter=int(info[0]), status=int(info[1]), solve_calls=int(info[2]), lq_wanted=int(info[3]), res=info[4:8])
    return out


def solve_batch(prob: RefProblem, x0, p, lh, yref, yref_e, nthreads=1):
    lib = C.CDLL(LIB)
    lib.usvo_solve_batch.restype = C.c_double
    c = lambda a: np.ascontiguousarray(a, dtype=np.float64)
    x0, p, lh, yref, yref_e = map(c, (x0, p, lh, yref, yref_e))
    B = x0.shape[0]
    x = np.zeros((B, prob.N + 1, prob.nx)); u = np.zeros((B, prob.N, prob.nu)); stats = np.zeros((B, 9))
    secs = lib.usvo_solve_batch(_i(prob.icfg), _d(prob.dcfg), _d(prob.W), _d(prob.We), _d(prob.lbu), _d(prob.ubu),
                                _i(prob.idxbx), _d(prob.lbx), _d(prob.ubx), B, _d(x0), _d(p), int(p.ndim > 2),
                                _d(lh), int(lh.ndim > 2), _d(yref), int(yref.ndim > 2), _d(yref_e), _d(x), _d(u),
                                _d(stats), nthreads)
    return dict(x=x, u=u, status=stats[:, 0].astype(int), sqp_iter=stats[:, 1].astype(int),
                qp_iter=stats[:, 2].astype(int), res=stats[:, 3:7], lq_calls=stats[:, 7].astype(int),
                solve_calls=stats[:, 8].astype(int), seconds=secs)
