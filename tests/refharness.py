"""ctypes binding of oracle/_ref/libusv_ref.so (the UNMODIFIED reference stack + our hand-rendered
acados_create; see oracle/ref_harness.c).  Test infrastructure only."""
import ctypes as C
import os
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "..", "oracle", "_ref", "libusv_ref.so")

ICFG = ["model", "N", "K", "num_steps", "num_stages", "nlp_type", "max_iter", "qp_iter_max", "cond_N", "nbx", "nbu", "print", "nsh"]
DCFG = ["dt", "tol_stat", "tol_eq", "tol_ineq", "tol_comp", "uh", "lsh", "ush", "zl", "zu", "Zl", "Zu"]

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def available():
    return os.path.exists(LIB)


def _d(a):
    return None if a is None else a.ctypes.data_as(_dp)


def _i(a):
    return None if a is None else a.ctypes.data_as(_ip)


class RefProblem:
    """Problem description shared by the reference harness, the C oracle and the engine."""

    def __init__(self, model=0, N=20, K=3, num_steps=1, num_stages=4, nlp_type=0, max_iter=100, qp_iter_max=50,
                 cond_N=0, dt=0.05, tol=1e-6, uh=1e6, W=None, We=None, lbu=None, ubu=None, idxbx=None, lbx=None,
                 ubx=None, nsh=0, lsh=0.0, ush=0.0, zl=0.0, zu=0.0, Zl=0.0, Zu=0.0):
        self.model, self.N, self.K, self.nsh = model, N, K, nsh
        self.nx, self.nu = {0: (6, 2), 1: (4, 1), 2: (8, 1)}[model]
        self.ny = self.nx + self.nu
        if W is None and model == 2:
            # the deployed CA solver: usv_guidance_ca1/acados_settings.py:70-118
            W = np.diag([0, 0, 0.05, 0.01, 0, 0, 0, 0, 0.2])
            We = np.diag([0, 0, 0.1, 0.05, 0, 0, 0, 0.0])
            lbu, ubu = np.array([-0.5]), np.array([0.5])
            idxbx, lbx, ubx = np.array([], dtype=np.int32), np.array([]), np.array([])
        if W is None:
            W = np.diag([1, 1, 0.1, 10, 0.1, 0.1, 1e-3, 1e-3])
            We = 5 * np.diag([1, 1, 0.1, 10, 0.1, 0.1])
            lbu, ubu = np.array([-30.0, -30.0]), np.array([35.0, 35.0])
            idxbx, lbx, ubx = np.array([3, 4, 5]), np.array([-1.5, -1.5, -1.0]), np.array([1.5, 1.5, 1.0])
        self.W = np.asfortranarray(W, dtype=np.float64)
        self.We = np.asfortranarray(We, dtype=np.float64)
        self.lbu = np.ascontiguousarray(lbu, dtype=np.float64)
        self.ubu = np.ascontiguousarray(ubu, dtype=np.float64)
        self.idxbx = np.ascontiguousarray(idxbx if idxbx is not None else [], dtype=np.int32)
        self.lbx = np.ascontiguousarray(lbx if lbx is not None else [], dtype=np.float64)
        self.ubx = np.ascontiguousarray(ubx if ubx is not None else [], dtype=np.float64)
        self.icfg = np.array([model, N, K, num_steps, num_stages, nlp_type, max_iter, qp_iter_max, cond_N,
                              len(self.idxbx), len(self.lbu), 0, nsh], dtype=np.int32)
        tols = tol if isinstance(tol, (list, tuple)) else [tol] * 4
        self.dcfg = np.array([dt, *tols, uh, lsh, ush, zl, zu, Zl, Zu], dtype=np.float64)
        self.slack = dict(lsh=lsh, ush=ush, zl=zl, zu=zu, Zl=Zl, Zu=Zu)
        self.dt, self.uh = dt, uh
        self.nbm = max(self.nx, len(self.idxbx)) + len(self.lbu) + K   # stage stride /2 of lam, t (+ nsh slack-bound rows)


class RefSolver:
    def __init__(self, prob: RefProblem):
        self.lib = C.CDLL(LIB)
        self.lib.usvref_create.restype = C.c_void_p
        self.lib.usvref_solve_batch.restype = C.c_double
        self.p = prob
        self.h = C.c_void_p(self.lib.usvref_create(_i(prob.icfg), _d(prob.dcfg), _d(prob.W), _d(prob.We), _d(prob.lbu),
                                                   _d(prob.ubu), _i(prob.idxbx), _d(prob.lbx), _d(prob.ubx)))
        assert self.h.value

    def __del__(self):
        if getattr(self, "h", None) is not None and self.h.value:
            self.lib.usvref_free(self.h)
            self.h = None

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
        self.lib.usvref_solve(self.h, _d(x0), _d(p), int(p.ndim > 1), _d(lh), int(lh.ndim > 1), _d(yref),
                              int(yref.ndim > 1), _d(yref_e), _d(xinit), _d(uinit), _d(piinit), _d(x), _d(u), _d(pi),
                              _d(lam), _d(t), _d(stats))
        out = dict(x=x, u=u, pi=pi, lam=lam, t=t, status=int(stats[0]), sqp_iter=int(stats[1]),
                   qp_iter=int(stats[2]), res=stats[3:7].copy(), lq_calls=int(stats[7]), solve_calls=int(stats[8]))
        if P.nsh:
            sl = np.zeros((N, P.nsh)); su = np.zeros((N, P.nsh))
            self.lib.usvref_get_slacks(self.h, _d(sl), _d(su))
            out.update(sl=sl, su=su)
        return out

    def solve_capture_qp(self, want, *args, **kw):
        """Solve while capturing the `want`-th QP HPIPM sees (after x0 elimination)."""
        P = self.p
        N, nvm, nxm = P.N, P.nx + P.nu, P.nx
        ncm = 2 * P.nbm
        buf = dict(BAbt=np.zeros((N, nvm * nxm)), b=np.zeros((N, nxm)), RSQrq=np.zeros((N + 1, nvm * nvm)),
                   rqz=np.zeros((N + 1, nvm)), DCt=np.zeros((N + 1, nvm * max(P.K, 1))), d=np.zeros((N + 1, ncm)),
                   idxb=np.zeros((N + 1, P.nbm), dtype=np.int32), ux=np.zeros((N + 1, nvm)), pi=np.zeros((N, nxm)),
                   lam=np.zeros((N + 1, ncm)), t=np.zeros((N + 1, ncm)))
        b = buf
        self.lib.usvref_tap_arm(want, _d(b["BAbt"]), nvm * nxm, _d(b["b"]), nxm, _d(b["RSQrq"]), nvm * nvm, _d(b["rqz"]),
                                nvm, _d(b["DCt"]), nvm * max(P.K, 1), _d(b["d"]), ncm, _i(b["idxb"]), P.nbm, _d(b["ux"]),
                                nvm, _d(b["pi"]), nxm, _d(b["lam"]), _d(b["t"]), ncm)
        out = self.solve(*args, **kw)
        dims = np.zeros((N + 1, 4), dtype=np.int32); info = np.zeros(3, dtype=np.int32)
        self.lib.usvref_tap_read(_i(dims), _i(info))
        buf.update(dims=dims, got=int(info[0]), iter=int(info[1]), status=int(info[2]))
        return out, buf


def solve_batch(prob: RefProblem, x0, p, lh, yref, yref_e, nthreads=1):
    lib = C.CDLL(LIB)
    lib.usvref_solve_batch.restype = C.c_double
    c = lambda a: np.ascontiguousarray(a, dtype=np.float64)
    x0, p, lh, yref, yref_e = map(c, (x0, p, lh, yref, yref_e))
    B = x0.shape[0]
    x = np.zeros((B, prob.N + 1, prob.nx)); u = np.zeros((B, prob.N, prob.nu)); stats = np.zeros((B, 9))
    secs = lib.usvref_solve_batch(_i(prob.icfg), _d(prob.dcfg), _d(prob.W), _d(prob.We), _d(prob.lbu), _d(prob.ubu),
                                  _i(prob.idxbx), _d(prob.lbx), _d(prob.ubx), B, _d(x0), _d(p), int(p.ndim > 2),
                                  _d(lh), int(lh.ndim > 2), _d(yref), int(yref.ndim > 2), _d(yref_e), _d(x), _d(u),
                                  _d(stats), nthreads)
    return dict(x=x, u=u, status=stats[:, 0].astype(int), sqp_iter=stats[:, 1].astype(int),
                qp_iter=stats[:, 2].astype(int), res=stats[:, 3:7], lq_calls=stats[:, 7].astype(int), solve_calls=stats[:, 8].astype(int), seconds=secs)
