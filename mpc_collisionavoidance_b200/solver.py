"""BatchedAcadosOcpSolver -- the reference's `AcadosOcpSolver.{set, get, solve, cost_set, constraints_set,
get_stats, get_residuals, options_set}` surface (interfaces/acados_template/acados_template/acados_ocp_solver.py:
644-1238) for B independent NMPC instances on one B200, over the C ABI of include/usvmpc.h.

Same method names, field strings, stage indexing and error behaviour (Python `Exception` on a bad field, stage or
size); every per-instance value carries a leading batch dimension:

    solver.set(0, "lbx", x0)         x0: [B, nx]           (reference: [nx])
    solver.set(j, "p", pobs)         pobs: [B, 2K]
    status = solver.solve()          -> np.ndarray[B] of acados status codes
    solver.get(1, "x")               -> [B, nx]

Values may be numpy arrays (host; copied inside the call) or torch CUDA tensors (device; ordered on the current
torch stream, no host round trip).  `batch=None` gives the unbatched drop-in: 1-D values, `solve()` returns an int.
Extensions for batched use: stage = "all" ([B, n_stages, dim]) and stage = "every" (one [B, dim] row for every
stage), `get_all`, `device=True` getters returning torch tensors.
"""
import ctypes as C

import numpy as np

from . import _lib
from .ocp import config_from_ocp

_OUT_FIELDS = ["x", "u", "z", "pi", "lam", "t"]
_MEM_FIELDS = ["sl", "su"]
_COST_FIELDS = ["y_ref", "yref"]
_CONSTR_FIELDS = ["lbx", "ubx", "lbu", "ubu"]
_STAT_COL = {"status": 0, "sqp_iter": 1, "qp_iter": 2, "res_stat": 3, "res_eq": 4, "res_ineq": 5, "res_comp": 6,
             "lq_fact": 7, "solve_sweeps": 8, "qp_status": 9, "qp_iter_last": 10, "refinement_solves": 11,
             "time_tot": 12, "time_lin": 13, "time_qp": 14, "fp32_factorisations": 15}


def _is_torch(v):
    return type(v).__module__.startswith("torch")


class BatchedAcadosOcpSolver:
    def __init__(self, acados_ocp, batch=None, device=0, json_file=None):
        import torch  # device memory, streams: plumbing only
        self._torch = torch
        self.acados_ocp = acados_ocp
        self.unbatched = batch is None
        self.B = 1 if batch is None else int(batch)
        self.device = int(device)
        self.lib = _lib.load()
        self.cfg, self.model, self.nx, self.nu = config_from_ocp(acados_ocp)
        self.N, self.K = self.cfg.N, self.cfg.K
        h = C.c_void_p()
        _lib.check(self.lib.usvmpc_create(C.byref(self.cfg), self.B, self.device, C.byref(h)), "create")
        self.h = h
        self.status = np.zeros(self.B, dtype=np.int64)
        self.sync_host_sets = True  # False: caller keeps (pinned) host buffers alive until it synchronises
        k, c = acados_ocp.constraints, acados_ocp.cost
        # initial values the generated acados_create() would bake in (acados_solver.in.c:796-1449, 1595-1623)
        unbatched, self.unbatched = self.unbatched, False   # the initial values below are written batched
        if k.x0 is not None:
            x0 = np.tile(np.asarray(k.x0, dtype=np.float64), (self.B, 1))
            self._call_set(0, "lbx", x0)
            self._call_set("all", "x", np.repeat(x0[:, None, :], self.N + 1, axis=1))
        if c.yref is not None:
            self._call_set("every", "yref", np.tile(np.asarray(c.yref, dtype=np.float64), (self.B, 1)))
        if c.yref_e is not None:
            self._call_set(self.N, "yref", np.tile(np.asarray(c.yref_e, dtype=np.float64), (self.B, 1)))
        if self.K:
            self._call_set("every", "lh", np.tile(np.asarray(k.lh, dtype=np.float64), (self.B, 1)))
            pv = np.asarray(acados_ocp.parameter_values, dtype=np.float64).ravel()
            if pv.size == 2 * self.K:
                self._call_set("every", "p", np.tile(pv, (self.B, 1)))
        self.unbatched = unbatched

    # ------------------------------------------------------------------ plumbing
    def __del__(self):
        h = getattr(self, "h", None)
        if h is not None and h.value:
            self.lib.usvmpc_free(h)
            self.h = None

    def _stream(self):
        return C.c_void_p(self._torch.cuda.current_stream(self.device).cuda_stream)

    def _stage(self, stage_, allow_special=True):
        if isinstance(stage_, str):
            if allow_special and stage_ in ("all", "every"):
                return _lib.ALL_STAGES if stage_ == "all" else _lib.EVERY_STAGE
            raise Exception("stage index must be Integer.")
        if not isinstance(stage_, (int, np.integer)):
            raise Exception("stage index must be Integer.")
        if stage_ < 0 or stage_ > self.N:
            raise Exception("stage index must be in [0, N], got: {}.".format(stage_))
        return int(stage_)

    def _dims(self, stage, field):
        d = self.lib.usvmpc_dims_get_from_attr(self.h, max(stage, 0), field.encode())
        if d < 0:
            raise Exception(self.lib.usvmpc_last_error().decode())
        return d

    def _nstages(self, field):
        return {"x": self.N + 1, "u": self.N, "pi": self.N, "p": self.N + 1, "lh": self.N, "yref": self.N,
                "y_ref": self.N}.get(field)

    def _value(self, value_, stage, field):
        """-> (pointer, on_device, keepalive); checks the size like AcadosOcpSolver.set (:974-984)"""
        dims = self._dims(stage, field)
        if _is_torch(value_):
            v = value_.to(self._torch.float64).contiguous()
        else:
            v = np.ascontiguousarray(value_, dtype=np.float64)
        if self.unbatched:
            if v.ndim != 1:
                raise Exception("unbatched solver: values are 1-D")
            v = v[None]
        want = (self.B, self._nstages(field), dims) if stage == _lib.ALL_STAGES else (self.B, dims)
        if tuple(v.shape) != want:
            msg = 'mismatching dimension for field "{}" '.format(field)
            msg += "with dimension {} (you have {})".format(want, tuple(v.shape))
            raise Exception(msg)
        if _is_torch(v):
            if v.is_cuda:
                return C.c_void_p(v.data_ptr()), 1, v
            return C.c_void_p(v.data_ptr()), 0, v
        return C.c_void_p(v.ctypes.data), 0, v

    def _call_set(self, stage_, field_, value_):
        stage = self._stage(stage_)
        lib, h, st = self.lib, self.h, self._stream()
        dev = 1
        if field_ == "p":
            ptr, dev, keep = self._value(value_, stage, "p")
            _lib.check(lib.usvmpc_update_params(h, stage, ptr, 2 * self.K, dev, st), "set p")
        elif field_ in _CONSTR_FIELDS + ["lh"]:
            if (stage > 0 and field_ in _CONSTR_FIELDS) or field_ in ("lbu", "ubu"):
                if hasattr(value_, "detach"):       # a torch tensor (possibly on the device): these few numbers go by the host
                    value_ = value_.detach().cpu().numpy()
                v = np.ascontiguousarray(value_, dtype=np.float64).ravel()  # shared by the batch, copied by the call
                if v.shape[0] != self._dims(stage, field_):
                    raise Exception('mismatching dimension for field "{}" with dimension {} (you have {})'.format(
                        field_, self._dims(stage, field_), v.shape[0]))
                _lib.check(lib.usvmpc_constraints_model_set(h, stage, field_.encode(), C.c_void_p(v.ctypes.data), 0, st), "set")
            else:
                ptr, dev, keep = self._value(value_, stage, field_)
                _lib.check(lib.usvmpc_constraints_model_set(h, stage, field_.encode(), ptr, dev, st), "set " + field_)
        elif field_ in _COST_FIELDS:
            ptr, dev, keep = self._value(value_, stage, field_)
            _lib.check(lib.usvmpc_cost_model_set(h, stage, field_.encode(), ptr, dev, st), "set " + field_)
        elif field_ in ["x", "u", "pi", "lam", "t"]:
            ptr, dev, keep = self._value(value_, stage, field_)
            _lib.check(lib.usvmpc_out_set(h, stage, field_.encode(), ptr, dev, st), "set " + field_)
        else:
            raise Exception("AcadosOcpSolver.set(): {} is not a valid argument.\nPossible values are {}. Exiting.".format(
                field_, _CONSTR_FIELDS + _COST_FIELDS + ["x", "u", "pi", "lam", "t", "p"]))
        if not dev and self.sync_host_sets:
            # host values: the reference copies before returning; wait until the asynchronous copy has consumed them
            self._torch.cuda.current_stream(self.device).synchronize()

    # ------------------------------------------------------------------ the reference's surface
    def solve(self):
        """solve all instances with the current inputs; returns the acados status per instance"""
        self.solve_async()
        st = self.get_stats("status")
        self.status = st
        return int(st[0]) if self.unbatched else st

    def solve_async(self):
        """launch the solve on the current torch stream and return immediately"""
        _lib.check(self.lib.usvmpc_solve(self.h, self._stream()), "solve")

    def set(self, stage_, field_, value_):
        self._call_set(stage_, field_, value_)

    def cost_set(self, stage_, field_, value_, api="warn"):
        if field_ == "W":
            stage = self._stage(stage_, allow_special=False)
            n = self.nx if stage == self.N else self.nx + self.nu
            v = np.asarray(value_, dtype=np.float64)
            if v.shape != (n, n):
                raise Exception("AcadosOcpSolver.cost_set(): mismatching dimension for field W: {} vs {}".format(v.shape, (n, n)))
            v = np.ascontiguousarray(v.flatten(order="F"))  # column-major like the reference (:1032-1062)
            _lib.check(self.lib.usvmpc_cost_model_set(self.h, stage, b"W", C.c_void_p(v.ctypes.data), 0, self._stream()), "cost_set W")
            return
        if field_ in ("zl", "zu", "Zl", "Zu"):
            # slack penalties: shared by the batch and the stages
            v = np.ascontiguousarray(value_, dtype=np.float64).ravel()
            if v.shape[0] != self.cfg.nsh:
                raise Exception("AcadosOcpSolver.cost_set(): mismatching dimension for field {}: {} vs {}".format(field_, v.shape[0], self.cfg.nsh))
            _lib.check(self.lib.usvmpc_cost_model_set(self.h, self._stage(stage_, allow_special=False), field_.encode(),
                                                      C.c_void_p(v.ctypes.data), 0, self._stream()), "cost_set " + field_)
            return
        if field_ not in _COST_FIELDS:
            raise Exception("AcadosOcpSolver.cost_set(): {} is not a valid argument (yref, y_ref, W, zl, zu, Zl, Zu)".format(field_))
        self._call_set(stage_, field_, value_)

    def constraints_set(self, stage_, field_, value_, api="warn"):
        if field_ in ("uh", "lsh", "ush"):
            v = np.ascontiguousarray(value_, dtype=np.float64).ravel()
            want = self.K if field_ == "uh" else self.cfg.nsh
            if v.shape[0] != want:
                raise Exception('mismatching dimension for field "{}" with dimension {} (you have {})'.format(field_, want, v.shape[0]))
            _lib.check(self.lib.usvmpc_constraints_model_set(self.h, self._stage(stage_), field_.encode(), C.c_void_p(v.ctypes.data), 0, self._stream()), field_)
            return
        if field_ not in _CONSTR_FIELDS + ["lh"]:
            raise Exception("AcadosOcpSolver.constraints_set(): {} is not a valid argument".format(field_))
        self._call_set(stage_, field_, value_)

    def get(self, stage_, field_, device=False):
        if field_ not in _OUT_FIELDS + _MEM_FIELDS:
            raise Exception("AcadosOcpSolver.get(): {} is an invalid argument.\n Possible values are {}. Exiting.".format(
                field_, _OUT_FIELDS + _MEM_FIELDS))
        if not isinstance(stage_, (int, np.integer)):
            raise Exception("AcadosOcpSolver.get(): stage index must be Integer.")
        if stage_ < 0 or stage_ > self.N:
            raise Exception("AcadosOcpSolver.get(): stage index must be in [0, N], got: {}.".format(stage_))
        if stage_ == self.N and field_ == "pi":
            raise Exception("AcadosOcpSolver.get(): field {} does not exist at final stage {}.".format(field_, stage_))
        dims = self._dims(int(stage_), field_)
        return self._get(int(stage_), field_, (self.B, dims), device)

    def get_all(self, field_, device=False):
        """[B, n_stages, dim] of x, u or pi in one call"""
        if field_ not in ("x", "u", "pi"):
            raise Exception("get_all(): field must be x, u or pi")
        return self._get(_lib.ALL_STAGES, field_, (self.B, self._nstages(field_), self._dims(0, field_)), device)

    def _get(self, stage, field, shape, device):
        st = self._stream()
        if device:
            out = self._torch.empty(shape, dtype=self._torch.float64, device=f"cuda:{self.device}")
            if out.numel():
                _lib.check(self.lib.usvmpc_out_get(self.h, stage, field.encode(), C.c_void_p(out.data_ptr()), 1, st), "get")
            return out[0] if self.unbatched else out
        out = np.zeros(shape, dtype=np.float64)
        if out.size:
            _lib.check(self.lib.usvmpc_out_get(self.h, stage, field.encode(), C.c_void_p(out.ctypes.data), 0, st), "get")
        return out[0] if self.unbatched else out

    def stats_table(self, device=False):
        """[B, 16] statistics record (include/usvmpc.h)"""
        if device:
            out = self._torch.empty((self.B, _lib.NSTAT), dtype=self._torch.float64, device=f"cuda:{self.device}")
            _lib.check(self.lib.usvmpc_get_stats(self.h, C.c_void_p(out.data_ptr()), 1, self._stream()), "stats")
            return out
        out = np.zeros((self.B, _lib.NSTAT))
        _lib.check(self.lib.usvmpc_get_stats(self.h, C.c_void_p(out.ctypes.data), 0, self._stream()), "stats")
        return out

    def get_stats(self, field_):
        """fields of the reference (:816-865) that the engine tracks: sqp_iter, qp_iter, status, res_*"""
        if field_ == "statistics":
            return self.stats_table()
        if field_ not in _STAT_COL:
            raise Exception("AcadosOcpSolver.get_stats(): {} is not a valid argument.\n Possible values are {}. Exiting.".format(
                field_, list(_STAT_COL) + ["statistics"]))
        col = self.stats_table()[:, _STAT_COL[field_]]
        if field_.startswith("time_"):
            # the reference's timers (ocp_nlp_sqp_rti.c:984-1014) are wall-clock seconds; the engine counts SM clock cycles
            # per instance inside the kernel (time one block spent on the instance)
            return col / self.info("sm_clock_hz")
        if not field_.startswith("res_"):
            col = col.astype(np.int64)
        return col

    def get_residuals(self):
        """[res_stat, res_eq, res_ineq, res_comp] per instance ([B, 4]; [4] unbatched)"""
        r = self.stats_table()[:, 3:7]
        return r[0] if self.unbatched else r

    def get_cost(self):
        """cost value of the current solution per instance ([B]; float unbatched), like AcadosOcpSolver.get_cost (:878)"""
        out = np.zeros(self.B)
        _lib.check(self.lib.usvmpc_eval_cost(self.h, C.c_void_p(out.ctypes.data), 0, self._stream()), "get_cost")
        return float(out[0]) if self.unbatched else out

    def options_set(self, field_, value_):
        if field_ == "globalization":
            if value_ != "fixed_step":
                raise Exception("only globalization = fixed_step is implemented")
            return
        if field_ == "initialize_t_slacks":
            if value_:
                raise Exception("initialize_t_slacks = 1 is not implemented")
            return
        _lib.check(self.lib.usvmpc_solver_opts_set(self.h, field_.encode(), float(value_)), "options_set")

    def qp_solve(self, G, b, rq, gxy, d):
        """The QP seam of the reference (qp_solver_config.evaluate, ocp_qp_common.h:62-76): solve one given OCP QP per
        instance (engine layout, see include/usvmpc.h:usvmpc_qp_solve) with the engine's IPM.  Arrays: numpy or torch;
        returns dict(ux, pi, lam, t, iter, status, res) of numpy arrays."""
        torch = self._torch
        dev = f"cuda:{self.device}"
        nv, nx, N, B, K = self.nx + self.nu, self.nx, self.N, self.B, self.K
        r2 = 2 * (self.cfg.nbu + self.cfg.nbx + K)
        want = {"G": (B, N, nv * nx), "b": (B, N, nx), "rq": (B, N + 1, nv), "gxy": (B, N, 2 * K), "d": (B, N, r2)}
        t = {}
        for name, v in (("G", G), ("b", b), ("rq", rq), ("gxy", gxy), ("d", d)):
            a = torch.as_tensor(np.asarray(v) if not torch.is_tensor(v) else v, dtype=torch.float64, device=dev).contiguous()
            if tuple(a.shape) != want[name]:
                raise Exception('mismatching dimension for field "{}" with dimension {} (you have {})'.format(name, want[name], tuple(a.shape)))
            t[name] = a
        out = {"ux": torch.zeros((B, N + 1, nv), dtype=torch.float64, device=dev),
               "pi": torch.zeros((B, N, nx), dtype=torch.float64, device=dev),
               "lam": torch.zeros((B, N, r2), dtype=torch.float64, device=dev),
               "t": torch.zeros((B, N, r2), dtype=torch.float64, device=dev)}
        p = lambda a: C.c_void_p(a.data_ptr()) if a.numel() else None
        _lib.check(self.lib.usvmpc_qp_solve(self.h, p(t["G"]), p(t["b"]), p(t["rq"]), p(t["gxy"]), p(t["d"]), p(out["ux"]),
                                            p(out["pi"]), p(out["lam"]), p(out["t"]), self._stream()), "qp_solve")
        st = self.stats_table()
        res = {k: v.cpu().numpy() for k, v in out.items()}
        res.update(iter=st[:, 2].astype(np.int64), status=st[:, 0].astype(np.int64), res=st[:, 3:7])
        return res

    def set_result_buffer(self, tensor):
        """torch CUDA float64 tensor [B, width] (or None): the solve kernel writes each instance's packed result row
        (x | u | status, sqp_iter, qp_iter, 4 residuals) into it -- the send buffer of the multi-GPU all-gather."""
        if tensor is None:
            self._result_buffer = None
            return _lib.check(self.lib.usvmpc_set_result_buffer(self.h, None), "set_result_buffer")
        width = (self.N + 1) * self.nx + self.N * self.nu + 7
        if tuple(tensor.shape) != (self.B, width) or not tensor.is_cuda or not tensor.is_contiguous() \
                or tensor.dtype != self._torch.float64:
            raise Exception("result buffer must be a contiguous CUDA float64 tensor of shape {}".format((self.B, width)))
        self._result_buffer = tensor   # keep it alive
        return _lib.check(self.lib.usvmpc_set_result_buffer(self.h, C.c_void_p(tensor.data_ptr())), "set_result_buffer")

    def info(self, what):
        v = C.c_double()
        _lib.check(self.lib.usvmpc_info(self.h, what.encode(), C.byref(v)), "info")
        return v.value

    def print_statistics(self):
        t = self.stats_table()
        print("inst\tstatus\tsqp_iter\tqp_iter\tres_stat\tres_eq\t\tres_ineq\tres_comp")
        for i, r in enumerate(t[:16]):
            print("{:d}\t{:d}\t{:d}\t\t{:d}\t{:e}\t{:e}\t{:e}\t{:e}".format(i, int(r[0]), int(r[1]), int(r[2]), *r[3:7]))
