"""ctypes binding of libusvmpc.so (include/usvmpc.h).  The engine has no CPU path: if the library is missing
this raises, and if there is no CUDA device usvmpc_create() fails."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("USVMPC_LIB", os.path.join(HERE, "libusvmpc.so"))

ALL_STAGES = -1
EVERY_STAGE = -2
NSTAT = 16


class Config(C.Structure):
    """mirror of `usvmpc_config` (include/usvmpc.h)"""
    _fields_ = [("model", C.c_int), ("N", C.c_int), ("K", C.c_int), ("num_steps", C.c_int), ("num_stages", C.c_int),
                ("nlp_type", C.c_int), ("max_iter", C.c_int), ("qp_iter_max", C.c_int), ("nbx", C.c_int),
                ("nbu", C.c_int), ("idxbx", C.c_int * 8), ("dt", C.c_double), ("tol", C.c_double * 4),
                ("uh", C.c_double), ("lbu", C.c_double * 4), ("ubu", C.c_double * 4), ("lbx", C.c_double * 8),
                ("ubx", C.c_double * 8), ("W", C.c_double * 256), ("W_e", C.c_double * 256), ("nsh", C.c_int),
                ("lsh", C.c_double * 32), ("ush", C.c_double * 32), ("zl", C.c_double * 32), ("zu", C.c_double * 32),
                ("Zl", C.c_double * 32), ("Zu", C.c_double * 32)]


SYMBOLS = ["usvmpc_last_error", "usvmpc_version", "usvmpc_config_default", "usvmpc_create", "usvmpc_free",
           "usvmpc_solve", "usvmpc_update_params", "usvmpc_cost_model_set", "usvmpc_constraints_model_set",
           "usvmpc_out_set", "usvmpc_out_get", "usvmpc_dims_get_from_attr", "usvmpc_get_stats",
           "usvmpc_solver_opts_set", "usvmpc_info", "usvmpc_eval_cost", "usvmpc_obstacle_frontend",
           "usvmpc_set_result_buffer", "usvmpc_qp_solve", "usvmpc_acados_configure", "usvmpc_config_guidance_ca1"]
# include/acados_compat.h: the generated solver's and acados_c's own names for one instance
ACADOS_SYMBOLS = ["acados_create", "acados_update_params", "acados_solve", "acados_free", "acados_print_stats",
                  "acados_get_nlp_in", "acados_get_nlp_out", "acados_get_nlp_solver", "acados_get_nlp_config",
                  "acados_get_nlp_opts", "acados_get_nlp_dims", "acados_get_nlp_plan", "ocp_nlp_cost_model_set",
                  "ocp_nlp_constraints_model_set", "ocp_nlp_out_set", "ocp_nlp_out_get", "ocp_nlp_dims_get_from_attr",
                  "ocp_nlp_get", "ocp_nlp_solver_opts_set", "ocp_nlp_eval_residuals", "ocp_nlp_eval_cost",
                  "ocp_nlp_cost_dims_get_from_attr", "ocp_nlp_constraint_dims_get_from_attr", "ocp_nlp_get_at_stage"]

_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} not found: build it with `python -m mpc_collisionavoidance_b200.build` "
                           "(the engine is CUDA-only; there is no fallback)")
    lib = C.CDLL(LIB_PATH)
    vp, cp, dp, ci = C.c_void_p, C.c_char_p, C.c_void_p, C.c_int
    lib.usvmpc_last_error.restype = cp
    lib.usvmpc_version.restype = cp
    lib.usvmpc_config_default.argtypes = [C.POINTER(Config), ci]
    lib.usvmpc_config_guidance_ca1.argtypes = [C.POINTER(Config)]
    lib.usvmpc_acados_configure.argtypes = [C.POINTER(Config)]
    lib.usvmpc_create.argtypes = [C.POINTER(Config), ci, ci, C.POINTER(vp)]
    lib.usvmpc_free.argtypes = [vp]
    lib.usvmpc_solve.argtypes = [vp, vp]
    lib.usvmpc_update_params.argtypes = [vp, ci, dp, ci, ci, vp]
    lib.usvmpc_cost_model_set.argtypes = [vp, ci, cp, dp, ci, vp]
    lib.usvmpc_constraints_model_set.argtypes = [vp, ci, cp, dp, ci, vp]
    lib.usvmpc_out_set.argtypes = [vp, ci, cp, dp, ci, vp]
    lib.usvmpc_out_get.argtypes = [vp, ci, cp, dp, ci, vp]
    lib.usvmpc_dims_get_from_attr.argtypes = [vp, ci, cp]
    lib.usvmpc_get_stats.argtypes = [vp, dp, ci, vp]
    lib.usvmpc_solver_opts_set.argtypes = [vp, cp, C.c_double]
    lib.usvmpc_eval_cost.argtypes = [vp, dp, ci, vp]
    lib.usvmpc_obstacle_frontend.argtypes = [dp, dp, vp, ci, ci, ci, C.c_double, C.c_double, dp, dp, vp]
    lib.usvmpc_info.argtypes = [vp, cp, C.POINTER(C.c_double)]
    lib.usvmpc_set_result_buffer.argtypes = [vp, dp]
    lib.usvmpc_qp_solve.argtypes = [vp] + [dp] * 9 + [vp]
    for name in SYMBOLS:
        getattr(lib, name)
    _lib = lib
    return lib


def check(rc, what=""):
    if rc < 0:
        raise Exception(f"usvmpc {what}: {load().usvmpc_last_error().decode()} (code {rc})")
    return rc
