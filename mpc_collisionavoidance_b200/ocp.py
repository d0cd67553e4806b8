"""Problem description for the batched NMPC engine.

These classes carry the SAME attribute names as the reference's description objects
(`AcadosOcp`, `AcadosOcpDims/Cost/Constraints/Options`, `AcadosModel`:
interfaces/acados_template/acados_template/acados_ocp.py, acados_model.py) so that an
`acados_settings.py` of the nmpc_ca scripts can be pointed at this module with its numbers unchanged.
The reference's classes need CasADi (symbolic model) and a code generator; here the model is one of the
engine's built-in device models, selected by `model.name`, and no code is generated.  A genuine
`acados_template.AcadosOcp` instance is accepted as well (duck typing) as long as it stays inside the path
the engine implements: LINEAR_LS cost with Vx=[I;0], Vu=[0;I]; BGH constraints (input boxes, state boxes,
h = obstacle distances, hard or soft: idxsh / lsh / ush with the penalties zl, zu, Zl, Zu); ERK; GAUSS_NEWTON;
PARTIAL_CONDENSING_HPIPM; SQP or SQP_RTI.  Anything outside that path (soft boxes, other cost / constraint /
integrator / QP plugins, a step length != 1, Levenberg-Marquardt, qp_solver_cond_N != N) is refused with an
Exception instead of being silently ignored.
"""
import numpy as np

from . import _lib

# model.name -> device model.  "usv_model_guidance_ca1" is the model the deployed CA node generates its solver from
# (nmpc_ca/scripts/usv_guidance_ca1/usv_model.py:45).
MODELS = {"usv3": 0, "usv_model_ca": 0, "usv3_ca": 0, "pendulum": 1, "pendulum_ode": 1, "usv_model_guidance_ca1": 2,
          "usv_guidance_ca1": 2}
MODEL_DIMS = {0: (6, 2), 1: (4, 1), 2: (8, 1)}


class AcadosModel:
    def __init__(self, name="usv3"):
        self.name = name


class AcadosOcpDims:
    def __init__(self):
        self.N = None
        self.nh = 0  # number of obstacle rows K (np = 2K)


class AcadosOcpCost:
    def __init__(self):
        self.cost_type = "LINEAR_LS"
        self.cost_type_e = "LINEAR_LS"
        self.W = None
        self.W_e = None
        self.Vx = None
        self.Vu = None
        self.Vx_e = None
        self.yref = None
        self.yref_e = None
        # slack penalties of the soft rows (acados_ocp.py: zl, zu, Zl, Zu)
        self.zl = np.array([]); self.zu = np.array([]); self.Zl = np.array([]); self.Zu = np.array([])


class AcadosOcpConstraints:
    def __init__(self):
        self.constr_type = "BGH"
        self.x0 = None
        self.lbu = np.array([]); self.ubu = np.array([]); self.idxbu = np.array([], dtype=int)
        self.lbx = np.array([]); self.ubx = np.array([]); self.idxbx = np.array([], dtype=int)
        self.lh = np.array([]); self.uh = np.array([])
        # soft nonlinear rows: bounds of the lower / upper slacks and the rows they belong to
        self.lsh = np.array([]); self.ush = np.array([]); self.idxsh = np.array([], dtype=int)


class AcadosOcpOptions:
    """defaults of acados_template/acados_ocp.py:1747-1780"""

    def __init__(self):
        self.qp_solver = "PARTIAL_CONDENSING_HPIPM"
        self.hessian_approx = "GAUSS_NEWTON"
        self.integrator_type = "ERK"
        self.tf = None
        self.nlp_solver_type = "SQP_RTI"
        self.nlp_solver_step_length = 1.0
        self.levenberg_marquardt = 0.0
        self.sim_method_num_stages = 4
        self.sim_method_num_steps = 1
        self.qp_solver_iter_max = 50
        self.qp_solver_cond_N = None
        self.nlp_solver_tol_stat = 1e-6
        self.nlp_solver_tol_eq = 1e-6
        self.nlp_solver_tol_ineq = 1e-6
        self.nlp_solver_tol_comp = 1e-6
        self.nlp_solver_max_iter = 100
        self.print_level = 0


class AcadosOcp:
    def __init__(self):
        self.model = AcadosModel()
        self.dims = AcadosOcpDims()
        self.cost = AcadosOcpCost()
        self.constraints = AcadosOcpConstraints()
        self.solver_options = AcadosOcpOptions()
        self.parameter_values = np.array([])


def _arr(a):
    return np.zeros(0) if a is None else np.atleast_1d(np.asarray(a, dtype=np.float64))


def config_from_ocp(ocp):
    """Validate an OCP description against the engine's path and turn it into a `usvmpc_config`."""
    name = getattr(ocp.model, "name", "usv3")
    if name not in MODELS:
        raise Exception(f"model '{name}' is not one of the engine's device models {sorted(MODELS)}")
    model = MODELS[name]
    nx, nu = MODEL_DIMS[model]
    ny = nx + nu
    o, c, k = ocp.solver_options, ocp.cost, ocp.constraints
    if c.cost_type != "LINEAR_LS" or c.cost_type_e != "LINEAR_LS":
        raise Exception("only LINEAR_LS cost is implemented (the cost type of every nmpc_ca script)")
    if getattr(k, "constr_type", "BGH") != "BGH":
        raise Exception("only BGH constraints are implemented")
    if o.qp_solver != "PARTIAL_CONDENSING_HPIPM" or o.hessian_approx != "GAUSS_NEWTON" or o.integrator_type != "ERK":
        raise Exception("the engine implements PARTIAL_CONDENSING_HPIPM (cond_N = N) + GAUSS_NEWTON + ERK")
    if o.nlp_solver_type not in ("SQP", "SQP_RTI"):
        raise Exception(f"unknown nlp_solver_type {o.nlp_solver_type}")
    if o.tf is None or ocp.dims.N is None:
        raise Exception("solver_options.tf and dims.N must be set")
    Vx = np.zeros((ny, nx)); Vx[:nx, :nx] = np.eye(nx)
    Vu = np.zeros((ny, nu)); Vu[nx:, :] = np.eye(nu)
    for given, want, nm in ((c.Vx, Vx, "Vx"), (c.Vu, Vu, "Vu"), (c.Vx_e, np.eye(nx), "Vx_e")):
        if given is not None and not np.array_equal(np.asarray(given, dtype=float), want):
            raise Exception(f"cost.{nm} must be the selector of y = [x; u] used by the nmpc_ca scripts")
    cfg = _lib.Config()
    _lib.check(_lib.load().usvmpc_config_default(cfg, model), "config_default")
    N = int(ocp.dims.N)
    K = int(len(_arr(k.lh)))
    cfg.N, cfg.K = N, K
    cfg.num_steps, cfg.num_stages = int(o.sim_method_num_steps), int(o.sim_method_num_stages)
    cfg.nlp_type = 0 if o.nlp_solver_type == "SQP" else 1
    cfg.max_iter, cfg.qp_iter_max = int(o.nlp_solver_max_iter), int(o.qp_solver_iter_max)
    cfg.dt = float(o.tf) / N
    cfg.tol[:] = [o.nlp_solver_tol_stat, o.nlp_solver_tol_eq, o.nlp_solver_tol_ineq, o.nlp_solver_tol_comp]
    W = np.asarray(c.W, dtype=float); We = np.asarray(c.W_e, dtype=float)
    if W.shape != (ny, ny) or We.shape != (nx, nx):
        raise Exception(f"cost.W must be {ny}x{ny} and cost.W_e {nx}x{nx}")
    cfg.W[:ny * ny] = W.flatten(order="F").tolist()
    cfg.W_e[:nx * nx] = We.flatten(order="F").tolist()
    lbu, ubu, idxbu = _arr(k.lbu), _arr(k.ubu), np.asarray(k.idxbu, dtype=int).ravel()
    if len(lbu) != len(ubu) or len(lbu) != len(idxbu) or len(lbu) > nu or not np.array_equal(idxbu, np.arange(len(lbu))):
        raise Exception("input bounds must be on u[0..nbu) in order (idxbu = 0..nbu-1)")
    lbx, ubx, idxbx = _arr(k.lbx), _arr(k.ubx), np.asarray(k.idxbx, dtype=int).ravel()
    if len(lbx) != len(ubx) or len(lbx) != len(idxbx) or len(lbx) > nx:
        raise Exception("lbx, ubx, idxbx sizes are inconsistent")
    cfg.nbu, cfg.nbx = len(lbu), len(lbx)
    for i in range(len(lbu)):
        cfg.lbu[i], cfg.ubu[i] = lbu[i], ubu[i]
    for i in range(len(lbx)):
        cfg.lbx[i], cfg.ubx[i], cfg.idxbx[i] = lbx[i], ubx[i], int(idxbx[i])
    uh = _arr(k.uh)
    if K:
        if len(uh) != K or not np.all(uh == uh[0]):
            raise Exception("uh must have one (common) value per obstacle row")
        cfg.uh = float(uh[0])
    # ---- soft constraints: only soft h rows, and they must be the leading rows of h (every nmpc_ca script that uses
    # slacks softens all of its h rows: usv_guidance_ca1/acados_settings.py:160-178)
    for nm in ("idxsbx", "idxsbu", "idxsg", "idxsphi", "idxsbx_e", "idxsh_e", "idxsg_e", "idxsphi_e", "lsbx", "lsbu", "usbx", "usbu"):
        v = getattr(k, nm, None)
        if v is not None and np.size(v) > 0:
            raise Exception(f"constraints.{nm}: only soft nonlinear rows (idxsh) are implemented")
    idxsh = np.asarray(getattr(k, "idxsh", []), dtype=int).ravel()
    nsh = len(idxsh)
    if nsh:
        if nsh > K or not np.array_equal(idxsh, np.arange(nsh)):
            raise Exception("idxsh must be 0..nsh-1 (the leading rows of h)")

        def per_row(v, name):
            v = _arr(v)
            if len(v) != nsh:
                raise Exception(f"{name} must have one value per soft row (nsh = {nsh})")
            return v
        lsh, ush = per_row(getattr(k, "lsh", None), "lsh"), per_row(getattr(k, "ush", None), "ush")
        z = [per_row(getattr(c, nm, None), nm)[:nsh] if len(_arr(getattr(c, nm, None))) == nsh
             else per_row(_arr(getattr(c, nm, None))[:nsh], nm) for nm in ("zl", "zu", "Zl", "Zu")]
        cfg.nsh = nsh
        for i in range(nsh):
            cfg.lsh[i], cfg.ush[i] = lsh[i], ush[i]
            cfg.zl[i], cfg.zu[i], cfg.Zl[i], cfg.Zu[i] = z[0][i], z[1][i], z[2][i], z[3][i]
    else:
        for nm in ("zl", "zu", "Zl", "Zu"):
            if np.size(_arr(getattr(c, nm, None))) > 0:
                raise Exception(f"cost.{nm} given but no soft rows (idxsh) declared")
    # ---- options the engine does not implement are refused, not dropped
    if float(getattr(o, "nlp_solver_step_length", 1.0)) != 1.0:
        raise Exception("nlp_solver_step_length != 1: the engine takes full SQP steps like the nmpc_ca scripts")
    if float(getattr(o, "levenberg_marquardt", 0.0)) != 0.0:
        raise Exception("levenberg_marquardt != 0 is not implemented")
    cond_N = getattr(o, "qp_solver_cond_N", None)
    if cond_N is not None and int(cond_N) != N:
        raise Exception("qp_solver_cond_N != N: the engine solves the uncondensed QP (cond_N = N, the default of the template)")
    return cfg, model, nx, nu


# ---------------------------------------------------------------------------------------------------------------
# acados JSON problem description: the on-disk format the reference writes with
# ocp_formulation_json_dump (acados_template/acados_ocp_solver.py:416-444): one dict per description class, attribute
# names without the class prefix, numpy arrays as nested lists; x0 appears as constraints.lbx_0 / ubx_0.
_JSON_FIELDS = {
    "cost": ["cost_type", "cost_type_e", "W", "W_e", "Vx", "Vu", "Vx_e", "yref", "yref_e", "zl", "zu", "Zl", "Zu"],
    "constraints": ["constr_type", "lbu", "ubu", "idxbu", "lbx", "ubx", "idxbx", "lh", "uh", "lsh", "ush", "idxsh",
                    "idxsbx", "idxsbu", "idxsg", "idxsphi", "lsbx", "lsbu", "usbx", "usbu"],
    "solver_options": ["qp_solver", "hessian_approx", "integrator_type", "tf", "nlp_solver_type", "nlp_solver_step_length",
                       "levenberg_marquardt", "qp_solver_cond_N",
                       "sim_method_num_stages", "sim_method_num_steps", "qp_solver_iter_max", "nlp_solver_tol_stat",
                       "nlp_solver_tol_eq", "nlp_solver_tol_ineq", "nlp_solver_tol_comp", "nlp_solver_max_iter", "print_level"],
}


def ocp_to_dict(ocp):
    """AcadosOcp description -> dict in the layout of the reference's acados_ocp_nlp.json"""
    tolist = lambda v: v.tolist() if isinstance(v, np.ndarray) else v
    d = {"model": {"name": ocp.model.name}, "dims": {"N": ocp.dims.N, "nh": int(len(_arr(ocp.constraints.lh)))},
         "parameter_values": tolist(np.asarray(ocp.parameter_values))}
    for sec, names in _JSON_FIELDS.items():
        obj = getattr(ocp, sec)
        d[sec] = {n: tolist(getattr(obj, n, None)) for n in names}
    if ocp.constraints.x0 is not None:
        x0 = np.asarray(ocp.constraints.x0, dtype=float)
        d["constraints"].update(lbx_0=x0.tolist(), ubx_0=x0.tolist(), idxbx_0=list(range(len(x0))))
    return d


def ocp_from_dict(d):
    """dict in the layout of acados_ocp_nlp.json (as written by the reference or by ocp_to_dict) -> AcadosOcp"""
    ocp = AcadosOcp()
    ocp.model.name = d.get("model", {}).get("name", "usv3")
    ocp.dims.N = d["dims"]["N"]
    for sec, names in _JSON_FIELDS.items():
        obj, src = getattr(ocp, sec), d.get(sec, {})
        for n in names:
            if n in src and src[n] is not None:
                v = src[n]
                setattr(obj, n, np.asarray(v) if isinstance(v, list) else v)
    c = d.get("constraints", {})
    if c.get("lbx_0") is not None and len(c["lbx_0"]):
        if c.get("ubx_0") is not None and not np.array_equal(np.asarray(c["lbx_0"]), np.asarray(c["ubx_0"])):
            raise Exception("lbx_0 != ubx_0: only a fixed initial state is implemented")
        ocp.constraints.x0 = np.asarray(c["lbx_0"], dtype=float)
    ocp.parameter_values = np.asarray(d.get("parameter_values", []), dtype=float)
    return ocp


def ocp_formulation_json_dump(acados_ocp, json_file="acados_ocp_nlp.json"):
    import json
    with open(json_file, "w") as f:
        json.dump(ocp_to_dict(acados_ocp), f, indent=4, sort_keys=True)


def ocp_formulation_json_load(json_file="acados_ocp_nlp.json"):
    import json
    with open(json_file) as f:
        return ocp_from_dict(json.load(f))
