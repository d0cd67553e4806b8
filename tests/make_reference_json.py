"""Write the acados JSON description of the deployed CA solver (usv_guidance_ca1) with the REFERENCE's own
ocp_formulation_json_dump, to pin the engine's JSON loader on a file the reference itself wrote.  casadi and
future_fstrings are not installed here: a stub module / codec stand in for them (no symbolic model is needed to
dump the description); the dimensions make_ocp_dims_consistent would derive are filled in by hand."""
import codecs, json, os, sys, types
import numpy as np, scipy.linalg

def _search(name):
    if name.replace('-', '_') == 'future_fstrings':
        u = codecs.lookup('utf-8')
        return codecs.CodecInfo(name='future_fstrings', encode=u.encode, decode=u.decode, incrementalencoder=u.incrementalencoder,
                                incrementaldecoder=u.incrementaldecoder, streamreader=u.streamreader, streamwriter=u.streamwriter)
codecs.register(_search)
stub = types.ModuleType('casadi')
class _Sym:
    def __init__(self, *a, **k): pass
for n in ('SX', 'MX', 'DM', 'Function'):
    setattr(stub, n, type(n, (_Sym,), {}))
stub.CasadiMeta = type('CasadiMeta', (), {'version': staticmethod(lambda: '3.5.1')})
for n in ('transpose', 'vertcat', 'horzcat', 'jacobian'):
    setattr(stub, n, lambda *a: None)
stub.__all__ = ['SX', 'MX', 'DM', 'Function', 'CasadiMeta', 'transpose', 'vertcat', 'horzcat', 'jacobian']
sys.modules['casadi'] = stub
sys.path.insert(0, '/root/reference/catkin_ws/src/nmpc_ca/acados/interfaces/acados_template')
from acados_template import AcadosOcp, AcadosModel
from acados_template.acados_ocp_solver import ocp_formulation_json_dump

def guidance_ca1_ocp():
    # numbers of /root/reference/catkin_ws/src/nmpc_ca/scripts/usv_guidance_ca1/acados_settings.py:42-208, main.py:54-55
    ocp = AcadosOcp()
    m = AcadosModel(); m.name = "usv_model_guidance_ca1"
    ocp.model = m
    nx, nu = 8, 1
    ny, ny_e = nx + nu, nx
    N, Tf = 100, 5.0
    ocp.dims.N = N
    ns = 8
    Q = np.diag([0, 0, 0.05, 0.01, 0, 0, 0, 0]); R = np.eye(nu); R[0, 0] = 0.2; Qe = np.diag([0, 0, 0.1, 0.05, 0, 0, 0, 0])
    ocp.cost.cost_type = "LINEAR_LS"; ocp.cost.cost_type_e = "LINEAR_LS"
    ocp.cost.W = scipy.linalg.block_diag(Q, R); ocp.cost.W_e = Qe
    Vx = np.zeros((ny, nx)); Vx[:nx, :nx] = np.eye(nx); ocp.cost.Vx = Vx
    Vu = np.zeros((ny, nu)); Vu[8, 0] = 1.0; ocp.cost.Vu = Vu
    Vx_e = np.zeros((ny_e, nx)); Vx_e[:nx, :nx] = np.eye(nx); ocp.cost.Vx_e = Vx_e
    ocp.cost.zl = 1 * np.ones((ns,)); ocp.cost.Zl = 0 * np.ones((ns,)); ocp.cost.zu = 1 * np.ones((ns,)); ocp.cost.Zu = 0 * np.ones((ns,))
    ocp.cost.yref = np.zeros(9); ocp.cost.yref_e = np.zeros(8)
    ocp.constraints.lbu = np.array([-0.5]); ocp.constraints.ubu = np.array([0.5]); ocp.constraints.idxbu = np.array([0])
    ocp.constraints.lh = np.full(8, 1.5); ocp.constraints.uh = np.full(8, 1000000.0)
    ocp.constraints.lsh = np.full(8, -0.2); ocp.constraints.ush = np.zeros(8); ocp.constraints.idxsh = np.arange(8)
    ocp.constraints.x0 = np.zeros(8)
    ocp.parameter_values = np.full(16, 100.0)
    o = ocp.solver_options
    o.tf = Tf; o.qp_solver = "PARTIAL_CONDENSING_HPIPM"; o.nlp_solver_type = "SQP_RTI"; o.hessian_approx = "GAUSS_NEWTON"; o.integrator_type = "ERK"
    d = ocp.dims
    d.nx, d.nu, d.np, d.nz = nx, nu, 16, 0
    d.ny, d.ny_e = ny, ny_e
    d.nbu, d.nbx, d.nbx_0, d.nbx_e, d.nbxe_0 = 1, 0, 8, 0, 8
    d.nh, d.nh_e, d.nsh, d.ns = 8, 0, 8, 8
    o.time_steps = np.full(N, Tf / N)
    return ocp

def benchmark_usv3_ocp():
    # the benchmark OCP of SURVEY.md section 8d, config 1 (mpc_collisionavoidance_b200/workloads.py:benchmark_ocp), written
    # into the REFERENCE's description classes
    ocp = AcadosOcp()
    m = AcadosModel(); m.name = "usv3"
    ocp.model = m
    nx, nu, N, K = 6, 2, 20, 3
    ocp.dims.N = N
    Q = np.diag([1, 1, 0.1, 10, 0.1, 0.1]); R = np.diag([1e-3, 1e-3])
    ocp.cost.cost_type = "LINEAR_LS"; ocp.cost.cost_type_e = "LINEAR_LS"
    ocp.cost.W = scipy.linalg.block_diag(Q, R); ocp.cost.W_e = 5 * Q
    Vx = np.zeros((nx + nu, nx)); Vx[:nx, :nx] = np.eye(nx); ocp.cost.Vx = Vx
    Vu = np.zeros((nx + nu, nu)); Vu[nx:, :] = np.eye(nu); ocp.cost.Vu = Vu
    ocp.cost.Vx_e = np.eye(nx)
    ocp.cost.yref = np.array([6, 0, 0, 1, 0, 0, 0, 0.0]); ocp.cost.yref_e = np.array([6, 0, 0, 1, 0, 0.0])
    c = ocp.constraints
    c.lbu = np.array([-30.0, -30.0]); c.ubu = np.array([35.0, 35.0]); c.idxbu = np.array([0, 1])
    c.lbx = np.array([-1.5, -1.5, -1.0]); c.ubx = np.array([1.5, 1.5, 1.0]); c.idxbx = np.array([3, 4, 5])
    c.lh = np.full(K, 0.8); c.uh = np.full(K, 1e6)
    c.x0 = np.array([0, 0, 0, 0.7, 0, 0.0])
    ocp.parameter_values = np.array([2.0, 0.2, 3.5, -0.6, 5.0, 0.5])
    o = ocp.solver_options
    o.tf = 0.05 * N; o.qp_solver = "PARTIAL_CONDENSING_HPIPM"; o.nlp_solver_type = "SQP"; o.hessian_approx = "GAUSS_NEWTON"
    o.integrator_type = "ERK"; o.sim_method_num_stages = 4; o.sim_method_num_steps = 1
    d = ocp.dims
    d.nx, d.nu, d.np, d.nz = nx, nu, 2 * K, 0
    d.ny, d.ny_e = nx + nu, nx
    d.nbu, d.nbx, d.nbx_0, d.nbx_e, d.nbxe_0 = 2, 3, 6, 0, 6
    d.nh, d.nh_e, d.nsh, d.ns = K, 0, 0, 0
    o.time_steps = np.full(N, 0.05)
    return ocp


if __name__ == '__main__':
    G = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
    for name, ocp in (("acados_ocp_usv_guidance_ca1.json", guidance_ca1_ocp()), ("acados_ocp_usv3_cfg1.json", benchmark_usv3_ocp())):
        out = os.path.join(G, name)
        ocp_formulation_json_dump(ocp, out)
        j = json.load(open(out))
        j["acados_include_path"] = j["acados_lib_path"] = ""   # machine paths of the writer: not part of the problem
        json.dump(j, open(out, "w"), indent=4, sort_keys=True)
        print(name, j['dims'])
