"""CPU-only checks of the drop-in boundary: the C-ABI library loads, exports every symbol include/usvmpc.h declares,
the description -> config translation mirrors the reference's field names, and the engine refuses loudly to run
without a CUDA device (no CPU fallback).  No compute calls."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import refharness as rh
from enginehelper import ocp_from_problem
from mpc_collisionavoidance_b200 import _lib, build
from mpc_collisionavoidance_b200.ocp import config_from_ocp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    build.build()
    return _lib.load()


def test_library_exports_every_declared_symbol(lib):
    hdr = ""
    for h in sorted(os.listdir(os.path.join(ROOT, "include"))):
        hdr += open(os.path.join(ROOT, "include", h)).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = sorted(set(re.findall(r"\b(usvmpc_[a-z_0-9]+)\s*\(", hdr)))
    assert len(declared) >= 21
    for name in declared:
        assert hasattr(lib, name), name
    assert sorted(_lib.SYMBOLS) == declared
    # include/acados_compat.h: the reference's own symbol names (acados_solver.in.h:44-56, ocp_nlp_interface.h)
    theirs = sorted(set(re.findall(r"\b((?:acados|ocp_nlp)_[a-z_]+)\s*\(", hdr)))
    assert len(theirs) == 24 and theirs == sorted(_lib.ACADOS_SYMBOLS)
    for name in theirs:
        assert hasattr(lib, name), name


def test_deployed_solver_description(lib):
    # what acados_create() builds by default == the description of the deployed node (workloads.guidance_ca1_ocp)
    from mpc_collisionavoidance_b200.ocp import config_from_ocp
    from mpc_collisionavoidance_b200.workloads import guidance_ca1_ocp
    a = _lib.Config()
    assert lib.usvmpc_config_guidance_ca1(C.byref(a)) == 0
    b = config_from_ocp(guidance_ca1_ocp())[0]
    for name, _ in _lib.Config._fields_:
        va, vb = getattr(a, name), getattr(b, name)
        if hasattr(va, "__len__"):
            assert list(va) == list(vb), name
        else:
            assert va == vb, name


def test_config_struct_matches_header(lib):
    # sizeof(usvmpc_config) as laid out by ctypes must be what the C side reads: probe with config_default
    cfg = _lib.Config()
    assert lib.usvmpc_config_default(C.byref(cfg), 0) == 0
    assert (cfg.N, cfg.num_stages, cfg.nlp_type, cfg.max_iter, cfg.qp_iter_max) == (20, 4, 1, 100, 50)
    assert cfg.uh == 1e6 and list(cfg.tol) == [1e-6] * 4
    assert cfg.W[0] == 1.0 and cfg.W[9] == 1.0 and cfg.W[1] == 0.0 and cfg.W_e[7] == 1.0
    assert lib.usvmpc_config_default(C.byref(cfg), 7) < 0
    assert b"unknown model" in lib.usvmpc_last_error()


def test_description_to_config(lib):
    P = rh.RefProblem(N=40, K=5, num_steps=4)
    cfg, model, nx, nu = config_from_ocp(ocp_from_problem(P))
    assert (model, nx, nu, cfg.N, cfg.K, cfg.nbx, cfg.nbu, cfg.num_steps, cfg.nlp_type) == (0, 6, 2, 40, 5, 3, 2, 4, 0)
    assert abs(cfg.dt - 0.05) < 1e-15 and list(cfg.idxbx)[:3] == [3, 4, 5]
    np.testing.assert_array_equal(np.array(cfg.W[:64]).reshape(8, 8, order="F"), P.W)
    ocp = ocp_from_problem(P)
    ocp.cost.cost_type = "NONLINEAR_LS"
    with pytest.raises(Exception, match="LINEAR_LS"):
        config_from_ocp(ocp)
    ocp = ocp_from_problem(P)
    ocp.solver_options.qp_solver = "FULL_CONDENSING_QPOASES"
    with pytest.raises(Exception, match="PARTIAL_CONDENSING_HPIPM"):
        config_from_ocp(ocp)
    ocp = ocp_from_problem(P)
    ocp.model.name = "race_car"
    with pytest.raises(Exception, match="device models"):
        config_from_ocp(ocp)


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from mpc_collisionavoidance_b200 import BatchedAcadosOcpSolver
    with pytest.raises(Exception, match="CUDA|cuda"):
        BatchedAcadosOcpSolver(ocp_from_problem(rh.RefProblem(N=20, K=3)), batch=4)


def test_json_description_round_trip(lib, tmp_path):
    # acados_ocp_nlp.json layout (acados_ocp_solver.py:416-444): description -> json -> description -> same config
    from mpc_collisionavoidance_b200.ocp import ocp_formulation_json_dump, ocp_formulation_json_load
    from mpc_collisionavoidance_b200.workloads import benchmark_ocp
    ocp = benchmark_ocp(2)
    f = str(tmp_path / "acados_ocp_nlp.json")
    ocp_formulation_json_dump(ocp, f)
    back = ocp_formulation_json_load(f)
    a, b = config_from_ocp(ocp)[0], config_from_ocp(back)[0]
    assert bytes(a) == bytes(b)
    np.testing.assert_array_equal(back.constraints.x0, ocp.constraints.x0)
    import json
    d = json.load(open(f))
    assert d["constraints"]["lbx_0"] == d["constraints"]["ubx_0"] and d["solver_options"]["nlp_solver_type"] == "SQP"


def build_node_driver(out_dir):
    """gcc tests/c/ca_node_step.c against libusvmpc.so alone: the generated solver's symbol names must all resolve"""
    import subprocess
    exe = os.path.join(str(out_dir), "ca_node_step")
    pkg = os.path.join(ROOT, "mpc_collisionavoidance_b200")
    subprocess.run(["gcc", "-O1", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "c", "ca_node_step.c"),
                    "-o", exe, "-L" + pkg, "-lusvmpc", "-Wl,-rpath," + pkg, "-Wl,--no-undefined"], check=True)
    return exe


def test_node_driver_links_with_the_generated_solver_symbols(lib, tmp_path):
    import subprocess
    exe = build_node_driver(tmp_path)
    # no input file: exits before any device call
    assert subprocess.run([exe]).returncode == 2
