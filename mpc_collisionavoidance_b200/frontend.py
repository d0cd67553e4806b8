"""Batched obstacle front end: the step right before the solver in the guidance node
(nmpc_ca/src/nmpc_guidance_ca1.cpp:251-363): nearest-K selection, body -> NED transform, radius inflation, on the GPU."""
import ctypes as C

from . import _lib


def select_obstacles(pose, obs_body, lens, K, boat_radius=0.5, init_obs_pos=1000.0):
    """pose [B,3] (nedx, nedy, psi), obs_body [B,M,3] (x, y, radius in the body frame), lens [B] int32: CUDA tensors.
    Returns (p [B,2K], r [B,K]) CUDA float64 tensors ready for solver.set(j, "p", p) / constraints_set(j, "lh", r)."""
    import torch
    lib = _lib.load()
    pose = pose.to(torch.float64).contiguous(); obs_body = obs_body.to(torch.float64).contiguous()
    lens = lens.to(torch.int32).contiguous()
    if not (pose.is_cuda and obs_body.is_cuda and lens.is_cuda):
        raise Exception("select_obstacles(): tensors must live on the GPU (the engine has no CPU path)")
    B, M = obs_body.shape[0], obs_body.shape[1]
    p = torch.empty((B, 2 * K), dtype=torch.float64, device=pose.device)
    r = torch.empty((B, K), dtype=torch.float64, device=pose.device)
    st = C.c_void_p(torch.cuda.current_stream(pose.device).cuda_stream)
    _lib.check(lib.usvmpc_obstacle_frontend(C.c_void_p(pose.data_ptr()), C.c_void_p(obs_body.data_ptr()), C.c_void_p(lens.data_ptr()),
                                            B, M, int(K), float(boat_radius), float(init_obs_pos), C.c_void_p(p.data_ptr()),
                                            C.c_void_p(r.data_ptr()), st), "obstacle_frontend")
    return p, r
