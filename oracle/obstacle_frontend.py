"""oracle/obstacle_frontend.py -- TEST INFRASTRUCTURE ONLY.

numpy restatement of the obstacle front end of the reference's guidance node
(catkin_ws/src/nmpc_ca/src/nmpc_guidance_ca1.cpp): obstaclesCallback :251-344 (radius inflation by boat_radius_ :267,
clearance sqrt(x^2+y^2) - radius :268, ascending sort :275 / sortVec :422-438, keep the obs_num_ nearest :284-294, or
pad with initializeObstacles :365-376 when there are at most obs_num_), body2NED :346-363 (float rotation by psi, then
+ nedx / nedy).  The node's vectors are Eigen::Vector3f, so positions and radii are rounded to float32 here too.
PARITY PIN: none exists in the reference (no tests, no fixtures for this node): parity unpinned; this file follows the
C++ line by line and the GPU kernel is compared with it.
"""
import numpy as np


def obstacle_frontend(pose, obs_body, lens, K, boat_radius=0.5, init_obs_pos=1000.0):
    pose = np.asarray(pose, dtype=np.float64); obs_body = np.asarray(obs_body, dtype=np.float64)
    B, M, _ = obs_body.shape
    p = np.zeros((B, 2 * K)); r = np.zeros((B, K)); chosen = -np.ones((B, K), dtype=int)
    for b in range(B):
        n = int(min(max(lens[b], 0), M))
        nedx, nedy, psi = pose[b]
        c, s = np.float32(np.cos(psi)), np.float32(np.sin(psi))
        if n > K:
            rad = obs_body[b, :n, 2] + boat_radius
            d = np.sqrt(obs_body[b, :n, 0] ** 2 + obs_body[b, :n, 1] ** 2) - rad
            idx = np.argsort(d, kind="stable")[:K]          # ascending clearance, ties: lower index first
        else:
            idx = np.arange(n)
        for i in range(K):
            if i < len(idx):
                j = idx[i]
                bx, by = np.float32(obs_body[b, j, 0]), np.float32(obs_body[b, j, 1])
                x = np.float32(np.float64(np.float32(c * bx) + np.float32(-s * by)) + nedx)   # R * body in float, + ned
                y = np.float32(np.float64(np.float32(s * bx) + np.float32(c * by)) + nedy)
                p[b, 2 * i], p[b, 2 * i + 1], r[b, i] = x, y, np.float32(obs_body[b, j, 2] + boat_radius)
                chosen[b, i] = j
            else:
                p[b, 2 * i] = p[b, 2 * i + 1] = np.float32(init_obs_pos)
    return p, r, chosen
