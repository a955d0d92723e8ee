"""Synthetic problem generators for the BASELINE.json configurations.

Host-side numpy only: they produce the arrays a caller would hand to ba_cuda_set_model_*.
Intrinsics are those of the reference's camera 821312061029
(Common/Calibration/Intrinsics/821312061029.xml), 640x480, OpenCV +z-forward convention.
"""
from dataclasses import dataclass

import numpy as np

FX = 624.01068115234375
PPX = 315.53594970703125
PPY = 231.18597412109375
INTR = np.array([FX, FX, PPX, PPY], np.float64)


def _rodrigues(rv):
    rv = np.asarray(rv, np.float64)
    t = np.linalg.norm(rv, axis=-1, keepdims=True)
    k = rv / np.maximum(t, 1e-300)
    K = np.zeros(rv.shape[:-1] + (3, 3))
    K[..., 0, 1], K[..., 0, 2] = -k[..., 2], k[..., 1]
    K[..., 1, 0], K[..., 1, 2] = k[..., 2], -k[..., 0]
    K[..., 2, 0], K[..., 2, 1] = -k[..., 1], k[..., 0]
    t = t[..., None]
    return np.eye(3) + np.sin(t) * K + (1 - np.cos(t)) * (K @ K)


def _rvec_from_R(R):
    # inverse Rodrigues for a single rotation (angles well inside (0, pi))
    tr = np.clip((np.trace(R) - 1) / 2, -1, 1)
    th = np.arccos(tr)
    if th < 1e-12:
        return np.zeros(3)
    v = np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]]) / (2 * np.sin(th))
    return v * th


@dataclass
class ModelA:
    n_cam: int
    n_pt: int
    cam_idx: np.ndarray
    pt_idx: np.ndarray
    obs_xy: np.ndarray
    intr: np.ndarray      # [n_cam, 4]
    params: np.ndarray    # initial guess, [6 n_cam + 3 n_pt]
    truth: np.ndarray

    @property
    def n_obs(self):
        return int(self.cam_idx.shape[0])


@dataclass
class ModelB:
    n_cam: int
    n_time: int
    n_marker: int
    time_idx: np.ndarray
    cam_idx: np.ndarray
    marker_idx: np.ndarray
    obs8: np.ndarray
    intr: np.ndarray
    params: np.ndarray
    truth: np.ndarray
    marker_side: float
    counts: np.ndarray = None

    @property
    def n_mobs(self):
        return int(self.time_idx.shape[0])


def _project(cams, X, cam_idx):
    """cams [n,6] angle-axis|t ; X [m,3] ; returns pixel coords of X[i] in camera cam_idx[i] and depth."""
    R = _rodrigues(cams[:, :3])
    p = np.einsum("nij,nj->ni", R[cam_idx], X) + cams[cam_idx, 3:]
    return np.stack([FX * p[:, 0] / p[:, 2] + PPX, FX * p[:, 1] / p[:, 2] + PPY], 1), p[:, 2]


def bal_like(n_cam, n_pt, obs_per_pt, window, seed, noise_px=0.5, perturb=(0.02, 0.005, 0.005), variable_degree=False):
    """BAL-shaped Model A problem (cfg4 / cfg5): cameras along a smooth trajectory looking at a slab of points;
    every point is seen by `obs_per_pt` (or a clipped Poisson number of) distinct cameras inside a window."""
    rng = np.random.default_rng(seed)
    s = np.linspace(0.0, 1.0, n_cam)
    length = 0.05 * n_cam
    cam_pos = np.stack([length * s, 0.3 * np.sin(8 * np.pi * s), 0.1 * np.cos(6 * np.pi * s)], 1)
    rv_true = 0.05 * np.stack([np.sin(5 * np.pi * s), np.cos(7 * np.pi * s), np.sin(3 * np.pi * s)], 1)
    R = _rodrigues(rv_true)
    t_true = -np.einsum("nij,nj->ni", R, cam_pos)
    cams = np.concatenate([rv_true, t_true], 1)
    if variable_degree:
        k = np.clip(rng.poisson(obs_per_pt, n_pt), 2, 16).astype(np.int64)
    else:
        k = np.full(n_pt, obs_per_pt, np.int64)
    w = max(window, int(k.max()))
    centre = np.sort(rng.integers(0, n_cam, n_pt))
    start = np.clip(centre - w // 2, 0, max(n_cam - w, 0))
    n_obs = int(k.sum())
    pt_idx = np.repeat(np.arange(n_pt, dtype=np.int64), k)
    # distinct cameras per point: random offsets without replacement inside the window
    first = np.concatenate([[0], np.cumsum(k)[:-1]])
    rank_in_pt = np.arange(n_obs) - np.repeat(first, k)
    stride = np.maximum(w // np.repeat(k, k), 1)
    jitter = rng.integers(0, 1 << 30, n_obs) % stride
    cam_idx = np.minimum(np.repeat(start, k) + rank_in_pt * stride + jitter, n_cam - 1)
    # points in front of their window's central camera
    cc = np.minimum(start + w // 2, n_cam - 1)
    local = np.stack([rng.uniform(-0.6, 0.6, n_pt), rng.uniform(-0.45, 0.45, n_pt), rng.uniform(2.0, 4.0, n_pt)], 1)
    X = np.einsum("nji,nj->ni", R[cc], local - t_true[cc])
    uv, depth = _project(cams, X[pt_idx], cam_idx)
    keep = depth > 0.2
    if not np.all(keep):  # never drop: re-place offending points straight ahead instead
        uv[~keep] = (PPX, PPY)
    uv = uv + rng.normal(0.0, noise_px, uv.shape)
    truth = np.concatenate([cams.ravel(), X.ravel()])
    # initial guess: rotation perturbed about the camera CENTRE (perturbing t = -R c directly would move a camera that
    # sits 50 m down the trajectory by metres), centre and points jittered
    rv_init = rv_true + rng.normal(0, perturb[0], (n_cam, 3))
    c_init = cam_pos + rng.normal(0, perturb[1], (n_cam, 3))
    t_init = -np.einsum("nij,nj->ni", _rodrigues(rv_init), c_init)
    init_c = np.concatenate([rv_init, t_init], 1)
    init_X = X + rng.normal(0, perturb[2], X.shape)
    params = np.concatenate([init_c.ravel(), init_X.ravel()])
    return ModelA(n_cam, n_pt, cam_idx.astype(np.int32), pt_idx.astype(np.int32), uv, np.tile(INTR, (n_cam, 1)), params, truth)


def _rig(n_cam, n_marker, n_time, seed, marker_side, radius=0.35):
    rng = np.random.default_rng(seed)
    # cameras on a ring looking inward; camera 0 is the base camera (identity)
    ang = 2 * np.pi * np.arange(n_cam) / n_cam
    world_R = []
    world_c = []
    for a in ang:
        c = np.array([radius * np.sin(a), 0.0, radius * (1 - np.cos(a))])  # camera 0 at origin looking +z to the centre
        z = np.array([0.0, 0.0, radius]) - c
        z /= np.linalg.norm(z)
        x = np.cross(np.array([0.0, 1.0, 0.0]), z); x /= np.linalg.norm(x)
        y = np.cross(z, x)
        world_R.append(np.stack([x, y, z], 0))  # world(base cam) -> camera
        world_c.append(c)
    cams = np.zeros((n_cam, 6))
    for i in range(n_cam):
        Rw = world_R[i] @ world_R[0].T
        cams[i, :3] = _rvec_from_R(Rw)
        cams[i, 3:] = -Rw @ (world_R[0] @ (world_c[i] - world_c[0]))
    cams[0] = 0.0
    # markers on a small polyhedron around the object origin (marker 0 = base marker, identity)
    mk = np.zeros((n_marker, 6))
    dirs = rng.normal(size=(n_marker, 3)); dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    for m in range(1, n_marker):
        mk[m, :3] = rng.normal(0, 0.8, 3)
        mk[m, 3:] = 0.03 * dirs[m] * (1 + 0.3 * (m % 5))
    # object poses per frame, in front of camera 0 near the ring centre
    fr = np.zeros((n_time, 6))
    fr[:, :3] = rng.normal(0, 0.5, (n_time, 3))
    fr[:, 3:] = np.array([0.0, 0.0, radius]) + rng.normal(0, 0.03, (n_time, 3))
    return rng, cams, fr, mk


def marker_rig_b(n_cam, n_marker, n_time, seed, marker_side=0.0148, visibility=1.0, noise_px=0.5,
                 perturb=(0.02, 0.005)):
    """Model B marker rig (cfg2 / cfg3): (time, camera, marker) observations of 4 corners each."""
    rng, cams, fr, mk = _rig(n_cam, n_marker, n_time, seed, marker_side)
    t, c, m = np.meshgrid(np.arange(n_time), np.arange(n_cam), np.arange(n_marker), indexing="ij")
    t, c, m = t.ravel(), c.ravel(), m.ravel()
    if visibility < 1.0:
        keep = rng.random(t.shape[0]) < visibility
        t, c, m = t[keep], c[keep], m[keep]
    h = marker_side / 2
    corners = np.array([[-h, h, 0], [h, h, 0], [h, -h, 0], [-h, -h, 0]])
    Rm, Rt, Rc = _rodrigues(mk[:, :3]), _rodrigues(fr[:, :3]), _rodrigues(cams[:, :3])
    obs = np.zeros((t.shape[0], 8))
    for j in range(4):
        p = np.einsum("nij,j->ni", Rm[m], corners[j]) + mk[m, 3:]
        p = np.einsum("nij,nj->ni", Rt[t], p) + fr[t, 3:]
        p = np.einsum("nij,nj->ni", Rc[c], p) + cams[c, 3:]
        obs[:, 2 * j] = FX * p[:, 0] / p[:, 2] + PPX
        obs[:, 2 * j + 1] = FX * p[:, 1] / p[:, 2] + PPY
    obs += rng.normal(0, noise_px, obs.shape)
    truth = np.concatenate([cams.ravel(), fr.ravel(), mk.ravel()])

    def pert(a, fixed0):
        b = a + np.concatenate([rng.normal(0, perturb[0], (a.shape[0], 3)), rng.normal(0, perturb[1], (a.shape[0], 3))], 1)
        if fixed0:
            b[0] = a[0]
        return b
    params = np.concatenate([pert(cams, True).ravel(), pert(fr, False).ravel(), pert(mk, True).ravel()])
    counts = np.zeros((n_time, n_cam), np.int32)
    np.add.at(counts, (t, c), 1)
    return ModelB(n_cam, n_time, n_marker, t.astype(np.int32), c.astype(np.int32), m.astype(np.int32), obs,
                  np.tile(INTR, (n_cam, 1)), params, truth, marker_side, counts)


def marker_rig_a(n_cam, n_marker, n_time, seed, marker_side=0.0148, noise_px=0.5, perturb=(0.02, 0.005, 0.005)):
    """Model A reading of the rig (cfg3-A, "pose + marker-corner BA"): every marker corner at every frame is a free
    3-D point seen by all cameras."""
    rng, cams, fr, mk = _rig(n_cam, n_marker, n_time, seed, marker_side)
    h = marker_side / 2
    corners = np.array([[-h, h, 0], [h, h, 0], [h, -h, 0], [-h, -h, 0]])
    Rm, Rt = _rodrigues(mk[:, :3]), _rodrigues(fr[:, :3])
    pm = np.einsum("mij,cj->mci", Rm, corners) + mk[:, None, 3:]            # [M,4,3] in the object frame
    X = np.einsum("tij,mcj->tmci", Rt, pm) + fr[:, None, None, 3:]          # [T,M,4,3] in the base camera
    X = X.reshape(-1, 3)
    n_pt = X.shape[0]
    pt_idx = np.repeat(np.arange(n_pt), n_cam)
    cam_idx = np.tile(np.arange(n_cam), n_pt)
    uv, _ = _project(cams, X[pt_idx], cam_idx)
    uv += rng.normal(0, noise_px, uv.shape)
    truth = np.concatenate([cams.ravel(), X.ravel()])
    init_c = cams + np.concatenate([rng.normal(0, perturb[0], (n_cam, 3)), rng.normal(0, perturb[1], (n_cam, 3))], 1)
    init_X = X + rng.normal(0, perturb[2], X.shape)
    params = np.concatenate([init_c.ravel(), init_X.ravel()])
    return ModelA(n_cam, n_pt, cam_idx.astype(np.int32), pt_idx.astype(np.int32), uv, np.tile(INTR, (n_cam, 1)), params, truth)


def two_cam_like(n_frames, seed, marker_side=0.032, noise_px=0.3):
    """cfg1: two_cam_data.txt-shaped Model A problem: one relative camera pose, one marker re-posed for n_frames,
    every corner observed once (num_observations == num_points, Test1 bundle_adjustmenter.cpp:64)."""
    rng = np.random.default_rng(seed)
    cam = np.array([[-0.04, -0.02, 0.004, 0.15, 0.0, 0.004]])
    h = marker_side / 2
    corners = np.array([[-h, h, 0], [h, h, 0], [h, -h, 0], [-h, -h, 0]])
    rv = rng.normal(0, 0.3, (n_frames, 3))
    tv = np.stack([rng.uniform(-0.08, 0.08, n_frames), rng.uniform(-0.06, 0.06, n_frames), rng.uniform(0.25, 0.4, n_frames)], 1)
    X = (np.einsum("tij,cj->tci", _rodrigues(rv), corners) + tv[:, None, :]).reshape(-1, 3)
    n_pt = X.shape[0]
    cam_idx = np.zeros(n_pt, np.int32); pt_idx = np.arange(n_pt, dtype=np.int32)
    uv, _ = _project(cam, X, cam_idx)
    uv += rng.normal(0, noise_px, uv.shape)
    truth = np.concatenate([cam.ravel(), X.ravel()])
    init = np.concatenate([(cam + rng.normal(0, 0.01, cam.shape)).ravel(), (X + rng.normal(0, 0.003, X.shape)).ravel()])
    return ModelA(1, n_pt, cam_idx, pt_idx, uv, INTR.reshape(1, 4), init, truth)
