"""Host-side shard maps for the multi-GPU path: eliminated blocks are independent under Schur, so
Model A is sharded by point and Model B by frame, each shard carrying all observations of its blocks; kept blocks
(cameras, markers) are replicated.  Pure index bookkeeping (numpy); the split points come from the C ABI's
ba_cuda_shard_blocks so that the C++ host and Python agree."""
from dataclasses import dataclass

import numpy as np

from . import cuda


@dataclass
class ShardA:
    lo: int
    hi: int
    n_pt: int
    cam_idx: np.ndarray
    pt_idx: np.ndarray      # local point index
    obs_xy: np.ndarray
    params: np.ndarray      # [6 n_cam | 3 n_pt_local]
    obs_sel: np.ndarray     # indices of this shard's observations in the caller's order


def shard_ranges(block_of_obs, n_blocks, world):
    deg = np.bincount(np.asarray(block_of_obs), minlength=n_blocks).astype(np.int64)
    return cuda.shard_blocks(deg, world)


def shard_model_a(n_cam, n_pt, cam_idx, pt_idx, obs_xy, params, rank, world):
    cam_idx = np.asarray(cam_idx); pt_idx = np.asarray(pt_idx); obs_xy = np.asarray(obs_xy); params = np.asarray(params)
    if world == 1:
        return ShardA(0, n_pt, n_pt, cam_idx, pt_idx, obs_xy, params, np.arange(cam_idx.shape[0]))
    r = shard_ranges(pt_idx, n_pt, world)
    lo, hi = int(r[rank]), int(r[rank + 1])
    sel = np.nonzero((pt_idx >= lo) & (pt_idx < hi))[0]
    pts = params[6 * n_cam:].reshape(-1, 3)[lo:hi]
    return ShardA(lo, hi, hi - lo, cam_idx[sel], (pt_idx[sel] - lo).astype(np.int32), obs_xy[sel],
                  np.concatenate([params[:6 * n_cam], pts.ravel()]), sel)


def merge_model_a(n_cam, n_pt, shard_results):
    """shard_results: list of (lo, hi, local parameter vector) from every rank -> the full parameter vector.
    Cameras are replicated (identical on every rank); points are concatenated in shard order."""
    x = np.zeros(6 * n_cam + 3 * n_pt)
    x[:6 * n_cam] = shard_results[0][2][:6 * n_cam]
    for lo, hi, xl in shard_results:
        x[6 * n_cam + 3 * lo:6 * n_cam + 3 * hi] = xl[6 * n_cam:]
    return x


@dataclass
class ShardB:
    lo: int
    hi: int
    n_time: int
    time_idx: np.ndarray    # local frame index
    cam_idx: np.ndarray
    marker_idx: np.ndarray
    obs8: np.ndarray
    params: np.ndarray      # [6 n_cam | 6 n_time_local | 6 n_marker]
    obs_sel: np.ndarray


def shard_model_b(n_cam, n_time, n_marker, time_idx, cam_idx, marker_idx, obs8, params, rank, world):
    time_idx = np.asarray(time_idx); cam_idx = np.asarray(cam_idx); marker_idx = np.asarray(marker_idx)
    obs8 = np.asarray(obs8).reshape(-1, 8); params = np.asarray(params)
    if world == 1:
        return ShardB(0, n_time, n_time, time_idx, cam_idx, marker_idx, obs8, params, np.arange(time_idx.shape[0]))
    r = shard_ranges(time_idx, n_time, world)
    lo, hi = int(r[rank]), int(r[rank + 1])
    sel = np.nonzero((time_idx >= lo) & (time_idx < hi))[0]
    C, T = n_cam, n_time
    x = np.concatenate([params[:6 * C], params[6 * C + 6 * lo:6 * C + 6 * hi], params[6 * (C + T):]])
    return ShardB(lo, hi, hi - lo, (time_idx[sel] - lo).astype(np.int32), cam_idx[sel], marker_idx[sel], obs8[sel], x, sel)


def merge_model_b(n_cam, n_time, n_marker, shard_results):
    C, T, M = n_cam, n_time, n_marker
    x = np.zeros(6 * (C + T + M))
    first = shard_results[0][2]
    x[:6 * C] = first[:6 * C]
    x[6 * (C + T):] = first[len(first) - 6 * M:]
    for lo, hi, xl in shard_results:
        x[6 * C + 6 * lo:6 * C + 6 * hi] = xl[6 * C:6 * C + 6 * (hi - lo)]
    return x
