"""N > 1 coverage.  CPU part (world_size 2, gloo): the shard maps used by every rank form a partition, are
balanced and merge back to the caller's layout.  GPU part: tests/mgpu_check.py under torchrun on 2 GPUs (NCCL),
skipped when the box has a single GPU."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from realsensecalibration_b200 import cuda, sharding, synthetic as S

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        pr = S.bal_like(30, 2000, 5, 12, 5, variable_degree=True)
        sh = sharding.shard_model_a(pr.n_cam, pr.n_pt, pr.cam_idx, pr.pt_idx, pr.obs_xy, pr.params, rank, world)
        got = [None] * world
        dist.all_gather_object(got, (sh.lo, sh.hi, sh.obs_sel, sh.params))
        # every observation belongs to exactly one rank, shards are contiguous point ranges covering [0, n_pt)
        allsel = np.sort(np.concatenate([g[2] for g in got]))
        ok = np.array_equal(allsel, np.arange(pr.n_obs))
        ok = ok and got[0][0] == 0 and got[-1][1] == pr.n_pt and all(got[i][1] == got[i + 1][0] for i in range(world - 1))
        loads = np.array([len(g[2]) for g in got])
        ok = ok and loads.max() - loads.min() <= 2 * 16
        # local indices address the local parameter vector correctly
        ok = ok and np.array_equal(sh.params[6 * pr.n_cam:].reshape(-1, 3)[sh.pt_idx],
                                   pr.params[6 * pr.n_cam:].reshape(-1, 3)[pr.pt_idx[sh.obs_sel]])
        # a "solve" that adds rank-independent camera updates and local point updates merges back exactly
        merged = sharding.merge_model_a(pr.n_cam, pr.n_pt, [(g[0], g[1], g[3] + 1.0) for g in got])
        ok = ok and np.array_equal(merged, pr.params + 1.0)
        # the sum every rank would hand to ncclAllReduce: per-camera observation counts add up to the global ones
        t = torch.from_numpy(np.bincount(sh.cam_idx, minlength=pr.n_cam).astype(np.int64))
        dist.all_reduce(t)
        ok = ok and np.array_equal(t.numpy(), np.bincount(pr.cam_idx, minlength=pr.n_cam))
        # Model B by frame
        pb = S.marker_rig_b(3, 5, 9, 2, visibility=0.6)
        sb = sharding.shard_model_b(pb.n_cam, pb.n_time, pb.n_marker, pb.time_idx, pb.cam_idx, pb.marker_idx, pb.obs8, pb.params, rank, world)
        gb = [None] * world
        dist.all_gather_object(gb, (sb.lo, sb.hi, sb.params, sb.obs_sel))
        ok = ok and np.array_equal(np.sort(np.concatenate([g[3] for g in gb])), np.arange(pb.n_mobs))
        ok = ok and np.array_equal(sharding.merge_model_b(pb.n_cam, pb.n_time, pb.n_marker, [(g[0], g[1], g[2]) for g in gb]), pb.params)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_shard_maps_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(60)
    assert sorted(res) == [(0, True), (1, True)]


def test_single_rank_shard_is_identity():
    pr = S.bal_like(10, 100, 4, 6, 1)
    sh = sharding.shard_model_a(pr.n_cam, pr.n_pt, pr.cam_idx, pr.pt_idx, pr.obs_xy, pr.params, 0, 1)
    assert sh.lo == 0 and sh.hi == pr.n_pt and np.array_equal(sh.pt_idx, pr.pt_idx) and np.array_equal(sh.params, pr.params)


@pytest.mark.gpu
def test_two_gpu_solve_matches_oracle():
    if cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29611", os.path.join(ROOT, "tests", "mgpu_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "[mgpu] OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
