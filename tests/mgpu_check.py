"""Multi-GPU parity check, one rank per GPU:  torchrun --nnodes=1 --nproc-per-node N tests/mgpu_check.py

Every rank solves its shard (points / frames) of the same problem through the C ABI with the NCCL-summed reduced
camera system; rank 0 then compares the merged result with the single-process CPU oracle on the whole problem.
Exit code 0 = parity green.  Used by tests/test_multi_gpu.py (skipped when fewer than 2 GPUs are visible)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from realsensecalibration_b200 import abi, cuda, formats as F, sharding, synthetic as S  # noqa: E402


def rel(a, b):
    return abs(a - b) / max(abs(b), 1e-300)


def stage(rank, what):
    if os.environ.get("MGPU_VERBOSE"):
        print("[mgpu stage] rank %d: %s" % (rank, what), file=sys.stderr, flush=True)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    stage(rank, "start")
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ids = [cuda.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    stage(rank, "unique id broadcast")
    P = cuda.Problem(local)
    P.comm_init(rank, world, ids[0])
    stage(rank, "comm_init done")
    failures = []

    def check(name, cond, msg=""):
        if not cond:
            failures.append("%s: %s" % (name, msg))

    # ---- Model A, dense and PCG reduced-system solvers -------------------------------------------------
    # (name, problem, solver, fixed CG count or None, environment of the build)
    chain = S.bal_like(300, 20000, 5, 20, 17)
    band = S.bal_like(200, 12000, 8, 24, 19)   # >= 4 pair products per observation: pass 1 on strips
    cases = (("balA-dense", S.bal_like(60, 5000, 6, 16, 13, variable_degree=True), abi.RCS_DENSE_CHOLESKY, None, {}),
             ("chain-pcg", chain, abi.RCS_PCG, None, {}),
             ("chain-pcg-k15", chain, abi.RCS_PCG, 15, {}),                       # sparse all-gather exchange (ba_exchange.cuh)
             ("chain-pcg-k15-allreduce", chain, abi.RCS_PCG, 15, {"BA_SX": "0"}),  # the all-reduce path on the same problem
             ("band-pcg-k15", band, abi.RCS_PCG, 15, {}),
             ("band-pcg-k15-tiles", band, abi.RCS_PCG, 15, {"BA_SA": "0"}))
    for name, pr, solver, fixed_cg, env in cases:
        old_env = {k: os.environ.get(k) for k in env}
        os.environ.update(env)
        sh = sharding.shard_model_a(pr.n_cam, pr.n_pt, pr.cam_idx, pr.pt_idx, pr.obs_xy, pr.params, rank, world)
        stage(rank, name + ": set_model_a")
        P.set_model_a(pr.n_cam, sh.n_pt, sh.cam_idx, sh.pt_idx, sh.obs_xy, pr.intr)
        P.set_parameters(sh.params)
        opt = cuda.default_options()
        opt.rcs_solver = solver
        opt.max_num_iterations = 8
        if fixed_cg is not None:
            opt.pcg_min_iterations = fixed_cg; opt.pcg_max_iterations = fixed_cg
        stage(rank, name + ": solve")
        s, rows = P.solve(opt)
        stage(rank, name + ": solved")
        for k, v in old_env.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
        xl = P.get_parameters()
        got = [None] * world
        dist.all_gather_object(got, (sh.lo, sh.hi, xl))
        if rank == 0:
            from oracle import oracle_py as O
            x = sharding.merge_model_a(pr.n_cam, pr.n_pt, got)
            oo = O.default_options(); oo.rcs_solver = solver; oo.max_num_iterations = 8
            if fixed_cg is not None:
                oo.pcg_min_iterations = fixed_cg; oo.pcg_max_iterations = fixed_cg
            xo, so, rows_o = O.solve_model_a(pr.n_cam, pr.n_pt, pr.cam_idx, pr.pt_idx, pr.obs_xy, pr.intr, pr.params, options=oo,
                                             linear_solver=O.SCHUR_DENSE if solver == abi.RCS_DENSE_CHOLESKY else O.SCHUR_PCG, n_threads=8)
            check(name, len(rows) == len(rows_o), "rows %d vs %d" % (len(rows), len(rows_o)))
            same = True
            exact = solver == abi.RCS_DENSE_CHOLESKY or fixed_cg is not None
            worst = 0.0
            for a, b in zip(rows, rows_o):
                same = same and a["linear_solver_iterations"] == b["linear_solver_iterations"]
                # dense Cholesky, or CG with the iteration count fixed on both sides: the same method, rows to round-off.
                # Ceres' stopping rule can flip by one CG iteration on the different summation order: looser bound there
                tol = (1e-10 if solver == abi.RCS_DENSE_CHOLESKY else 1e-9) if exact else (1e-5 if same else 1e-3)
                worst = max(worst, rel(a["cost"], b["cost"]))
                check(name, rel(a["cost"], b["cost"]) <= tol, "row %d cost %.15e vs %.15e" % (a["iteration"], a["cost"], b["cost"]))
                check(name, a["step_is_successful"] == b["step_is_successful"], "row %d accept flag" % a["iteration"])
            check(name, np.abs(x - xo).max() < (1e-7 if exact else 1e-2), "max |x - oracle| = %.3e" % np.abs(x - xo).max())
            print("[mgpu] %-24s world=%d rows=%d path=%d final cost %.12e  worst row %.1e  max|x-oracle| %.2e  coll %.3f ms" %
                  (name, world, len(rows), s.path_used, s.final_cost, worst, np.abs(x - xo).max(), s.ms_collective))
        # cameras must be bit-identical on every rank (replicated LM step)
        cams = torch.tensor(xl[:6 * pr.n_cam], device="cuda")
        lo_t, hi_t = cams.clone(), cams.clone()
        dist.all_reduce(lo_t, op=dist.ReduceOp.MIN); dist.all_reduce(hi_t, op=dist.ReduceOp.MAX)
        check(name, bool(torch.equal(lo_t, hi_t)), "camera blocks differ between ranks")

    # ---- Model B sharded by frame (dense RCS over cameras + markers) ------------------------------------
    pr = S.marker_rig_b(4, 12, 30, 7)
    sh = sharding.shard_model_b(pr.n_cam, pr.n_time, pr.n_marker, pr.time_idx, pr.cam_idx, pr.marker_idx, pr.obs8, pr.params, rank, world)
    P.set_model_b(pr.n_cam, sh.n_time, pr.n_marker, sh.time_idx, sh.cam_idx, sh.marker_idx, sh.obs8, pr.intr, pr.marker_side, 1)
    P.set_parameters(sh.params)
    s, rows = P.solve()
    xl = P.get_parameters()
    got = [None] * world
    dist.all_gather_object(got, (sh.lo, sh.hi, xl))
    if rank == 0:
        from oracle import oracle_py as O
        x = sharding.merge_model_b(pr.n_cam, pr.n_time, pr.n_marker, got)
        pb = F.ModelBFile(pr.n_time, pr.n_cam, pr.n_marker, pr.counts, pr.time_idx, pr.cam_idx, pr.marker_idx, pr.obs8, pr.params)
        xo, so, rows_o = O.solve_model_b(pb, pr.intr, pr.marker_side, 1)
        check("rigB", len(rows) == len(rows_o), "rows %d vs %d" % (len(rows), len(rows_o)))
        for a, b in zip(rows, rows_o):
            check("rigB", rel(a["cost"], b["cost"]) <= 1e-10, "row %d cost %.15e vs %.15e" % (a["iteration"], a["cost"], b["cost"]))
        check("rigB", np.abs(x - xo).max() < 1e-7, "max |x - oracle| = %.3e" % np.abs(x - xo).max())
        print("[mgpu] %-12s world=%d rows=%d final cost %.12e  max|x-oracle| %.2e" % ("rigB", world, len(rows), s.final_cost, np.abs(x - xo).max()))

    flag = torch.tensor([len(failures)], device="cuda")
    dist.all_reduce(flag)
    if failures:
        print("[mgpu] rank %d FAILURES:\n  " % rank + "\n  ".join(failures))
    elif rank == 0 and flag.item() == 0:
        print("[mgpu] OK world=%d" % world)
    P.close()
    dist.destroy_process_group()
    return 1 if flag.item() else 0


if __name__ == "__main__":
    sys.exit(main())
