"""shard_build.py <workload> <world> <rank ...>: builds the problem of ONE rank's shard on one GPU (no NCCL) -- to reproduce and
locate build failures of a multi-GPU run under compute-sanitizer."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench
from realsensecalibration_b200 import cuda
name, world = sys.argv[1], int(sys.argv[2])
P = cuda.Problem(0)
for r in [int(x) for x in sys.argv[3:]]:
    job = bench.Job(name, r, world)
    try:
        job.set_model(P)
        P.set_parameters(job.params)
        s, rows = P.solve(bench.bench_options(cuda, False))
        print("rank", r, "ok: path", s.path_used, "final cost %.9e" % s.final_cost, flush=True)
    except Exception as e:
        print("rank", r, "FAILED:", e, flush=True)
        break
