"""e2e phase timing: python scratch/e2e.py cfg4"""
import os, sys, time, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import bench
from realsensecalibration_b200 import cuda
name = sys.argv[1]
job = bench.Job(name, 0, 1)
job.pin()
P = cuda.Problem(0)
opts = bench.bench_options(cuda, profile=False)
for rep in range(3):
    if rep == 2: os.environ["BA_CUDA_TIMING"] = "2"
    t = [time.perf_counter()]
    job.set_model(P); t.append(time.perf_counter())
    P.set_parameters(job.params); t.append(time.perf_counter())
    P.solve(opts); t.append(time.perf_counter())
    x = P.get_parameters(out=job.result); t.append(time.perf_counter())
    print(name, "rep", rep, "set_model %.2f set_params %.2f solve %.2f get %.2f total %.2f ms" % tuple(1e3 * v for v in (t[1]-t[0], t[2]-t[1], t[3]-t[2], t[4]-t[3], t[4]-t[0])), "h2d MB", job.h2d / 1e6, flush=True)
