import os, sys, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import bench
from realsensecalibration_b200 import cuda
job = bench.Job("hongo", 0, 1)
P = cuda.Problem(0)
opts = bench.bench_options(cuda, profile=False)
for i in range(6):
    if i == 5: os.environ["BA_CUDA_TIMING"] = "1"
    t0 = time.perf_counter(); job.set_model(P); t1 = time.perf_counter(); P.set_parameters(job.params); t2 = time.perf_counter()
    s, rows = P.solve(opts); t3 = time.perf_counter(); x = P.get_parameters(); t4 = time.perf_counter()
    print("set_model %.0f us  set_parameters %.0f  solve %.0f  get %.0f" % (1e6*(t1-t0), 1e6*(t2-t1), 1e6*(t3-t2), 1e6*(t4-t3)), flush=True)
