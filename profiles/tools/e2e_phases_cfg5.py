"""Host wall clock of the phases of one end-to-end solve of a workload (BA_CUDA_TIMING=1 prints the phases of set_model)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import bench
from realsensecalibration_b200 import cuda
name = sys.argv[1] if len(sys.argv) > 1 else "cfg5"
job = bench.Job(name, 0, 1)
job.pin()
P = cuda.Problem(0)
opts = bench.bench_options(cuda, profile=False)
for i in range(3):
    if i == 2: os.environ["BA_CUDA_TIMING"] = "2"
    t0 = time.perf_counter(); job.set_model(P); t1 = time.perf_counter(); P.set_parameters(job.params); t2 = time.perf_counter()
    s, rows = P.solve(opts); t3 = time.perf_counter(); x = P.get_parameters(); t4 = time.perf_counter()
    print("set_model %.2f ms  set_parameters %.2f  solve %.2f  get %.2f" % (1e3*(t1-t0), 1e3*(t2-t1), 1e3*(t3-t2), 1e3*(t4-t3)), flush=True)
