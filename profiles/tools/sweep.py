"""Sweep of the tuning switches of the fused Model A passes on one workload (tuning aid, not a bench):
   python profiles/tools/sweep.py cfg5 "SA_TOBS=768,SA_L=8 SA_TOBS=576,SA_L=12 SA=0,FA_TOBS=1024"   (BA_<name>=<value>)"""
import os, sys, time, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import bench
from realsensecalibration_b200 import cuda

name = sys.argv[1]
combos = [dict(kv.split("=") for kv in c.split(",") if kv) for c in sys.argv[2].split()]
job = bench.Job(name, 0, 1)
job.pin()
stream = torch.cuda.Stream()
P = cuda.Problem(0)
P.set_stream(stream.cuda_stream)
for combo in combos:
    for k in list(os.environ):
        if k.startswith("BA_FA_") or k.startswith("BA_SA"): del os.environ[k]
    for k, v in combo.items(): os.environ["BA_" + k] = v
    t_build = time.perf_counter()
    job.set_model(P)
    P.set_parameters(job.params)
    t_build = time.perf_counter() - t_build
    P.save_parameters()
    opts = bench.bench_options(cuda, profile=False)
    bench.run_steps(P, opts, 5)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record(stream)
        s, rows = bench.run_steps(P, opts, 10)
        e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    P.reset_stats()
    with torch.cuda.stream(stream):
        bench.run_steps(P, bench.bench_options(cuda, profile=True), 10)
    torch.cuda.synchronize()
    st = {k["name"]: (k["launches"], k["total_ms"] / k["launches"]) for k in P.kernel_stats()}
    print(json.dumps({"workload": name, "env": combo, "set_model_ms": round(1e3 * t_build, 1), "ms_per_it": round(ms, 4),
                      "final_cost": s.final_cost, "pcg_iters": [r["linear_solver_iterations"] for r in rows],
                      "kernels": {k: (v[0], round(v[1], 4)) for k, v in st.items() if v[1] > 0.02}}), flush=True)
