"""Where the one-CTA rig path stops paying: LM it/s of rig_b(4 cameras, 10 markers, F frames) with BA_RIG=1 and 0 (tuning aid)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import bench
from realsensecalibration_b200 import cuda

for frames in [int(a) for a in sys.argv[1:]] or [5, 12, 25, 50, 100]:
    bench.WORKLOADS["x"] = dict(desc="", kind="rig_b", args=(4, 10, frames, 0xBA02))
    out = {"frames": frames, "rows": 4 * 10 * frames * 8}
    for rig in ("1", "0"):
        os.environ["BA_RIG"] = rig
        os.environ["BA_RIG_MAX_ROWS"] = "1000000"
        job = bench.Job("x", 0, 1)
        P = cuda.Problem(0)
        stream = torch.cuda.Stream()
        P.set_stream(stream.cuda_stream)
        job.set_model(P)
        P.set_parameters(job.params)
        P.save_parameters()
        opts = bench.bench_options(cuda, profile=False)
        bench.run_steps(P, opts, 10)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record(stream)
            s, rows = bench.run_steps(P, opts, 20)
            e1.record(stream)
        torch.cuda.synchronize()
        out["rig" + rig + "_us_per_it"] = round(1e3 * e0.elapsed_time(e1) / 20, 1)
        out["path" + rig] = int(s.path_used)
        P.close()
    print(json.dumps(out), flush=True)
