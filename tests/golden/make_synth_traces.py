"""make_synth_traces.py -- oracle LM traces of the BASELINE.json configurations at FULL size, committed as
tests/golden/synth_traces.json.

Run here (CPU, oracle/libba_oracle.so):   python tests/golden/make_synth_traces.py [cfg2 cfg3 cfg4 cfg5]

For every workload of bench.py: the rows (cost, trust-region radius, CG iterations, accepted) of ONE solve with bench.py's
options (ITERS_PER_SOLVE LM iterations from the synthetic start, tolerances off) -- what `bench.py` compares its final cost
with at N > 1 (`parity`) and what tests/test_gpu_parity_at_size.py holds the CUDA rows to.  The BAL-shaped problems run on
truncated PCG, whose stopping rule can flip by one CG iteration on round-off; a second trace with the CG iteration count
FIXED (pcg_min_iterations = pcg_max_iterations = FIXED_CG) removes the rule from the comparison."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from oracle import oracle_py as O  # noqa: E402

FIXED_CG = 12
OUT = os.path.join(ROOT, "tests", "golden", "synth_traces.json")


def trace(job, fixed_cg=None):
    o = O.default_options()
    o.function_tolerance = 0.0; o.parameter_tolerance = 0.0; o.gradient_tolerance = 0.0
    o.max_num_iterations = bench.ITERS_PER_SOLVE
    if fixed_cg is not None:
        o.pcg_min_iterations = fixed_cg; o.pcg_max_iterations = fixed_cg
    t0 = time.time()
    x, s, rows = job.oracle_solve(O, o, bench.host_threads())
    keep = ("iteration", "cost", "cost_change", "trust_region_radius", "linear_solver_iterations", "step_is_successful", "step_is_valid",
            "gradient_max_norm", "step_norm")
    return {"rows": [{k: r[k] for k in keep} for r in rows], "final_cost": s.final_cost, "rcs_dim": s.rcs_dim,
            "seconds": round(time.time() - t0, 1), "x_head": [float(v) for v in x[:12]]}


def main():
    names = sys.argv[1:] or ["cfg1", "cfg2", "cfg3", "cfg4", "cfg5"]
    try:
        with open(OUT) as f:
            out = json.load(f)
    except (OSError, ValueError):
        out = {}
    for name in names:
        job = bench.Job(name, 0, 1)
        pcg = job.oracle_solver(O) == O.SCHUR_PCG
        ent = {"how": "oracle/ba_oracle.cpp, %s, %d LM iterations, bench.py options" % ("Schur + block-Jacobi PCG (eta 0.1)" if pcg else "Schur + dense Cholesky",
                                                                                       bench.ITERS_PER_SOLVE),
               "description": bench.WORKLOADS[name]["desc"]}
        ent.update(trace(job))
        if pcg:
            ent["fixed_cg"] = FIXED_CG
            ent["fixed"] = trace(job, FIXED_CG)
        out[name] = ent
        print(name, "final cost %.12e" % ent["final_cost"], "%.1f s" % ent["seconds"], flush=True)
        with open(OUT, "w") as f:
            json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
