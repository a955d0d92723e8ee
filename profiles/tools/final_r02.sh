#!/bin/bash
# End-of-round evidence of the final code on one B200: tests, ncu captures, launch lists, bench lines (-> gpurun_out/, copied to profiles/).
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/r02f_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02f_smoke.log 2>&1
B="python bench.py --steps 5 --warmup 0 --no-cpu-baseline --no-e2e --no-jacobian"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_sa_pass1|k_fa_pass2|k_reduce_items|k_pcg" -s 6 -c 8 -o gpurun_out/r02f_prof_cfg5 -f $B --workload cfg5 > gpurun_out/ncu_cfg5.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_jac_b|k_inc_W|k_pairs_partial|k_model_cost|k_fobs_partial|k_dobs_partial|k_e_M|k_chol_blocked" -s 10 -c 10 -o gpurun_out/r02f_prof_cfg3 -f $B --workload cfg3 > gpurun_out/ncu_cfg3.log 2>&1
python profiles/tools/ncu_summary.py gpurun_out/r02f_prof_cfg5.ncu-rep > gpurun_out/r02f_ncu_full_cfg5.txt 2>/dev/null
python profiles/tools/ncu_summary.py gpurun_out/r02f_prof_cfg3.ncu-rep > gpurun_out/r02f_ncu_full_cfg3.txt 2>/dev/null
rm -f gpurun_out/r02f_prof_cfg5.ncu-rep gpurun_out/r02f_prof_cfg3.ncu-rep   # gpurun brings back at most 64 MiB
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02f_launches_cfg5.csv python bench.py --steps 5 --warmup 0 --no-cpu-baseline --no-e2e > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02f_launches_hongo.csv python bench.py --workload hongo --steps 10 --warmup 0 --no-cpu-baseline --no-e2e > /dev/null 2>&1
for w in cfg5 hongo cfg1 cfg2 cfg3 cfg3a cfg4; do
  python bench.py --workload $w > gpurun_out/r02f_bench_${w}_n1.json 2> gpurun_out/bench_$w.err
done
python bench.py --impl reference --workload hongo --steps 20 --warmup 3 > gpurun_out/r02f_bench_hongo_reference_arm.json 2>/dev/null
tail -2 gpurun_out/r02f_pytest_gpu.log; cat gpurun_out/r02f_smoke.log | tail -1
for w in cfg5 hongo cfg1 cfg2 cfg3 cfg3a cfg4; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02f_bench_${w}_n1.json").read().strip().splitlines()[-1])
    print("$w", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "cpu", round(d["cpu_baseline"]["value"],2), d.get("path_used"), "frac", (d.get("roofline") or {}).get("frac"), "parity", (d.get("parity") or {}).get("rel"))
except Exception as e: print("$w", "ERR", e)
PY
done
