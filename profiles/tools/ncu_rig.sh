set -e
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_rig_lm -c 1 -s 3 -o gpurun_out/r02_rig_hongo -f python bench.py --workload hongo --steps 5 --warmup 5 > gpurun_out/ncu_rig.log 2>&1 || tail -5 gpurun_out/ncu_rig.log
ls -la gpurun_out/r02_rig_hongo.ncu-rep
