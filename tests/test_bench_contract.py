"""The driver's bench contract, CPU side: `bench.py --impl reference` (the CPU oracle arm) prints ONE JSON line with the
keys the driver reads, on the smoke-sized workload; under torchrun every rank but 0 exits without work."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra_env=None):
    env = dict(os.environ)
    env.update(extra_env or {})
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "small", "--steps", "2", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    return [l for l in r.stdout.splitlines() if l.strip()]


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    lines = _run()
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "LM iterations/sec" and d["unit"] == "it/s" and d["higher_is_better"] is True
    assert d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 0 and d["value"] > 0 and d["ms_per_step"] > 0
    assert d["config"]["workload"] == "small" and d["dtype"] == "f64" and d["data"] == "synthetic" and d["vs_baseline"] is None
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "it/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_without_output():
    assert _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}) == []


def test_reference_arm_ignores_omp_num_threads_of_torchrun():
    # torch.distributed.run exports OMP_NUM_THREADS=1 to its workers: the CPU arm must still use every core it may run
    # on (round 1: the 30 M-observation reference run was single threaded under torchrun and timed out at N = 2, 4, 8)
    lines = _run({"OMP_NUM_THREADS": "1", "RANK": "0", "WORLD_SIZE": "2", "LOCAL_RANK": "0"})
    assert len(lines) == 1
    d = json.loads(lines[0])
    want = len(os.sched_getaffinity(0))
    assert d["cpu_baseline"]["cores"] == want
    assert d["cpu_baseline"]["final_cost"] > 0


def test_both_arms_print_the_same_config_keys():
    # the driver compares the two `config` objects: same keys (values are checked on the GPU box where both arms run)
    import bench
    ref = json.loads(_run()[0])["config"]
    job = bench.Job("small", 0, 1)
    ours = bench.config_of(job, "small", 1, 2, 300)
    assert sorted(ref) == sorted(ours)
    assert all(ref[k] == ours[k] for k in ref if k not in ("rcs_solver", "rcs_dim"))


def test_committed_traffic_numbers_belong_to_this_tree():
    """roofline.traffic comes from profiles/traffic.json; bench.py only uses an entry whose kernel_source_hash is the hash of
    the csrc/ files it runs.  The committed entries of the default workload must be the ones of the committed kernels."""
    import json
    import bench
    with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
        t = json.load(f)
    h = bench.kernel_source_hash()
    entries = t[bench.DEFAULT_WORKLOAD]
    assert entries and all(e.get("kernel_source_hash") == h for e in entries.values()), (h, entries)
