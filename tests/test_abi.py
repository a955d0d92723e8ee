"""CPU-side checks of the drop-in boundary: the library loads without a GPU and exports every symbol the
header declares; the ctypes PODs have the C layout; the shard map is a balanced partition."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from realsensecalibration_b200 import abi, cuda

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "ba_cuda.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ba_cuda_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    L = cuda.lib()
    syms = _declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(L, s), "libba_cuda.so does not export %s" % s
    assert sorted(cuda.EXPORTS) == syms


def test_options_defaults_are_ceres_1_14():
    o = cuda.default_options()
    assert o.max_num_iterations == 50 and o.max_num_consecutive_invalid_steps == 5 and o.jacobi_scaling == 1
    assert o.initial_trust_region_radius == 1e4 and o.max_trust_region_radius == 1e16
    assert o.min_trust_region_radius == 1e-32 and o.min_relative_decrease == 1e-3
    assert o.min_lm_diagonal == 1e-6 and o.max_lm_diagonal == 1e32
    assert o.function_tolerance == 1e-6 and o.gradient_tolerance == 1e-10 and o.parameter_tolerance == 1e-8
    assert o.pcg_max_iterations == 500 and o.pcg_eta == 0.1 and o.pcg_r_tolerance == -1.0


def test_pod_sizes_match_header():
    # 10 int32 + 11 double ; 4 int32 + 8 double ; 10 int32 + 2 int64 + 9 double + 2 int32 ; char[32] + int64 + 2 double
    assert C.sizeof(abi.Options) == 10 * 4 + 11 * 8 + 2 * 4 + 8
    assert C.sizeof(abi.KernelStat) == 32 + 8 + 2 * 8
    assert C.sizeof(abi.Iteration) == 4 * 4 + 8 * 8
    assert C.sizeof(abi.Summary) == 10 * 4 + 2 * 8 + 9 * 8 + 2 * 4


def test_no_cpu_fallback_without_device():
    if cuda.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(cuda.BAError) as e:
        cuda.Problem(0)
    assert "no CPU fallback" in str(e.value)


def test_shard_blocks_partition():
    rng = np.random.default_rng(0)
    w = rng.integers(1, 17, 10000)
    for world in (1, 2, 3, 4, 8):
        r = cuda.shard_blocks(w, world)
        assert r[0] == 0 and r[-1] == w.shape[0] and np.all(np.diff(r) >= 0)
        loads = np.array([w[r[i]:r[i + 1]].sum() for i in range(world)])
        assert loads.sum() == w.sum()
        assert loads.max() - loads.min() <= 2 * w.max()
    assert list(cuda.shard_blocks(np.zeros(0, np.int64), 2)) == [0, 0, 0]
