"""The source-compatible C++ host classes (realsensecalibration_b200/host: BALProblem, BAManager,
ReprojectionCheck, Test1's BALProblem) driven the way Main_Calibration/main.cpp:35-43 drives the reference's."""
import os
import re
import subprocess

import numpy as np
import pytest

from realsensecalibration_b200 import cuda, formats as F
from tests import helpers as H

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "realsensecalibration_b200", "host")
EXE = os.path.join(HOST, "ba_main_calibration")


def _build():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "realsensecalibration_b200", "csrc")])
    subprocess.check_call(["make", "-s", "-C", HOST])


def test_host_library_builds_and_keeps_the_reference_interface():
    _build()
    assert os.path.exists(os.path.join(HOST, "libba_host.so")) and os.path.exists(EXE)
    sym = subprocess.run(["nm", "-DC", os.path.join(HOST, "libba_host.so")], capture_output=True, text=True).stdout
    for s in ["RSCalibration::BALProblem::loadFile(char const*)", "RSCalibration::BALProblem::getPoint3dCoordinates(",
              "RSCalibration::BALProblem::mutable_marker_transform_from_base_marker(int)",
              "RSCalibration::BALProblem::num_observations_per_time_camera(int, int) const",
              "RSCalibration::BAManager::StartBA()", "RSCalibration::BAManager::Write()",
              "RSCalibration::BAManager::BAManager(std::map<", "RSCalibration::ReprojectionCheck::Reproject(std::vector<std::map<"]:
        assert s in sym, s


def test_host_fails_loudly_without_a_gpu(tmp_path):
    if cuda.device_count() > 0:
        pytest.skip("a GPU is present")
    _build()
    r = subprocess.run([EXE, H.GOLDEN, str(tmp_path)], capture_output=True, text=True)
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)


@pytest.mark.gpu
def test_main_calibration_ba_stage_reproduces_the_committed_outputs(tmp_path):
    r = subprocess.run([EXE, H.GOLDEN, str(tmp_path)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    out = r.stdout
    # Ceres' progress table, 7 rows, same costs
    costs = [float(m.group(1)) for m in re.finditer(r"^\s*\d+\s+([0-9.]+e[+-]\d+)\s", out, re.M)]
    assert len(costs) == 7
    for c, g in zip(costs, H.HONGO_COSTS):
        assert abs(c - g) <= 1e-6 * g          # printed with 6 digits
    m = re.search(r"summary: iterations (\d+) initial (\S+) final (\S+) reprojection (\S+) rms (\S+)", out)
    assert m and int(m.group(1)) == 7
    assert H.rel(float(m.group(2)), H.HONGO_COSTS[0]) < 1e-12 and H.rel(float(m.group(3)), H.HONGO_COSTS[-1]) < 1e-10
    assert abs(float(m.group(4)) - 143.63) < 0.5 and abs(float(m.group(5)) - 0.7267) < 2e-3
    assert "Reprojection Error (After BA): " in out and "Average Reprojection Error per One Coordinate: " in out
    # Camera_Transform.xml: R as 3x3, 17 digits (bundle_adjustment_manager.cpp:130-131) vs the reference's committed file
    got = F.load_opencv_xml(os.path.join(tmp_path, "Camera_Transform.xml"))
    gold = F.load_opencv_xml(os.path.join(H.GOLDEN, "Correspondence", "hongo", "Camera_Transform.xml"))
    for c in range(4):
        assert got["R%d" % c].shape == (3, 3) and got["t%d" % c].shape == (3, 1)
        assert np.abs(got["R%d" % c] - gold["R%d" % c]).max() < 1e-10 and np.abs(got["t%d" % c] - gold["t%d" % c]).max() < 1e-10
        ext = F.load_extrinsics(os.path.join(tmp_path, "mat%d.txt" % c))
        ext_gold = F.load_extrinsics(os.path.join(H.GOLDEN, "Calibration", "Extrinsics", "mat%d.txt" % c))
        assert np.abs(ext - ext_gold).max() < 2e-6
    (hdr, pts), (hdr_g, pts_g) = (F.load_point3d(os.path.join(tmp_path, "point3d.txt")),
                                  F.load_point3d(os.path.join(H.GOLDEN, "Correspondence", "hongo", "point3d.txt")))
    assert pts.shape == pts_g.shape == (272, 3) and np.abs(pts - pts_g).max() < 2e-6
    # the text files are byte-identical to the reference's where the digits allow (same ofstream default formatting)
    same = sum(a == b for a, b in zip(open(os.path.join(tmp_path, "point3d.txt")).read().split(),
                                      open(os.path.join(H.GOLDEN, "Correspondence", "hongo", "point3d.txt")).read().split()))
    assert same >= 0.98 * (272 * 3)


@pytest.mark.gpu
def test_test1_program_path(tmp_path, oracle):
    r = subprocess.run([EXE, H.GOLDEN, str(tmp_path), "test1"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    m = re.search(r"test1: iterations (\d+) initial (\S+) final (\S+) camera (.*)", r.stdout)
    assert m and int(m.group(1)) == 3
    assert H.rel(float(m.group(2)), H.TWO_CAM_COSTS[0]) < 1e-10 and float(m.group(3)) < 1e-10
    pa, intr = H.two_cam()
    xo, so, _ = oracle.solve_model_a(pa.n_cam, pa.n_pt, pa.cam_idx, pa.pt_idx, pa.obs_xy, intr, pa.params)
    cam = np.array([float(v) for v in m.group(4).split()])
    assert np.abs(cam - xo[:6]).max() < 1e-6


@pytest.mark.gpu
def test_test2_flow_writes_rvecs_and_reprojects_from_them(tmp_path, oracle):
    # Test2_BundleAdjustment's flow: two-functor dispatch, Camera_Transform.xml with R<i> as the 3x1 rotation vector
    # (Test2_BundleAdjustment/main.cpp:128), then ReprojectionCheck::Reproject reading that file back
    r = subprocess.run([EXE, H.GOLDEN, str(tmp_path), "test2"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    got = F.load_opencv_xml(os.path.join(tmp_path, "Camera_Transform.xml"))
    gold = F.load_opencv_xml(os.path.join(H.GOLDEN, "Correspondence", "test2", "Camera_Transform.xml"))
    for c in range(2):
        assert got["R%d" % c].shape == (3, 1)
        assert np.abs(got["R%d" % c] - gold["R%d" % c]).max() < 1e-10 and np.abs(got["t%d" % c] - gold["t%d" % c]).max() < 1e-10
    m = re.search(r"summary: iterations (\d+) initial (\S+) final (\S+) reprojection (\S+) rms (\S+)", r.stdout)
    assert m and int(m.group(1)) == 4
    # the check projects the 6-digit points of point3d.txt with float image points: close to the final cost of the solve
    pb, intr, side, fix0 = H.test2()
    xo, so, rows_o = oracle.solve_model_b(pb, intr, side, fix0)
    assert H.rel(float(m.group(3)), so.final_cost) < 1e-10
    assert abs(float(m.group(4)) - so.final_cost) < 0.02 * so.final_cost
