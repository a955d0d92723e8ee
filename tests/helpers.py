"""Shared fixtures-as-functions for the parity tests (host-side numpy only)."""
import os

import numpy as np

from realsensecalibration_b200 import formats as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden", "reference_common")

# Known-answer traces of the reference's fixtures (SURVEY.md 8c / BASELINE.md 2): the
# end states match the reference's committed Ceres outputs to 1e-15.
HONGO_COSTS = [1.387966960542683e+05, 1.044261171521932e+05, 1.619009257909599e+04, 1.140581757675756e+04,
               5.783243408891863e+02, 1.441617307924972e+02, 1.436293888518609e+02]
HONGO_RADII = [1e4, 9.589347493e3, 2.876804248e4, 2.770715625e4, 8.312146876e4, 2.493644063e5, 7.480932188e5]
TEST2_COSTS = [1.357816158655430e+02, 1.343560334876707e+01, 1.330172662558997e+01, 1.330170910708654e+01]
TEST2_RADII = [1e4, 3e4, 9e4, 2.7e5]
TWO_CAM_COSTS = [5.617747098021e+00, 2.512595816529e-04, 2.734342332794e-12]
# G2's configuration is no longer in the reference source tree
TEST2_MARKER_SIDE = 0.048
TEST2_SERIALS = ("819612072493", "825312072048")
TEST1_SERIAL = "825312072048"  # serial_numbers[1], Test1_BundleAdjustment/main.cpp:21,73


def intrinsics(serials):
    return np.stack([F.load_intrinsics(os.path.join(GOLDEN, "Calibration", "Intrinsics", "%s.xml" % s))[0] for s in serials])


def hongo():
    pb = F.load_correspondence(os.path.join(GOLDEN, "Correspondence", "hongo", "correspondence.txt"))
    return pb, intrinsics(F.SERIAL_NUMBERS), F.MARKER_SIDE, 1


def test2():
    pb = F.load_correspondence(os.path.join(GOLDEN, "Correspondence", "test2", "correspondence_test.txt"))
    return pb, intrinsics(TEST2_SERIALS), TEST2_MARKER_SIDE, 0


def two_cam():
    pa = F.load_two_cam_data(os.path.join(GOLDEN, "Correspondence", "two_cam_data.txt"))
    return pa, intrinsics((TEST1_SERIAL,))[0]


def rodrigues_np(rvec):
    """cv::Rodrigues rvec -> R, numpy twin used to compare against XML goldens."""
    r = np.asarray(rvec, np.float64)
    t = np.linalg.norm(r)
    if t < np.finfo(np.float64).eps:
        return np.eye(3)
    k = r / t
    K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.cos(t) * np.eye(3) + (1 - np.cos(t)) * np.outer(k, k) + np.sin(t) * K


def check_hongo_golden(x, tol_R=5e-14, tol_t=5e-14):
    gold = F.load_opencv_xml(os.path.join(GOLDEN, "Correspondence", "hongo", "Camera_Transform.xml"))
    for c in range(4):
        assert np.abs(rodrigues_np(x[6 * c:6 * c + 3]) - gold["R%d" % c]).max() <= tol_R, c
        assert np.abs(x[6 * c + 3:6 * c + 6] - gold["t%d" % c].ravel()).max() <= tol_t, c


def check_test2_golden(x, tol=5e-14):
    gold = F.load_opencv_xml(os.path.join(GOLDEN, "Correspondence", "test2", "Camera_Transform.xml"))
    for c in range(2):
        assert np.abs(x[6 * c:6 * c + 3] - gold["R%d" % c].ravel()).max() <= tol, c
        assert np.abs(x[6 * c + 3:6 * c + 6] - gold["t%d" % c].ravel()).max() <= tol, c


def rel(a, b):
    return abs(a - b) / max(abs(b), 1e-300)
