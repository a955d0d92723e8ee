"""Readers / writers of the reference's on-disk formats around the hot path.

  correspondence.txt   writer Main_Calibration/correspondencer.cpp:207-282,
                       reader Main_Calibration/bundle_adjustment.cpp:132-187
  two_cam_data.txt     writer Test1_ReprojectionError/main.cpp:162-183,
                       reader Test1_BundleAdjustment/bundle_adjustmenter.cpp:55-85
  point3d.txt          writer Main_Calibration/bundle_adjustment_manager.cpp:157-174,
                       reader Main_Calibration/reprojection_check.cpp:7-62
  Camera_Transform.xml / Intrinsics/<serial>.xml   OpenCV FileStorage XML
                       (bundle_adjustment_manager.cpp:108-131, my_io.cpp:19-23)
  Extrinsics/mat<i>.txt  bundle_adjustment_manager.cpp:135-149

Host-side file plumbing only (numpy); nothing here is on the compute path.
"""
import re
from dataclasses import dataclass

import numpy as np

# Main_Calibration/my_const.h:9-16
MARKER_SIDE = 0.0148
SERIAL_NUMBERS = ("821312061029", "816612062327", "821212062536", "821212061326")


@dataclass
class ModelBFile:
    n_time: int
    n_cam: int
    n_marker: int
    counts: np.ndarray      # [T, C] marker observations per (time, camera)
    time_idx: np.ndarray    # int32 [N]
    cam_idx: np.ndarray
    marker_idx: np.ndarray
    obs8: np.ndarray        # float64 [N, 8]
    params: np.ndarray      # float64 [6 * (C + T + M)]

    @property
    def n_mobs(self):
        return int(self.time_idx.shape[0])


@dataclass
class ModelAFile:
    n_cam: int
    n_pt: int
    cam_idx: np.ndarray
    pt_idx: np.ndarray
    obs_xy: np.ndarray      # float64 [N, 2]
    params: np.ndarray      # float64 [6 * n_cam + 3 * n_pt]


def _tokens(path):
    with open(path, "r") as f:
        return f.read().split()


def load_correspondence(path):
    """BALProblem::loadFile, Main_Calibration/bundle_adjustment.cpp:132-187."""
    tok = _tokens(path)
    T, Cn, M, N = (int(t) for t in tok[:4])
    pos = 4
    counts = np.zeros((T, Cn), np.int32)
    for t in range(T):
        pos += 1  # the time index itself is skipped ("tmp")
        for c in range(Cn):
            counts[t, c] = int(tok[pos]); pos += 1
    ti = np.zeros(N, np.int32); ci = np.zeros(N, np.int32); mi = np.zeros(N, np.int32)
    obs = np.zeros((N, 8), np.float64)
    for i in range(N):
        ti[i], ci[i], mi[i] = int(tok[pos]), int(tok[pos + 1]), int(tok[pos + 2]); pos += 3
        obs[i] = [float(x) for x in tok[pos:pos + 8]]; pos += 8
    n_par = 6 * (Cn + T + M)
    params = np.array([float(x) for x in tok[pos:pos + n_par]], np.float64)
    if params.shape[0] != n_par:
        raise ValueError("short correspondence file %s" % path)
    return ModelBFile(T, Cn, M, counts, ti, ci, mi, obs, params)


def write_correspondence(path, pb):
    """Correspondencer::Write layout, Main_Calibration/correspondencer.cpp:207-282."""
    with open(path, "w") as f:
        f.write("%d %d %d %d\n" % (pb.n_time, pb.n_cam, pb.n_marker, pb.n_mobs))
        for t in range(pb.n_time):
            f.write(" ".join([str(t)] + [str(int(v)) for v in pb.counts[t]]) + "\n")
        for i in range(pb.n_mobs):
            f.write("%d %d %d " % (pb.time_idx[i], pb.cam_idx[i], pb.marker_idx[i]))
            f.write(" ".join(repr(float(v)) for v in pb.obs8[i]) + "\n")
        for row in pb.params.reshape(-1, 6):
            f.write(" ".join(repr(float(v)) for v in row) + "\n")


def load_two_cam_data(path):
    """BALProblem::LoadFile, Test1_BundleAdjustment/bundle_adjustmenter.cpp:55-85
    (num_observations == num_points, :64)."""
    tok = _tokens(path)
    n_cam, n_pt = int(tok[0]), int(tok[1])
    pos = 2
    ci = np.zeros(n_pt, np.int32); pi = np.zeros(n_pt, np.int32); obs = np.zeros((n_pt, 2))
    for i in range(n_pt):
        ci[i], pi[i] = int(tok[pos]), int(tok[pos + 1])
        obs[i] = (float(tok[pos + 2]), float(tok[pos + 3])); pos += 4
    n_par = 6 * n_cam + 3 * n_pt
    params = np.array([float(x) for x in tok[pos:pos + n_par]], np.float64)
    if params.shape[0] != n_par:
        raise ValueError("short two_cam_data file %s" % path)
    return ModelAFile(n_cam, n_pt, ci, pi, obs, params)


def write_two_cam_data(path, pa):
    with open(path, "w") as f:
        f.write("%d %d\n" % (pa.n_cam, pa.n_pt))
        for i in range(pa.cam_idx.shape[0]):
            f.write("%d %d %r %r\n" % (pa.cam_idx[i], pa.pt_idx[i], float(pa.obs_xy[i, 0]), float(pa.obs_xy[i, 1])))
        for v in pa.params:
            f.write("%r\n" % float(v))


_MAT_RE = re.compile(r"<(\w+) type_id=\"opencv-matrix\">\s*<rows>(\d+)</rows>\s*<cols>(\d+)</cols>\s*"
                     r"<dt>(\w+)</dt>\s*<data>(.*?)</data>", re.S)


def load_opencv_xml(path):
    """Minimal reader of OpenCV FileStorage XML holding dense matrices."""
    with open(path, "r") as f:
        text = f.read()
    out = {}
    for name, rows, cols, _dt, data in _MAT_RE.findall(text):
        out[name] = np.array([float(x) for x in data.split()], np.float64).reshape(int(rows), int(cols))
    return out


def _fmt17(v):
    # cv::FileStorage writes doubles with "%.17g" and a trailing '.' for integers
    s = "%.17g" % v
    if "e" in s:
        m, e = s.split("e")
        if "." not in m:
            m += "."
        return "%se%+03d" % (m, int(e))
    return s if "." in s else s + "."


def write_opencv_xml(path, mats):
    with open(path, "w") as f:
        f.write('<?xml version="1.0"?>\n<opencv_storage>\n')
        for name, m in mats.items():
            m = np.asarray(m, np.float64)
            f.write('<%s type_id="opencv-matrix">\n  <rows>%d</rows>\n  <cols>%d</cols>\n  <dt>d</dt>\n  <data>\n    ' %
                    (name, m.shape[0], m.shape[1]))
            f.write(" ".join(_fmt17(float(v)) for v in m.ravel()))
            f.write("</data></%s>\n" % name)
        f.write("</opencv_storage>\n")


def load_intrinsics(path):
    """IO::GetIntrinsics, Main_Calibration/my_io.cpp:19-23 -> (fx, fy, ppx, ppy), distCoeffs."""
    m = load_opencv_xml(path)
    K = m["intrinsics"]
    return np.array([K[0, 0], K[1, 1], K[0, 2], K[1, 2]], np.float64), m["distCoeffs"].ravel()


def load_point3d(path):
    """reprojection_check.cpp:15-33,55-62."""
    tok = _tokens(path)
    n, T, Cn = int(tok[0]), int(tok[1]), int(tok[2])
    pos = 3
    counts = np.zeros((T, Cn), np.int32)
    for t in range(T):
        pos += 1
        for c in range(Cn):
            counts[t, c] = int(tok[pos]); pos += 1
    pts = np.array([float(x) for x in tok[pos:pos + 3 * n]], np.float64).reshape(n, 3)
    return counts, pts


def _fmt6(v):
    # default std::ofstream << double : "%g" with 6 significant digits
    return "%g" % v


def write_point3d(path, counts_x4, pts):
    with open(path, "w") as f:
        T, Cn = counts_x4.shape
        f.write("%d %d %d\n" % (pts.shape[0], T, Cn))
        for t in range(T):
            f.write(" ".join([str(t)] + [str(int(v)) for v in counts_x4[t]]) + "\n")
        for p in pts:
            f.write("%s %s %s\n" % (_fmt6(p[0]), _fmt6(p[1]), _fmt6(p[2])))


def load_extrinsics(path):
    return np.array([float(x) for x in _tokens(path)], np.float64).reshape(3, 4)


def write_extrinsics(path, inv12):
    with open(path, "w") as f:
        for v in np.asarray(inv12).ravel():
            f.write(_fmt6(float(v)) + "\n")
