"""ctypes binding of oracle/libba_oracle.so -- TEST INFRASTRUCTURE ONLY.

Importable only from tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from realsensecalibration_b200.abi import Iteration, Options, Summary

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

DENSE_NORMAL, SCHUR_DENSE, SCHUR_PCG = 0, 1, 2

_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libba_oracle.so")
        if not os.path.exists(path):
            build()
        _LIB = C.CDLL(path)
        _LIB.ba_oracle_options_init.argtypes = [C.POINTER(Options)]
        _LIB.ba_oracle_max_threads.restype = C.c_int
    return _LIB


def default_options():
    o = Options()
    lib().ba_oracle_options_init(C.byref(o))
    return o


def max_threads():
    return int(lib().ba_oracle_max_threads())


def _opt(v, dtype):
    return None if v is None else v.ctypes.data_as(C.c_void_p)


def _intr_stride(intr, n_cam):
    if intr.size == 4 * n_cam:
        return 4
    if intr.size == 4:
        return 0  # one set shared by all cameras (Test1_BundleAdjustment/main.cpp:73-74)
    raise ValueError("intr must hold 4 or 4*n_cam doubles")


def _rows(buf, n):
    return [buf[i].as_dict() for i in range(n)]


def solve_model_a(n_cam, n_pt, cam_idx, pt_idx, obs_xy, intr, params, options=None, linear_solver=SCHUR_DENSE,
                  n_threads=1, cap=256):
    cam_idx = np.ascontiguousarray(cam_idx, np.int32); pt_idx = np.ascontiguousarray(pt_idx, np.int32)
    obs_xy = np.ascontiguousarray(obs_xy, np.float64); intr = np.ascontiguousarray(intr, np.float64)
    x = np.array(params, np.float64, copy=True)
    stride = _intr_stride(intr, n_cam)
    opts = options or default_options()
    s = Summary(); rows = (Iteration * cap)(); n = C.c_int(0)
    rc = lib().ba_oracle_solve_model_a(
        C.c_int32(n_cam), C.c_int64(n_pt), C.c_int64(cam_idx.shape[0]), cam_idx.ctypes, pt_idx.ctypes, obs_xy.ctypes,
        intr.ctypes, C.c_int32(stride), x.ctypes, C.byref(opts), C.c_int(linear_solver), C.c_int(n_threads),
        C.byref(s), rows, C.c_int(cap), C.byref(n))
    if rc != 0:
        raise ValueError("oracle rejected the Model A problem")
    return x, s, _rows(rows, min(n.value, cap))


def solve_model_b(pb, intr4, marker_side, fix_marker0, params=None, options=None, linear_solver=SCHUR_DENSE,
                  n_threads=1, cap=256):
    x = np.array(pb.params if params is None else params, np.float64, copy=True)
    intr4 = np.ascontiguousarray(intr4, np.float64)
    obs8 = np.ascontiguousarray(pb.obs8, np.float64)
    ti = np.ascontiguousarray(pb.time_idx, np.int32); ci = np.ascontiguousarray(pb.cam_idx, np.int32)
    mi = np.ascontiguousarray(pb.marker_idx, np.int32)
    opts = options or default_options()
    s = Summary(); rows = (Iteration * cap)(); n = C.c_int(0)
    rc = lib().ba_oracle_solve_model_b(
        C.c_int32(pb.n_cam), C.c_int32(pb.n_time), C.c_int32(pb.n_marker), C.c_int64(pb.n_mobs), ti.ctypes, ci.ctypes,
        mi.ctypes, obs8.ctypes, intr4.ctypes, C.c_double(marker_side), C.c_int32(1), C.c_int32(int(fix_marker0)),
        x.ctypes, C.byref(opts), C.c_int(linear_solver), C.c_int(n_threads), C.byref(s), rows, C.c_int(cap), C.byref(n))
    if rc != 0:
        raise ValueError("oracle rejected the Model B problem")
    return x, s, _rows(rows, min(n.value, cap))


def eval_model_a(n_cam, n_pt, cam_idx, pt_idx, obs_xy, intr, params, n_threads=1, want_jac=True):
    cam_idx = np.ascontiguousarray(cam_idx, np.int32); pt_idx = np.ascontiguousarray(pt_idx, np.int32)
    obs_xy = np.ascontiguousarray(obs_xy, np.float64); intr = np.ascontiguousarray(intr, np.float64)
    x = np.ascontiguousarray(params, np.float64)
    n = cam_idx.shape[0]
    stride = _intr_stride(intr, n_cam)
    cost = C.c_double(0)
    res = np.zeros((n, 2)); jac = np.zeros((n, 18)) if want_jac else None
    rc = lib().ba_oracle_eval_model_a(
        C.c_int32(n_cam), C.c_int64(n_pt), C.c_int64(n), cam_idx.ctypes, pt_idx.ctypes, obs_xy.ctypes, intr.ctypes,
        C.c_int32(stride), x.ctypes, C.c_int(n_threads), C.byref(cost), res.ctypes, _opt(jac, np.float64))
    if rc != 0:
        raise ValueError("oracle rejected the Model A problem")
    return cost.value, res, jac


def eval_model_b(pb, intr4, marker_side, fix_marker0, params=None, n_threads=1, want_jac=True):
    x = np.ascontiguousarray(pb.params if params is None else params, np.float64)
    intr4 = np.ascontiguousarray(intr4, np.float64); obs8 = np.ascontiguousarray(pb.obs8, np.float64)
    ti = np.ascontiguousarray(pb.time_idx, np.int32); ci = np.ascontiguousarray(pb.cam_idx, np.int32)
    mi = np.ascontiguousarray(pb.marker_idx, np.int32)
    n = pb.n_mobs
    cost = C.c_double(0)
    res = np.zeros((n, 8)); jac = np.zeros((n, 144)) if want_jac else None
    rc = lib().ba_oracle_eval_model_b(
        C.c_int32(pb.n_cam), C.c_int32(pb.n_time), C.c_int32(pb.n_marker), C.c_int64(n), ti.ctypes, ci.ctypes, mi.ctypes,
        obs8.ctypes, intr4.ctypes, C.c_double(marker_side), C.c_int32(1), C.c_int32(int(fix_marker0)), x.ctypes,
        C.c_int(n_threads), C.byref(cost), res.ctypes, _opt(jac, np.float64))
    if rc != 0:
        raise ValueError("oracle rejected the Model B problem")
    return cost.value, res, jac


def rotate_point(aa, pt):
    aa = np.ascontiguousarray(aa, np.float64); pt = np.ascontiguousarray(pt, np.float64); out = np.zeros(3)
    lib().ba_oracle_angle_axis_rotate_point(aa.ctypes, pt.ctypes, out.ctypes)
    return out


def rodrigues(rvec):
    r = np.ascontiguousarray(rvec, np.float64); R = np.zeros(9)
    lib().ba_oracle_rodrigues(r.ctypes, R.ctypes)
    return R.reshape(3, 3)


def model_b_outputs(pb, params, marker_side):
    x = np.ascontiguousarray(params, np.float64)
    ti = np.ascontiguousarray(pb.time_idx, np.int32); mi = np.ascontiguousarray(pb.marker_idx, np.int32)
    rot = np.zeros((pb.n_cam, 9)); inv = np.zeros((pb.n_cam, 12)); corners = np.zeros((pb.n_mobs * 4, 3))
    lib().ba_oracle_model_b_outputs(C.c_int32(pb.n_cam), C.c_int32(pb.n_time), C.c_int32(pb.n_marker), C.c_int64(pb.n_mobs),
                                    ti.ctypes, mi.ctypes, x.ctypes, C.c_double(marker_side), rot.ctypes, inv.ctypes,
                                    corners.ctypes)
    return rot.reshape(-1, 3, 3), inv.reshape(-1, 3, 4), corners


def project_points_error(xyz, cam_of_point, rvec_tvec6, intr4, image_xy):
    xyz = np.ascontiguousarray(xyz, np.float64); cam = np.ascontiguousarray(cam_of_point, np.int32)
    rt = np.ascontiguousarray(rvec_tvec6, np.float64); K = np.ascontiguousarray(intr4, np.float64)
    img = np.ascontiguousarray(image_xy, np.float32)
    n = xyz.shape[0]
    err = C.c_double(0); rms = C.c_double(0); rep = np.zeros((n, 2))
    lib().ba_oracle_project_points_error(C.c_int64(n), xyz.ctypes, cam.ctypes, rt.ctypes, K.ctypes, img.ctypes,
                                         C.byref(err), C.byref(rms), rep.ctypes)
    return err.value, rms.value, rep
