"""Pins the CPU oracle to the reference's own committed outputs:
G1 hongo/Camera_Transform.xml (+ Extrinsics/mat*.txt, hongo/point3d.txt) and
G2 test2/Camera_Transform.xml (+ test2/point3d.txt), plus the known-answer traces."""
import os

import numpy as np
import pytest

from realsensecalibration_b200 import abi, formats as F
from tests import helpers as H


@pytest.mark.parametrize("solver", [0, 1])
def test_hongo_golden(oracle, solver):
    pb, intr, side, fix0 = H.hongo()
    x, s, rows = oracle.solve_model_b(pb, intr, side, fix0, linear_solver=solver)
    H.check_hongo_golden(x, 5e-15, 5e-15)
    assert s.termination_type == abi.CONVERGENCE and s.termination_reason == abi.REASON_FUNCTION_TOLERANCE
    assert s.num_iterations == 7 and len(rows) == 7
    assert s.num_free_parameters == 114 and s.num_residuals == 544 and s.rcs_dim == 6 * 13
    for r, c, rad in zip(rows, H.HONGO_COSTS, H.HONGO_RADII):
        assert H.rel(r["cost"], c) < 1e-11
        assert H.rel(r["trust_region_radius"], rad) < 1e-9
    assert H.rel(rows[0]["gradient_max_norm"], 2.740449918e+06) < 1e-9
    # camera 0 and marker 0 are not parameter blocks under the Main dispatch: untouched
    assert np.all(x[:6] == pb.params[:6])
    m0 = 6 * (pb.n_cam + pb.n_time)
    assert np.all(x[m0:m0 + 6] == pb.params[m0:m0 + 6])


def test_hongo_outputs_match_committed_files(oracle, golden_dir):
    pb, intr, side, fix0 = H.hongo()
    x, _, _ = oracle.solve_model_b(pb, intr, side, fix0)
    rot, inv, corners = oracle.model_b_outputs(pb, x, side)
    for c in range(4):
        ext = F.load_extrinsics(os.path.join(golden_dir, "Calibration", "Extrinsics", "mat%d.txt" % c))
        assert np.abs(inv[c] - ext).max() < 6e-7   # 6 significant digits
    counts, pts = F.load_point3d(os.path.join(golden_dir, "Correspondence", "hongo", "point3d.txt"))
    assert pts.shape == corners.shape == (272, 3)
    assert np.abs(pts - corners).max() < 6e-7
    assert np.array_equal(counts, pb.counts * 4)       # accessor returns x4, bundle_adjustment.cpp:29-32
    # the points really are post-BA: the initial parameters are far away
    _, _, c0 = oracle.model_b_outputs(pb, pb.params, side)
    assert np.abs(pts - c0).max() > 1e-2


@pytest.mark.parametrize("solver", [0, 1])
def test_test2_golden(oracle, solver, golden_dir):
    pb, intr, side, fix0 = H.test2()
    x, s, rows = oracle.solve_model_b(pb, intr, side, fix0, linear_solver=solver)
    H.check_test2_golden(x, 5e-15)
    assert s.termination_reason == abi.REASON_FUNCTION_TOLERANCE and s.num_iterations == 4
    assert s.num_free_parameters == 54 and s.num_residuals == 160
    for r, c, rad in zip(rows, H.TEST2_COSTS, H.TEST2_RADII):
        assert H.rel(r["cost"], c) < 1e-11 and H.rel(r["trust_region_radius"], rad) < 1e-9
    _, _, corners = oracle.model_b_outputs(pb, x, side)
    _, pts = F.load_point3d(os.path.join(golden_dir, "Correspondence", "test2", "point3d.txt"))
    assert np.abs(pts - corners).max() < 6e-7


def test_test2_golden_discriminates_variant(oracle):
    # fixing marker 0 (Main dispatch) on the Test2 fixture misses the golden by ~3e-8
    pb, intr, side, _ = H.test2()
    x, _, _ = oracle.solve_model_b(pb, intr, side, 1)
    gold = F.load_opencv_xml(os.path.join(H.GOLDEN, "Correspondence", "test2", "Camera_Transform.xml"))
    assert np.abs(x[6:9] - gold["R1"].ravel()).max() > 1e-9


@pytest.mark.parametrize("solver", [0, 1])
def test_two_cam_trace(oracle, solver):
    pa, intr = H.two_cam()
    x, s, rows = oracle.solve_model_a(pa.n_cam, pa.n_pt, pa.cam_idx, pa.pt_idx, pa.obs_xy, intr, pa.params,
                                      linear_solver=solver)
    assert s.termination_reason == abi.REASON_PARAMETER_TOLERANCE and s.num_iterations == 3
    assert H.rel(rows[0]["cost"], H.TWO_CAM_COSTS[0]) < 1e-11
    assert H.rel(rows[1]["cost"], H.TWO_CAM_COSTS[1]) < 1e-8
    assert rows[2]["cost"] < 1e-10


def test_solver_paths_agree(oracle):
    pb, intr, side, fix0 = H.hongo()
    xs = [oracle.solve_model_b(pb, intr, side, fix0, linear_solver=k)[0] for k in (0, 1)]
    assert np.abs(xs[0] - xs[1]).max() < 1e-11
    opt = oracle.default_options()
    opt.pcg_eta = 0.0; opt.pcg_r_tolerance = 1e-14; opt.pcg_max_iterations = 2000
    x2, s2, rows2 = oracle.solve_model_b(pb, intr, side, fix0, linear_solver=2, options=opt)
    assert s2.num_iterations == 7 and np.abs(xs[1] - x2).max() < 1e-8
    assert all(r["linear_solver_iterations"] > 0 for r in rows2[1:])


def test_jet_jacobian_vs_complex_step(oracle):
    """Independent check of the dual-number Jacobian: numpy complex-step twin of the functors."""
    def rot(w, p):
        t2 = w @ w
        if t2.real > np.finfo(float).eps:
            t = np.sqrt(t2); k = w / t
            return p * np.cos(t) + np.cross(k, p) * np.sin(t) + k * (k @ p) * (1 - np.cos(t))
        return p + np.cross(w, p)

    pb, intr, side, fix0 = H.hongo()
    _, res, jac = oracle.eval_model_b(pb, intr, side, fix0)
    h = side / 2
    corners = np.array([[-h, h, 0], [h, h, 0], [h, -h, 0], [-h, -h, 0]])
    C_, T_ = pb.n_cam, pb.n_time

    def f(o, cam, fr, mk):
        out = []
        K = intr[pb.cam_idx[o]]
        for j in range(4):
            p = corners[j].astype(complex)
            if not (fix0 and pb.marker_idx[o] == 0):
                p = rot(mk[:3], p) + mk[3:]
            p = rot(fr[:3], p) + fr[3:]
            if pb.cam_idx[o] != 0:
                p = rot(cam[:3], p) + cam[3:]
            out += [K[0] * p[0] / p[2] + K[2] - pb.obs8[o, 2 * j], K[1] * p[1] / p[2] + K[3] - pb.obs8[o, 2 * j + 1]]
        return np.array(out)

    for o in range(pb.n_mobs):
        blocks = [pb.params[6 * pb.cam_idx[o]:][:6], pb.params[6 * (C_ + pb.time_idx[o]):][:6],
                  pb.params[6 * (C_ + T_ + pb.marker_idx[o]):][:6]]
        r0 = f(o, *[b.astype(complex) for b in blocks]).real
        assert np.abs(r0 - res[o]).max() < 1e-9
        active = [pb.cam_idx[o] != 0, True, not (fix0 and pb.marker_idx[o] == 0)]
        for bi in range(3):
            J = jac[o, 48 * bi:48 * bi + 48].reshape(8, 6)
            if not active[bi]:
                assert np.all(J == 0)
                continue
            for k in range(6):
                args = [b.astype(complex) for b in blocks]
                args[bi][k] += 1e-30j
                d = f(o, *args).imag / 1e-30
                assert np.abs(d - J[:, k]).max() <= 1e-9 * max(1.0, np.abs(d).max())


def test_model_a_pinned_independently_of_the_cpp_oracle(oracle):
    """Model A (Test1_BundleAdjustment/bundle_adjustmenter.cpp:122-141) pinned without the C++ oracle's own arithmetic:
    the projection by cv2.projectPoints, the Jacobian by a numpy complex-step twin of the functor, and the whole
    trust-region loop by a dense numpy restatement of Ceres' LM (Jacobi scaling, D = sqrt(clamp(diag) / radius), the
    cubic radius update) on Common/Correspondence/two_cam_data.txt.  The C++ oracle (Schur path) must reproduce all three."""
    cv2 = pytest.importorskip("cv2")
    pa, intr = H.two_cam()
    K = np.array([[intr[0], 0, intr[2]], [0, intr[1], intr[3]], [0, 0, 1.0]])
    n_cam, n_pt, nobs = pa.n_cam, pa.n_pt, len(pa.cam_idx)
    obs = np.asarray(pa.obs_xy, np.float64).reshape(-1, 2)

    def rot(w, p):
        t2 = w @ w
        if t2.real > np.finfo(float).eps:
            t = np.sqrt(t2); k = w / t
            return p * np.cos(t) + np.cross(k, p) * np.sin(t) + k * (k @ p) * (1 - np.cos(t))
        return p + np.cross(w, p)

    def residuals(x):
        out = np.zeros(2 * nobs, dtype=x.dtype)
        for o in range(nobs):
            cam = x[6 * pa.cam_idx[o]:][:6]
            X = x[6 * n_cam + 3 * pa.pt_idx[o]:][:3]
            p = rot(cam[:3], X) + cam[3:]
            out[2 * o] = intr[0] * p[0] / p[2] + intr[2] - obs[o, 0]
            out[2 * o + 1] = intr[1] * p[1] / p[2] + intr[3] - obs[o, 1]
        return out

    def jacobian(x):
        J = np.zeros((2 * nobs, x.size))
        for k in range(x.size):
            xc = x.astype(complex); xc[k] += 1e-30j
            J[:, k] = residuals(xc).imag / 1e-30
        return J

    x0 = np.asarray(pa.params, np.float64)
    # (1) the projection: OpenCV's pinhole model, no distortion
    cost_o, res_o, jac_o = oracle.eval_model_a(n_cam, n_pt, pa.cam_idx, pa.pt_idx, pa.obs_xy, intr, x0)
    for o in range(nobs):
        cam = x0[6 * pa.cam_idx[o]:][:6]
        X = x0[6 * n_cam + 3 * pa.pt_idx[o]:][:3]
        uv, _ = cv2.projectPoints(X.reshape(1, 1, 3), cam[:3].copy(), cam[3:].copy(), K, None)
        assert np.abs(uv.ravel() - obs[o] - res_o[o]).max() < 1e-9
    # (2) the Jacobian: complex step against the oracle's dual numbers
    J0 = jacobian(x0)
    assert np.abs(residuals(x0) - res_o.ravel()).max() < 1e-10
    for o in range(nobs):
        jc = J0[2 * o:2 * o + 2, 6 * pa.cam_idx[o]:][:, :6]
        jp = J0[2 * o:2 * o + 2, 6 * n_cam + 3 * pa.pt_idx[o]:][:, :3]
        mine = np.concatenate([jc.ravel(), jp.ravel()])
        assert np.abs(mine - jac_o[o]).max() <= 1e-9 * max(1.0, np.abs(mine).max()), o
    # (3) the trust-region loop, dense normal equations in numpy
    opt = oracle.default_options()
    x = x0.copy()
    radius, decrease = opt.initial_trust_region_radius, 2.0
    r, J = residuals(x), J0
    scale = 1.0 / (1.0 + np.sqrt((J * J).sum(axis=0)))
    cost = 0.5 * r @ r
    costs, reason = [cost], None
    for it in range(opt.max_num_iterations):
        Js = J * scale
        H_ = Js.T @ Js
        D2 = np.clip(np.diag(H_), opt.min_lm_diagonal, opt.max_lm_diagonal) / radius
        delta = np.linalg.solve(H_ + np.diag(D2), -(Js.T @ r))
        Jd = Js @ delta
        model_change = -(Jd @ (r + Jd / 2))
        assert model_change > 0
        xn = x + scale * delta
        if np.linalg.norm(scale * delta) <= opt.parameter_tolerance * (np.linalg.norm(x) + opt.parameter_tolerance):
            reason = abi.REASON_PARAMETER_TOLERANCE
            break
        rn = residuals(xn)
        cost_n = 0.5 * rn @ rn
        if abs(cost - cost_n) <= opt.function_tolerance * cost:
            reason = abi.REASON_FUNCTION_TOLERANCE
            break
        rho = (cost - cost_n) / model_change
        if rho > opt.min_relative_decrease:
            x, r, cost = xn, rn, cost_n
            J = jacobian(x)
            radius = min(opt.max_trust_region_radius, radius / max(1.0 / 3.0, 1.0 - (2.0 * rho - 1.0) ** 3))
            decrease = 2.0
            costs.append(cost)
        else:
            radius /= decrease; decrease *= 2.0
            costs.append(cost_n)
    xo, so, rows_o = oracle.solve_model_a(n_cam, n_pt, pa.cam_idx, pa.pt_idx, pa.obs_xy, intr, pa.params)
    assert reason == so.termination_reason == abi.REASON_PARAMETER_TOLERANCE
    assert len(costs) == so.num_iterations == 3
    assert H.rel(costs[0], rows_o[0]["cost"]) < 1e-12 and H.rel(costs[0], H.TWO_CAM_COSTS[0]) < 1e-11
    assert H.rel(costs[1], rows_o[1]["cost"]) < 1e-7 and H.rel(costs[1], H.TWO_CAM_COSTS[1]) < 1e-7
    assert costs[2] < 1e-10 and rows_o[2]["cost"] < 1e-10
    assert np.abs(x - xo).max() < 1e-6
