"""Robust losses (default off): ceres::HuberLoss / ceres::CauchyLoss + Ceres' Corrector, restated in the oracle and
in the CUDA kernels.  The reference passes NULL (bundle_adjustment_manager.cpp:38,51,68,82; Test1 main.cpp:77): with the default
options nothing changes (the goldens are reproduced by the other tests with the very same code path)."""
import numpy as np
import pytest

from realsensecalibration_b200 import abi, cuda, formats as F, synthetic as S
from tests import helpers as H


def _with_outliers(pr, n, seed):
    rng = np.random.default_rng(seed)
    ob = pr.obs_xy.copy()
    idx = rng.choice(ob.shape[0], n, replace=False)
    ob[idx] += rng.normal(0.0, 40.0, (n, 2))
    return ob


def _rho(loss, a, s):
    if loss == abi.LOSS_HUBER:
        return np.where(s > a * a, 2.0 * a * np.sqrt(s) - a * a, s)
    if loss == abi.LOSS_CAUCHY:
        return a * a * np.log1p(s / (a * a))
    return s


@pytest.mark.parametrize("loss", [abi.LOSS_HUBER, abi.LOSS_CAUCHY])
def test_oracle_cost_is_half_the_sum_of_rho(oracle, loss):
    pr = S.bal_like(30, 2000, 5, 12, 5)
    ob = _with_outliers(pr, 60, 1)
    _, res, _ = oracle.eval_model_a(pr.n_cam, pr.n_pt, pr.cam_idx, pr.pt_idx, ob, pr.intr, pr.params, want_jac=False)
    s = (res ** 2).sum(axis=1)
    o = oracle.default_options()
    o.loss_function = loss; o.loss_scale = 1.5; o.max_num_iterations = 0
    _, summ, rows = oracle.solve_model_a(pr.n_cam, pr.n_pt, pr.cam_idx, pr.pt_idx, ob, pr.intr, pr.params, options=o)
    assert (s > 1.5 ** 2).sum() > 40          # the outliers are in the robust regime
    assert H.rel(rows[0]["cost"], 0.5 * _rho(loss, 1.5, s).sum()) < 1e-13
    # Model B: the residual block is the marker observation (8 residuals)
    pb, intr, side, fix0 = H.hongo()
    _, res, _ = oracle.eval_model_b(pb, intr, side, fix0, want_jac=False)
    s = (res ** 2).sum(axis=1)
    o.loss_scale = 10.0
    _, summ, rows = oracle.solve_model_b(pb, intr, side, fix0, options=o)
    assert H.rel(rows[0]["cost"], 0.5 * _rho(loss, 10.0, s).sum()) < 1e-13


def test_a_robust_loss_resists_outliers(oracle):
    # gross outliers (40 px) on 2 % of the observations: the plain least-squares solution moves, the Huber one stays close to
    # the outlier-free solution (compared through the reprojection residuals of the inliers, free of the gauge)
    pr = S.bal_like(30, 2000, 5, 12, 5)
    ob = _with_outliers(pr, 200, 2)
    inl = np.all(ob == pr.obs_xy, axis=1)
    out = {}
    for name, loss in (("none", abi.LOSS_NONE), ("huber", abi.LOSS_HUBER)):
        o = oracle.default_options()
        o.loss_function = loss; o.loss_scale = 2.0; o.max_num_iterations = 25
        x, _, _ = oracle.solve_model_a(pr.n_cam, pr.n_pt, pr.cam_idx, pr.pt_idx, ob, pr.intr, pr.params, options=o, n_threads=4)
        _, res, _ = oracle.eval_model_a(pr.n_cam, pr.n_pt, pr.cam_idx, pr.pt_idx, pr.obs_xy, pr.intr, x, want_jac=False)
        out[name] = np.sqrt((res[inl] ** 2).sum(axis=1).mean())
    assert out["huber"] < 1.0 < out["none"]   # inlier RMS in pixels (the noise is 0.5 px per coordinate)


@pytest.mark.gpu
@pytest.mark.parametrize("loss,env", [(abi.LOSS_HUBER, dict(BA_SA=2)), (abi.LOSS_CAUCHY, dict(BA_SA=2)), (abi.LOSS_HUBER, dict(BA_SA=0)),
                                      (abi.LOSS_CAUCHY, dict(BA_SA=0))])
def test_gpu_model_a_with_loss_matches_oracle(oracle, loss, env):
    import os
    pr = S.bal_like(60, 5000, 6, 16, 13, variable_degree=True)
    ob = _with_outliers(pr, 300, 3)
    opt_g, opt_o = cuda.default_options(), oracle.default_options()
    for o in (opt_g, opt_o):
        o.loss_function = loss; o.loss_scale = 2.0; o.max_num_iterations = 8
    xo, so, rows_o = oracle.solve_model_a(pr.n_cam, pr.n_pt, pr.cam_idx, pr.pt_idx, ob, pr.intr, pr.params, options=opt_o, n_threads=4)
    old = {k: os.environ.get(k) for k in env}
    os.environ.update({k: str(v) for k, v in env.items()})
    try:
        P = cuda.Problem(0)
        P.set_model_a(pr.n_cam, pr.n_pt, pr.cam_idx, pr.pt_idx, ob, pr.intr)
        P.set_parameters(pr.params)
        s, rows = P.solve(opt_g)
        x = P.get_parameters()
        # the generic (materialised Jacobian) pipeline on the same problem
        opt_g.force_generic_path = 1
        P.set_parameters(pr.params)
        s2, rows2 = P.solve(opt_g)
        P.close()
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    assert len(rows) == len(rows_o) == len(rows2)
    for a, b, c in zip(rows, rows_o, rows2):
        assert a["step_is_successful"] == b["step_is_successful"] == c["step_is_successful"]
        assert H.rel(a["cost"], b["cost"]) <= 1e-10 and H.rel(c["cost"], b["cost"]) <= 1e-10, (a, b, c)
    assert np.abs(x - xo).max() < 1e-7


@pytest.mark.gpu
@pytest.mark.parametrize("loss", [abi.LOSS_HUBER, abi.LOSS_CAUCHY])
def test_gpu_model_b_with_loss_matches_oracle(oracle, loss):
    pr = S.marker_rig_b(4, 12, 30, 7)
    rng = np.random.default_rng(4)
    obs8 = pr.obs8.copy()
    idx = rng.choice(obs8.shape[0], 40, replace=False)
    obs8[idx] += rng.normal(0.0, 25.0, (40, 8))
    pb = F.ModelBFile(pr.n_time, pr.n_cam, pr.n_marker, pr.counts, pr.time_idx, pr.cam_idx, pr.marker_idx, obs8, pr.params)
    opt_g, opt_o = cuda.default_options(), oracle.default_options()
    for o in (opt_g, opt_o):
        o.loss_function = loss; o.loss_scale = 3.0; o.max_num_iterations = 10
    xo, so, rows_o = oracle.solve_model_b(pb, pr.intr, pr.marker_side, 1, options=opt_o)
    P = cuda.Problem(0)
    P.set_model_b(pb.n_cam, pb.n_time, pb.n_marker, pb.time_idx, pb.cam_idx, pb.marker_idx, pb.obs8, pr.intr, pr.marker_side, 1)
    P.set_parameters(pb.params)
    s, rows = P.solve(opt_g)
    x = P.get_parameters()
    P.close()
    assert len(rows) == len(rows_o)
    for a, b in zip(rows, rows_o):
        assert a["step_is_successful"] == b["step_is_successful"]
        assert H.rel(a["cost"], b["cost"]) <= 1e-10, (a, b)
    assert np.abs(x - xo).max() < 1e-7
