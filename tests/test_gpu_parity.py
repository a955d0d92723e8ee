"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle and the reference's goldens.

Tolerances (BASELINE.json north_star): per-iteration cost within 1e-10 relative with the same trust-region
schedule and iteration count; final poses within 1e-7 rad / 1e-7 m."""
import os

import numpy as np
import pytest

from realsensecalibration_b200 import abi, cuda, formats as F, synthetic as S
from tests import helpers as H

pytestmark = pytest.mark.gpu

COST_RTOL = 1e-10
POSE_ATOL = 1e-7


@pytest.fixture(scope="module")
def gpu():
    if cuda.device_count() == 0:
        pytest.fail("no CUDA device: the gpu-marked tests must run on the B200 box")
    P = cuda.Problem(0)
    yield P
    P.close()


def _set_b(P, pb, intr, side, fix0, params=None):
    P.set_model_b(pb.n_cam, pb.n_time, pb.n_marker, pb.time_idx, pb.cam_idx, pb.marker_idx, pb.obs8, intr, side, fix0)
    P.set_parameters(pb.params if params is None else params)


def _check_rows(rows, rows_o):
    assert len(rows) == len(rows_o)
    g0 = rows_o[0]["gradient_max_norm"]   # near convergence the gradient is a difference of large terms: scale by row 0
    for a, b in zip(rows, rows_o):
        assert a["iteration"] == b["iteration"]
        assert a["step_is_successful"] == b["step_is_successful"] and a["step_is_valid"] == b["step_is_valid"]
        assert H.rel(a["cost"], b["cost"]) <= COST_RTOL, (a, b)
        # the radius update is a function of rho = cost_change / model_cost_change; near convergence cost_change is a
        # difference of two nearly equal sums of ~1e5 squares, so a 1e-14 relative difference between two summation orders
        # is amplified by cost / |cost_change| (and by <= 18 through radius / max(1/3, 1 - (2 rho - 1)^3))
        amp = abs(b["cost"]) / max(abs(b["cost_change"]), 1e-300) if b["cost_change"] != 0 else 0.0
        assert H.rel(a["trust_region_radius"], b["trust_region_radius"]) <= 1e-8 + 2e-13 * amp, (a, b)
        if b["gradient_max_norm"] > 0:
            assert abs(a["gradient_max_norm"] - b["gradient_max_norm"]) <= 1e-7 * b["gradient_max_norm"] + 1e-9 * g0
        if b["step_norm"] > 0:
            assert H.rel(a["step_norm"], b["step_norm"]) <= 1e-6


def _jac_close(j, jo):
    scale = np.maximum(np.abs(jo).max(axis=-1, keepdims=True), 1.0)
    return np.abs(j - jo).max() if j.size == 0 else (np.abs(j - jo) / scale).max()


# ---- K1: residual + Jacobian ------------------------------------------------------------------
def test_eval_model_b_hongo(gpu, oracle):
    pb, intr, side, fix0 = H.hongo()
    _set_b(gpu, pb, intr, side, fix0)
    cost, res, jac = gpu.eval()
    co, ro, jo = oracle.eval_model_b(pb, intr, side, fix0)
    assert H.rel(cost, co) < 1e-13 and H.rel(cost, H.HONGO_COSTS[0]) < 1e-12
    assert np.abs(res - ro).max() < 1e-9
    assert _jac_close(jac, jo) < 1e-11


def test_eval_model_b_test2_zero_angle_branch(gpu, oracle):
    # all marker rvecs are exactly 0 in this fixture: exercises the theta^2 <= eps branch and its derivative
    pb, intr, side, fix0 = H.test2()
    _set_b(gpu, pb, intr, side, fix0)
    cost, res, jac = gpu.eval()
    co, ro, jo = oracle.eval_model_b(pb, intr, side, fix0)
    assert H.rel(cost, co) < 1e-13
    assert np.abs(res - ro).max() < 1e-9
    assert _jac_close(jac, jo) < 1e-11


def test_eval_model_a_two_cam(gpu, oracle):
    pa, intr = H.two_cam()
    gpu.set_model_a(pa.n_cam, pa.n_pt, pa.cam_idx, pa.pt_idx, pa.obs_xy, intr)
    gpu.set_parameters(pa.params)
    cost, res, jac = gpu.eval()
    co, ro, jo = oracle.eval_model_a(pa.n_cam, pa.n_pt, pa.cam_idx, pa.pt_idx, pa.obs_xy, intr, pa.params)
    assert H.rel(cost, co) < 1e-13
    assert np.abs(res - ro).max() < 1e-10
    assert _jac_close(jac, jo) < 1e-11


@pytest.mark.parametrize("seed", [1, 2])
def test_eval_model_a_synthetic_unsorted(gpu, oracle, seed):
    pr = S.bal_like(40, 3000, 6, 16, seed, variable_degree=True)
    rng = np.random.default_rng(seed)
    sh = rng.permutation(pr.n_obs)  # caller order is arbitrary: the library sorts by point itself
    ci, pi, ob = pr.cam_idx[sh], pr.pt_idx[sh], pr.obs_xy[sh]
    gpu.set_model_a(pr.n_cam, pr.n_pt, ci, pi, ob, pr.intr)
    gpu.set_parameters(pr.params)
    cost, res, jac = gpu.eval()
    co, ro, jo = oracle.eval_model_a(pr.n_cam, pr.n_pt, ci, pi, ob, pr.intr, pr.params)
    assert H.rel(cost, co) < 1e-12
    assert np.abs(res - ro).max() < 1e-9
    assert _jac_close(jac, jo) < 1e-11


# ---- full LM solves against the reference's goldens -------------------------------------------
def test_solve_hongo_golden(gpu, oracle):
    pb, intr, side, fix0 = H.hongo()
    _set_b(gpu, pb, intr, side, fix0)
    s, rows = gpu.solve()
    x = gpu.get_parameters()
    xo, so, rows_o = oracle.solve_model_b(pb, intr, side, fix0)
    assert s.termination_type == abi.CONVERGENCE and s.termination_reason == abi.REASON_FUNCTION_TOLERANCE
    assert s.num_iterations == 7 and s.num_free_parameters == 114 and s.num_residuals == 544 and s.rcs_dim == 78
    _check_rows(rows, rows_o)
    for r, c in zip(rows, H.HONGO_COSTS):
        assert H.rel(r["cost"], c) < COST_RTOL
    H.check_hongo_golden(x, POSE_ATOL, POSE_ATOL)     # the reference's own committed Ceres output
    H.check_hongo_golden(x, 1e-10, 1e-10)             # and much tighter than the acceptance bar
    assert np.abs(x - xo).max() < 1e-9
    assert np.all(x[:6] == pb.params[:6])             # camera 0 is not a parameter block
    m0 = 6 * (pb.n_cam + pb.n_time)
    assert np.all(x[m0:m0 + 6] == pb.params[m0:m0 + 6])  # nor is marker 0 under the Main dispatch


def test_solve_test2_golden(gpu, oracle):
    pb, intr, side, fix0 = H.test2()
    _set_b(gpu, pb, intr, side, fix0)
    s, rows = gpu.solve()
    x = gpu.get_parameters()
    xo, so, rows_o = oracle.solve_model_b(pb, intr, side, fix0)
    assert s.termination_reason == abi.REASON_FUNCTION_TOLERANCE and s.num_iterations == 4
    _check_rows(rows, rows_o)
    H.check_test2_golden(x, 1e-10)
    assert np.abs(x - xo).max() < 1e-9


def test_solve_two_cam(gpu, oracle):
    pa, intr = H.two_cam()
    gpu.set_model_a(pa.n_cam, pa.n_pt, pa.cam_idx, pa.pt_idx, pa.obs_xy, intr)
    gpu.set_parameters(pa.params)
    s, rows = gpu.solve()
    x = gpu.get_parameters()
    xo, so, rows_o = oracle.solve_model_a(pa.n_cam, pa.n_pt, pa.cam_idx, pa.pt_idx, pa.obs_xy, intr, pa.params)
    assert s.termination_reason == so.termination_reason == abi.REASON_PARAMETER_TOLERANCE
    assert s.num_iterations == so.num_iterations == 3
    assert H.rel(rows[0]["cost"], rows_o[0]["cost"]) < COST_RTOL
    assert H.rel(rows[1]["cost"], rows_o[1]["cost"]) < 1e-7   # cost ~1e-4 of 32 residuals: near cancellation
    assert rows[2]["cost"] < 1e-10                            # exact fit: under-determined problem, cost -> 0
    assert np.abs(x - xo).max() < 1e-6
    # Test1's tail check (Test1_BundleAdjustment/main.cpp:89-126): the optimised points through the optimised camera
    err, rms = gpu.reprojection_error()
    assert 0.0 <= err < 1e-9
    pts = x[6 * pa.n_cam:].reshape(-1, 3)[pa.pt_idx]
    rt = x[:6 * pa.n_cam].reshape(-1, 6)
    img = np.asarray(pa.obs_xy, np.float64).reshape(-1, 2).astype(np.float32)
    e_g, r_g, rep_g = gpu.project_points_error(pts, pa.cam_idx, rt, np.asarray(intr, np.float64).reshape(-1, 4), img)
    e_o, r_o, rep_o = oracle.project_points_error(pts, pa.cam_idx, rt, np.asarray(intr, np.float64).reshape(-1, 4), img)
    assert np.abs(rep_g - rep_o).max() < 1e-9 and np.abs(rep_g - img).max() < 1e-3


@pytest.mark.parametrize("name", ["cfg1", "rigA", "balA", "rigB", "rigB_sparse"])
def test_solve_synthetic_vs_oracle(gpu, oracle, name):
    if name in ("rigB", "rigB_sparse"):
        pr = S.marker_rig_b(4, 12, 30, 7, visibility=1.0 if name == "rigB" else 0.3)
        pb = F.ModelBFile(pr.n_time, pr.n_cam, pr.n_marker, pr.counts, pr.time_idx, pr.cam_idx, pr.marker_idx, pr.obs8, pr.params)
        _set_b(gpu, pb, pr.intr, pr.marker_side, 1)
        xo, so, rows_o = oracle.solve_model_b(pb, pr.intr, pr.marker_side, 1)
    else:
        pr = {"cfg1": lambda: S.two_cam_like(50, 11), "rigA": lambda: S.marker_rig_a(8, 10, 20, 12),
              "balA": lambda: S.bal_like(60, 5000, 6, 16, 13, variable_degree=True)}[name]()
        gpu.set_model_a(pr.n_cam, pr.n_pt, pr.cam_idx, pr.pt_idx, pr.obs_xy, pr.intr)
        gpu.set_parameters(pr.params)
        xo, so, rows_o = oracle.solve_model_a(pr.n_cam, pr.n_pt, pr.cam_idx, pr.pt_idx, pr.obs_xy, pr.intr, pr.params)
    s, rows = gpu.solve()
    x = gpu.get_parameters()
    assert (s.termination_type, s.termination_reason) == (so.termination_type, so.termination_reason)
    assert s.num_free_parameters == so.num_free_parameters and s.rcs_dim == so.rcs_dim
    if name == "cfg1":   # zero-residual problem: the last rows sit at round-off level cost
        assert len(rows) == len(rows_o)
        assert H.rel(rows[0]["cost"], rows_o[0]["cost"]) <= COST_RTOL
    else:
        _check_rows(rows, rows_o)
        assert np.abs(x - xo).max() < POSE_ATOL


def test_model_a_generic_pipeline_matches_fused_and_oracle(gpu, oracle):
    # force_generic_path sends Model A through the materialised-Jacobian pipeline (K1 by k_fa_jac, the per-destination
    # incidence-pair lists built on demand): same rows as the fused two-pass path and as the oracle
    pr = S.bal_like(60, 5000, 6, 16, 13, variable_degree=True)
    gpu.set_model_a(pr.n_cam, pr.n_pt, pr.cam_idx, pr.pt_idx, pr.obs_xy, pr.intr)
    gpu.set_parameters(pr.params)
    xo, so, rows_o = oracle.solve_model_a(pr.n_cam, pr.n_pt, pr.cam_idx, pr.pt_idx, pr.obs_xy, pr.intr, pr.params)
    s_f, rows_f = gpu.solve()
    x_f = gpu.get_parameters()
    opt = cuda.default_options()
    opt.force_generic_path = 1
    gpu.set_parameters(pr.params)
    s_g, rows_g = gpu.solve(opt)
    x_g = gpu.get_parameters()
    _check_rows(rows_g, rows_o)
    _check_rows(rows_f, rows_o)
    assert np.abs(x_g - xo).max() < POSE_ATOL and np.abs(x_g - x_f).max() < POSE_ATOL
    assert (s_g.termination_type, s_g.termination_reason) == (so.termination_type, so.termination_reason)


def test_tiles_with_more_cameras_than_the_staged_tables(gpu, oracle):
    # window = all cameras: a 480-observation tile sees ~100 different cameras, more than the 48 tables the fused kernels
    # stage in shared memory, so most observations take the read-the-table-from-L2 branch of pass 1 / pass 2 / k_fa_jac
    pr = S.bal_like(120, 4000, 5, 120, 31)
    gpu.set_model_a(pr.n_cam, pr.n_pt, pr.cam_idx, pr.pt_idx, pr.obs_xy, pr.intr)
    gpu.set_parameters(pr.params)
    cost, res, jac = gpu.eval()
    co, ro, jo = oracle.eval_model_a(pr.n_cam, pr.n_pt, pr.cam_idx, pr.pt_idx, pr.obs_xy, pr.intr, pr.params)
    assert H.rel(cost, co) < 1e-12
    assert np.abs(res - ro).max() < 1e-8
    assert _jac_close(jac, jo) < 1e-10
    opt_g, opt_o = cuda.default_options(), oracle.default_options()
    for o in (opt_g, opt_o):
        o.max_num_iterations = 6
    xo, so, rows_o = oracle.solve_model_a(pr.n_cam, pr.n_pt, pr.cam_idx, pr.pt_idx, pr.obs_xy, pr.intr, pr.params, options=opt_o, n_threads=4)
    s, rows = gpu.solve(opt_g)
    x = gpu.get_parameters()
    _check_rows(rows, rows_o)
    assert np.abs(x - xo).max() < POSE_ATOL


@pytest.mark.parametrize("name", ["balA_n720", "rigB_n384", "balA_n1806_forced"])
def test_blocked_dense_cholesky_vs_oracle(gpu, oracle, name):
    # reduced camera systems too large for one CTA's shared memory (n > 160): blocked DMMA Cholesky over the whole GPU
    # (ba_dense_blocked.cuh); n not a multiple of the 32-wide panel / 64-wide tile in every case
    opt_g, opt_o = cuda.default_options(), oracle.default_options()
    for o in (opt_g, opt_o):
        o.rcs_solver = abi.RCS_DENSE_CHOLESKY
        o.max_num_iterations = 8
    if name == "rigB_n384":
        pr = S.marker_rig_b(6, 60, 40, 9, visibility=0.5)
        pb = F.ModelBFile(pr.n_time, pr.n_cam, pr.n_marker, pr.counts, pr.time_idx, pr.cam_idx, pr.marker_idx, pr.obs8, pr.params)
        _set_b(gpu, pb, pr.intr, pr.marker_side, 1)
        xo, so, rows_o = oracle.solve_model_b(pb, pr.intr, pr.marker_side, 1, options=opt_o)
    else:
        pr = S.bal_like(120, 8000, 5, 16, 21) if name == "balA_n720" else S.bal_like(301, 12000, 5, 20, 23)
        gpu.set_model_a(pr.n_cam, pr.n_pt, pr.cam_idx, pr.pt_idx, pr.obs_xy, pr.intr)
        gpu.set_parameters(pr.params)
        xo, so, rows_o = oracle.solve_model_a(pr.n_cam, pr.n_pt, pr.cam_idx, pr.pt_idx, pr.obs_xy, pr.intr, pr.params, options=opt_o,
                                              n_threads=4)
    s, rows = gpu.solve(opt_g)
    x = gpu.get_parameters()
    assert s.rcs_solver_used == abi.RCS_DENSE_CHOLESKY and s.rcs_dim == so.rcs_dim > 160
    _check_rows(rows, rows_o)
    assert np.abs(x - xo).max() < POSE_ATOL


@pytest.mark.parametrize("seed,perturb", [(24, (0.25, 0.07)), (27, (0.3, 0.08))])
def test_rejected_steps_follow_the_same_schedule(gpu, oracle, seed, perturb):
    # a poor start makes LM overshoot: rejected steps (2 resp. 3 of them), radius halving / quartering, then recovery.
    # The cases are picked so that the oracle's dense-normal and Schur paths agree to 1e-11 on every row, i.e. the
    # schedule is a property of the algorithm and not of round-off.
    pr = S.marker_rig_b(4, 8, 15, seed, perturb=perturb)
    pb = F.ModelBFile(pr.n_time, pr.n_cam, pr.n_marker, pr.counts, pr.time_idx, pr.cam_idx, pr.marker_idx, pr.obs8, pr.params)
    _set_b(gpu, pb, pr.intr, pr.marker_side, 1)
    s, rows = gpu.solve()
    x = gpu.get_parameters()
    xo, so, rows_o = oracle.solve_model_b(pb, pr.intr, pr.marker_side, 1)
    assert so.num_unsuccessful_steps >= 2, "test problem no longer produces rejected steps"
    assert s.num_unsuccessful_steps == so.num_unsuccessful_steps and s.num_iterations == so.num_iterations
    assert (s.termination_type, s.termination_reason) == (so.termination_type, so.termination_reason)
    for a, b in zip(rows, rows_o):
        assert a["step_is_successful"] == b["step_is_successful"]
        # the radius update is a function of rho = cost_change / model_cost_change; near convergence cost_change is a
        # difference of two nearly equal sums of ~1e5 squares, so a 1e-14 relative difference between two summation orders
        # is amplified by cost / |cost_change| (and by <= 18 through radius / max(1/3, 1 - (2 rho - 1)^3))
        amp = abs(b["cost"]) / max(abs(b["cost_change"]), 1e-300) if b["cost_change"] != 0 else 0.0
        assert H.rel(a["trust_region_radius"], b["trust_region_radius"]) <= 1e-8 + 2e-13 * amp, (a, b)
        assert H.rel(a["cost"], b["cost"]) <= 1e-9
    assert np.abs(x - xo).max() < POSE_ATOL


# ---- K3b: block-sparse PCG on the reduced camera system ------------------------------------------
@pytest.mark.parametrize("name", ["balA", "chain"])
def test_pcg_matches_oracle_pcg(gpu, oracle, name):
    # same truncated-CG rule on both sides (Ceres' ITERATIVE_SCHUR + SCHUR_JACOBI): same LM schedule, same
    # number of CG iterations per linear solve (the inexact steps are part of the algorithm's definition)
    pr = S.bal_like(60, 5000, 6, 16, 13, variable_degree=True) if name == "balA" else S.bal_like(300, 20000, 5, 20, 17)
    opt_g, opt_o = cuda.default_options(), oracle.default_options()
    for o in (opt_g, opt_o):
        o.rcs_solver = abi.RCS_PCG
        o.max_num_iterations = 6
    gpu.set_model_a(pr.n_cam, pr.n_pt, pr.cam_idx, pr.pt_idx, pr.obs_xy, pr.intr)
    gpu.set_parameters(pr.params)
    s, rows = gpu.solve(opt_g)
    x = gpu.get_parameters()
    xo, so, rows_o = oracle.solve_model_a(pr.n_cam, pr.n_pt, pr.cam_idx, pr.pt_idx, pr.obs_xy, pr.intr, pr.params, options=opt_o,
                                          linear_solver=oracle.SCHUR_PCG, n_threads=4)
    assert s.rcs_solver_used == abi.RCS_PCG
    assert len(rows) == len(rows_o)
    # The Q-based stopping test (zeta < eta) can flip by one CG iteration on round-off: from the first row where the
    # counts differ, the two runs take (slightly) different inexact steps and only the looser bound applies.
    same = True
    for a, b in zip(rows, rows_o):
        assert a["step_is_successful"] == b["step_is_successful"]
        ia, ib = a["linear_solver_iterations"], b["linear_solver_iterations"]
        assert abs(ia - ib) <= max(2, 0.1 * ib), (a, b)
        same = same and ia == ib
        assert H.rel(a["cost"], b["cost"]) <= (1e-5 if same else 1e-3), (a, b)   # truncated CG amplifies summation-order round-off
        # the radius follows rho = cost_change / model_cost_change through the cubic rule: late rows, where the cost
        # change is 1e-4 of the cost, turn a 1e-6 cost difference into a 1e-3 radius difference (the dense solver on
        # the same problem agrees with the oracle to 1e-13, see test_solve_synthetic_vs_oracle)
        assert H.rel(a["trust_region_radius"], b["trust_region_radius"]) <= (5e-3 if same else 0.5)
    assert np.abs(x - xo).max() < (1e-4 if same else 1e-2)
    assert rows[1]["linear_solver_iterations"] == rows_o[1]["linear_solver_iterations"] > 0


@pytest.mark.parametrize("name,k", [("balA", 10), ("balA", 40), ("chain", 15)])
def test_pcg_with_a_fixed_iteration_count_is_the_same_method(gpu, oracle, name, k):
    # pcg_min_iterations = pcg_max_iterations = k on both sides takes the stopping rule out of the comparison: the k-step
    # CG iterate is a fixed polynomial of S and the right-hand side, so the LM rows must agree to round-off level
    pr = S.bal_like(60, 5000, 6, 16, 13, variable_degree=True) if name == "balA" else S.bal_like(300, 20000, 5, 20, 17)
    opt_g, opt_o = cuda.default_options(), oracle.default_options()
    for o in (opt_g, opt_o):
        o.rcs_solver = abi.RCS_PCG
        o.max_num_iterations = 6
        o.pcg_min_iterations = k
        o.pcg_max_iterations = k
    xo, so, rows_o = oracle.solve_model_a(pr.n_cam, pr.n_pt, pr.cam_idx, pr.pt_idx, pr.obs_xy, pr.intr, pr.params, options=opt_o,
                                          linear_solver=oracle.SCHUR_PCG, n_threads=4)
    for env in (dict(BA_SA=2), dict(BA_SA=0)):   # pass 1 on strips and on single tiles
        with _Env(**env):
            s, rows, x = _solve_a(gpu, pr, opt_g)
        assert s.rcs_solver_used == abi.RCS_PCG and len(rows) == len(rows_o)
        for a, b in zip(rows, rows_o):
            assert a["linear_solver_iterations"] == b["linear_solver_iterations"] == (k if a["iteration"] > 0 else 0)
            assert a["step_is_successful"] == b["step_is_successful"]
            assert H.rel(a["cost"], b["cost"]) <= 1e-9, (env, a, b)
        assert np.abs(x - xo).max() < POSE_ATOL


def test_pcg_converges_to_the_dense_solution(gpu):
    # with a tight CG tolerance the PCG step is the exact step: the LM trace must then equal the dense-Cholesky trace
    pr = S.bal_like(40, 3000, 6, 12, 3)
    traces = []
    for solver in (abi.RCS_DENSE_CHOLESKY, abi.RCS_PCG):
        opt = cuda.default_options()
        opt.rcs_solver = solver
        opt.pcg_eta = 1e-14
        opt.pcg_max_iterations = 2000
        gpu.set_model_a(pr.n_cam, pr.n_pt, pr.cam_idx, pr.pt_idx, pr.obs_xy, pr.intr)
        gpu.set_parameters(pr.params)
        s, rows = gpu.solve(opt)
        traces.append((s, rows, gpu.get_parameters()))
    (s0, r0, x0), (s1, r1, x1) = traces
    assert s0.num_iterations == s1.num_iterations and s0.termination_reason == s1.termination_reason
    # Model A leaves the 7 gauge freedoms to the LM damping alone: along them the system is nearly singular and CG
    # stalls at ~1e-7, which is what bounds the agreement of the two traces
    for a, b in zip(r0, r1):
        assert H.rel(a["cost"], b["cost"]) <= 1e-5
    assert H.rel(s0.final_cost, s1.final_cost) <= 1e-6


def test_max_iterations_and_determinism(gpu):
    pr = S.bal_like(30, 2000, 5, 12, 5)
    opt = cuda.default_options()
    opt.max_num_iterations = 2
    xs = []
    for _ in range(2):
        gpu.set_model_a(pr.n_cam, pr.n_pt, pr.cam_idx, pr.pt_idx, pr.obs_xy, pr.intr)
        gpu.set_parameters(pr.params)
        s, rows = gpu.solve(opt)
        assert s.termination_type == abi.NO_CONVERGENCE and s.termination_reason == abi.REASON_MAX_ITERATIONS
        assert s.num_iterations == 3
        xs.append(gpu.get_parameters())
    assert np.array_equal(xs[0], xs[1])   # no floating-point atomics: bitwise reproducible


def test_solve_again_on_the_same_problem_and_without_jacobi_scaling(gpu, oracle):
    # bench.py restarts solves on one problem (restore + solve_begin): the second solve must repeat the first bit for
    # bit; then the same problem without Jacobi scaling (the scaling of the earlier solves must not leak into it), and
    # in steps through solve_begin / solve_iterate with an iteration limit that ends on the cost + gradient only pass
    pr = S.bal_like(40, 3000, 5, 12, 29)
    gpu.set_model_a(pr.n_cam, pr.n_pt, pr.cam_idx, pr.pt_idx, pr.obs_xy, pr.intr)
    gpu.set_parameters(pr.params)
    gpu.save_parameters()
    s1, rows1 = gpu.solve()
    x1 = gpu.get_parameters()
    gpu.restore_parameters()
    s2, rows2 = gpu.solve()
    assert np.array_equal(x1, gpu.get_parameters()) and [r["cost"] for r in rows1] == [r["cost"] for r in rows2]
    opt_g, opt_o = cuda.default_options(), oracle.default_options()
    for o in (opt_g, opt_o):
        o.jacobi_scaling = 0
        o.max_num_iterations = 4
    xo, so, rows_o = oracle.solve_model_a(pr.n_cam, pr.n_pt, pr.cam_idx, pr.pt_idx, pr.obs_xy, pr.intr, pr.params, options=opt_o, n_threads=4)
    gpu.restore_parameters()
    s3, rows3 = gpu.solve(opt_g)
    _check_rows(rows3, rows_o)
    assert np.abs(gpu.get_parameters() - xo).max() < POSE_ATOL
    # the same limit reached in two calls of solve_iterate, with scaling
    opt_g.jacobi_scaling = 1; opt_o.jacobi_scaling = 1
    xo, so, rows_o = oracle.solve_model_a(pr.n_cam, pr.n_pt, pr.cam_idx, pr.pt_idx, pr.obs_xy, pr.intr, pr.params, options=opt_o, n_threads=4)
    gpu.restore_parameters()
    gpu.solve_begin(opt_g)
    gpu.solve_iterate(2)
    gpu.solve_iterate(10)
    s4, rows4 = gpu.solve_end()
    _check_rows(rows4, rows_o)
    assert np.abs(gpu.get_parameters() - xo).max() < POSE_ATOL


# ---- K5 and the post-BA outputs ------------------------------------------------------------------
def test_outputs_and_reprojection_hongo(gpu, oracle, golden_dir):
    pb, intr, side, fix0 = H.hongo()
    _set_b(gpu, pb, intr, side, fix0)
    err0, rms0 = gpu.reprojection_error()
    assert H.rel(err0, H.HONGO_COSTS[0]) < 1e-12
    gpu.solve()
    x = gpu.get_parameters()
    err, rms = gpu.reprojection_error()
    assert H.rel(err, 143.62938885186) < 1e-10 and abs(rms - 0.72667) < 1e-4
    rot, inv, corners = gpu.model_b_outputs()
    rot_o, inv_o, corners_o = oracle.model_b_outputs(pb, x, side)
    assert np.abs(rot - rot_o).max() < 1e-14 and np.abs(inv - inv_o).max() < 1e-14
    assert np.abs(corners - corners_o).max() < 1e-14
    gold = F.load_opencv_xml(os.path.join(golden_dir, "Correspondence", "hongo", "Camera_Transform.xml"))
    for c in range(4):
        assert np.abs(rot[c] - gold["R%d" % c]).max() < 1e-10
        ext = F.load_extrinsics(os.path.join(golden_dir, "Calibration", "Extrinsics", "mat%d.txt" % c))
        assert np.abs(inv[c] - ext).max() < 6e-7
    _, pts = F.load_point3d(os.path.join(golden_dir, "Correspondence", "hongo", "point3d.txt"))
    assert np.abs(pts - corners).max() < 6e-7
    # ReprojectionCheck::Reproject on the 6-digit points of point3d.txt with float image points
    cam_of_point = np.repeat(pb.cam_idx, 4)
    rt = x[:24].reshape(4, 6)
    img = pb.obs8.reshape(-1, 2).astype(np.float32)
    e_g, r_g, rep_g = gpu.project_points_error(pts, cam_of_point, rt, intr, img)
    e_o, r_o, rep_o = oracle.project_points_error(pts, cam_of_point, rt, intr, img)
    assert H.rel(e_g, e_o) < 1e-12 and H.rel(r_g, r_o) < 1e-12 and np.abs(rep_g - rep_o).max() < 1e-9
    assert abs(e_g - 143.63) < 0.5   # 6-digit rounding of the points moves the error slightly


def test_empty_and_ragged_inputs(gpu, oracle):
    # a point / camera that no observation references is not part of the problem; zero observations is legal
    pr = S.bal_like(12, 300, 4, 8, 9)
    keep = (pr.pt_idx % 7 != 3) & (pr.cam_idx != 5)
    ci, pi, ob = pr.cam_idx[keep], pr.pt_idx[keep], pr.obs_xy[keep]
    gpu.set_model_a(pr.n_cam, pr.n_pt, ci, pi, ob, pr.intr)
    gpu.set_parameters(pr.params)
    s, rows = gpu.solve()
    x = gpu.get_parameters()
    xo, so, rows_o = oracle.solve_model_a(pr.n_cam, pr.n_pt, ci, pi, ob, pr.intr, pr.params)
    _check_rows(rows, rows_o)
    assert s.num_free_parameters == so.num_free_parameters
    assert np.abs(x - xo).max() < POSE_ATOL
    unused_pt = np.where(np.bincount(pi, minlength=pr.n_pt) == 0)[0]
    assert unused_pt.size > 0
    X0 = pr.params[6 * pr.n_cam:].reshape(-1, 3); X1 = x[6 * pr.n_cam:].reshape(-1, 3)
    assert np.array_equal(X0[unused_pt], X1[unused_pt]) and np.array_equal(x[30:36], pr.params[30:36])
    gpu.set_model_a(2, 3, np.zeros(0, np.int32), np.zeros(0, np.int32), np.zeros((0, 2)), pr.intr[:2])
    gpu.set_parameters(np.zeros(21))
    cost, res, jac = gpu.eval()
    assert cost == 0.0 and res.shape == (0, 2)


def test_error_paths(gpu):
    with pytest.raises(cuda.BAError):
        gpu.set_model_a(2, 3, np.array([2], np.int32), np.array([0], np.int32), np.zeros((1, 2)), np.ones(8))
    P2 = cuda.Problem(0)
    with pytest.raises(cuda.BAError):
        P2.solve()
    P2.close()


# ---- strip pass 1 (ba_strip_a.cuh): register-resident Schur accumulators across a strip of tiles -------------------
def _solve_a(gpu, pr, opt=None):
    gpu.set_model_a(pr.n_cam, pr.n_pt, pr.cam_idx, pr.pt_idx, pr.obs_xy, pr.intr)
    gpu.set_parameters(pr.params)
    s, rows = gpu.solve(opt)
    return s, rows, gpu.get_parameters()


class _Env:
    def __init__(self, **kv):
        self.kv = {k: str(v) for k, v in kv.items()}

    def __enter__(self):
        self.old = {k: os.environ.get(k) for k in self.kv}
        os.environ.update(self.kv)

    def __exit__(self, *a):
        for k, v in self.old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


@pytest.mark.parametrize("name", ["balA", "rigA", "chain"])
def test_strip_path_matches_tile_path_and_oracle(gpu, oracle, name):
    # the same problem through pass 1 on strips (default) and on single tiles (BA_SA=0, the first-generation kernel):
    # same rows as the oracle, and as each other
    pr = {"balA": lambda: S.bal_like(60, 5000, 6, 16, 13, variable_degree=True), "rigA": lambda: S.marker_rig_a(8, 10, 20, 12),
          "chain": lambda: S.bal_like(300, 20000, 5, 20, 17)}[name]()
    opt_g, opt_o = cuda.default_options(), oracle.default_options()
    for o in (opt_g, opt_o):
        o.max_num_iterations = 6
        o.rcs_solver = abi.RCS_DENSE_CHOLESKY
    xo, so, rows_o = oracle.solve_model_a(pr.n_cam, pr.n_pt, pr.cam_idx, pr.pt_idx, pr.obs_xy, pr.intr, pr.params, options=opt_o, n_threads=4)
    with _Env(BA_SA=2, BA_RIG=0):   # 2 = strips whenever the problem fits (the default also asks for >= 4 pair products per observation)
        s1, rows1, x1 = _solve_a(gpu, pr, opt_g)
    assert s1.path_used == abi.PATH_FUSED_STRIPS
    with _Env(BA_SA=0, BA_RIG=0):   # BA_RIG=0: the rig is small enough for the one-CTA path (tested further down)
        s0, rows0, x0 = _solve_a(gpu, pr, opt_g)
    assert s0.path_used == abi.PATH_FUSED_TILES
    _check_rows(rows1, rows_o)
    _check_rows(rows0, rows_o)
    assert np.abs(x1 - xo).max() < POSE_ATOL and np.abs(x1 - x0).max() < POSE_ATOL


@pytest.mark.parametrize("geom", [dict(BA_SA_TOBS=128, BA_SA_L=3), dict(BA_SA_TOBS=256, BA_SA_L=1), dict(BA_SA_TOBS=1024, BA_SA_L=16)])
def test_strip_geometries(gpu, oracle, geom):
    # many short tiles per strip (the bulk-copy prefetch runs across tile boundaries), one tile per strip, one strip
    pr = S.bal_like(60, 5000, 6, 16, 13, variable_degree=True)
    opt_g, opt_o = cuda.default_options(), oracle.default_options()
    for o in (opt_g, opt_o):
        o.max_num_iterations = 5
    xo, so, rows_o = oracle.solve_model_a(pr.n_cam, pr.n_pt, pr.cam_idx, pr.pt_idx, pr.obs_xy, pr.intr, pr.params, options=opt_o, n_threads=4)
    with _Env(BA_SA=2, **geom):
        s, rows, x = _solve_a(gpu, pr, opt_g)
    assert s.path_used == abi.PATH_FUSED_STRIPS
    _check_rows(rows, rows_o)
    assert np.abs(x - xo).max() < POSE_ATOL


def test_strip_flush_warp_and_split_owners(gpu, oracle):
    # window of 40 cameras: a strip couples up to ~55 x 54 / 2 camera pairs, more than the 512 owner threads, so the
    # lightest pairs of every strip go through the flush warp (one partial block per tile); the rig has 28 pairs and
    # 8 cameras for 512 threads, so every pair and camera is split over many owner threads
    pr = S.bal_like(120, 6000, 8, 40, 41)
    opt_g, opt_o = cuda.default_options(), oracle.default_options()
    for o in (opt_g, opt_o):
        o.max_num_iterations = 5
    xo, so, rows_o = oracle.solve_model_a(pr.n_cam, pr.n_pt, pr.cam_idx, pr.pt_idx, pr.obs_xy, pr.intr, pr.params, options=opt_o, n_threads=4)
    with _Env(BA_SA=2):
        s, rows, x = _solve_a(gpu, pr, opt_g)
    assert s.path_used == abi.PATH_FUSED_STRIPS
    _check_rows(rows, rows_o)
    assert np.abs(x - xo).max() < POSE_ATOL
    pr = S.marker_rig_a(8, 40, 60, 19)
    xo, so, rows_o = oracle.solve_model_a(pr.n_cam, pr.n_pt, pr.cam_idx, pr.pt_idx, pr.obs_xy, pr.intr, pr.params, options=opt_o, n_threads=4)
    s, rows, x = _solve_a(gpu, pr, opt_g)   # 4.5 pair products per observation: strips by default
    assert s.path_used == abi.PATH_FUSED_STRIPS
    _check_rows(rows, rows_o)
    assert np.abs(x - xo).max() < POSE_ATOL


def test_point_seen_twice_by_one_camera_takes_the_tile_path(gpu, oracle):
    # a duplicated observation (same point, same camera): the strip plan has no owner for a (camera, camera) pair of
    # two different observations, so the problem runs on single tiles and still matches the oracle
    pr = S.bal_like(30, 2000, 5, 12, 5)
    ci = np.concatenate([pr.cam_idx, pr.cam_idx[:1]]); pi = np.concatenate([pr.pt_idx, pr.pt_idx[:1]])
    ob = np.concatenate([pr.obs_xy, pr.obs_xy[:1] + 0.25])
    gpu.set_model_a(pr.n_cam, pr.n_pt, ci, pi, ob, pr.intr)
    gpu.set_parameters(pr.params)
    s, rows = gpu.solve()
    x = gpu.get_parameters()
    assert s.path_used == abi.PATH_FUSED_TILES
    xo, so, rows_o = oracle.solve_model_a(pr.n_cam, pr.n_pt, ci, pi, ob, pr.intr, pr.params)
    _check_rows(rows, rows_o)
    assert np.abs(x - xo).max() < POSE_ATOL


def test_point_with_more_than_64_observations_takes_the_generic_pipeline(gpu, oracle, capfd):
    # the fused passes keep the observations of a point in one tile and cap them at 64 (FA_KMAX); a denser point sends the
    # whole problem through the generic materialised-Jacobian pipeline: said so in summary.path_used (and once on stderr),
    # and still the oracle's rows
    pr = S.bal_like(70, 300, 66, 70, 33)
    opt_g, opt_o = cuda.default_options(), oracle.default_options()
    for o in (opt_g, opt_o):
        o.max_num_iterations = 5
    xo, so, rows_o = oracle.solve_model_a(pr.n_cam, pr.n_pt, pr.cam_idx, pr.pt_idx, pr.obs_xy, pr.intr, pr.params, options=opt_o, n_threads=4)
    s, rows, x = _solve_a(gpu, pr, opt_g)
    assert s.path_used == abi.PATH_GENERIC
    _check_rows(rows, rows_o)
    assert np.abs(x - xo).max() < POSE_ATOL


# ---- the numeric part of the correspondence stage --------------------------------------------------------
def test_correspondence_stage_pose_algebra_vs_opencv(gpu):
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(8)
    n = 500
    a = np.concatenate([rng.normal(0, 1.2, (n, 3)), rng.normal(0, 0.3, (n, 3))], 1)
    b = np.concatenate([rng.normal(0, 1.2, (n, 3)), rng.normal(0, 0.3, (n, 3))], 1)
    a[0, :3] = 0.0; b[1, :3] = 0.0                      # identity rotations
    a[2, :3] = b[2, :3]                                   # a o b^-1 = identity rotation
    a[3, :3] = [np.pi - 1e-7, 0.0, 0.0]; b[3, :3] = 0.0   # rotation by (almost) pi: the s < 1e-5 branch of cv::Rodrigues
    side = 0.0148
    # Correspondencer::GetCornersInCameraWorld (correspondencer.cpp:5-39)
    got = gpu.marker_corners(a, side)
    h = side / 2
    for i in range(n):
        R = cv2.Rodrigues(a[i, :3])[0]
        E, Fv, t = R[:, 0] * h, R[:, 1] * h, a[i, 3:]
        ref = np.stack([t - E + Fv, t + E + Fv, t + E - Fv, t - E - Fv])
        assert np.abs(got[i] - ref).max() < 1e-14
    # pose composition (correspondencer.cpp:118-121 and :141-146), rotation vectors through cv::Rodrigues both ways
    for inv in (False, True):
        out = gpu.compose_poses(a, b, inv)
        for i in range(n):
            Ra, Rb = cv2.Rodrigues(a[i, :3])[0], cv2.Rodrigues(b[i, :3])[0]
            R = Ra @ (Rb.T if inv else Rb)
            t = R @ (-b[i, 3:]) + a[i, 3:] if inv else Ra @ b[i, 3:] + a[i, 3:]
            assert np.abs(out[i, 3:] - t).max() < 1e-13
            if i != 3:
                assert np.abs(cv2.Rodrigues(out[i, :3])[0] - R).max() < 1e-12, (i, inv)
                assert np.abs(out[i, :3] - cv2.Rodrigues(R)[0].ravel()).max() < 1e-9
            else:   # within 1e-5 of pi cv::Rodrigues recovers the axis from the diagonal alone: both sides are good to ~1e-7 there
                assert np.abs(cv2.Rodrigues(out[i, :3])[0] - R).max() < 1e-6
                assert np.abs(np.abs(out[i, :3]) - np.abs(cv2.Rodrigues(R)[0].ravel())).max() < 1e-6


def test_before_ba_reprojection_error_from_the_initial_guess(gpu):
    # correspondencer.cpp:284-339 on the committed hongo file: marker-from-base-camera poses composed from the frame and marker
    # blocks, corners by GetCornersInCameraWorld, projected through the initial camera poses: the "Reprojection Error (Before BA)"
    # is the cost of row 0 of the solve (same corners, same projection, same ((dx)^2 + (dy)^2) / 2)
    pb, intr, side, fix0 = H.hongo()
    C, T, M = pb.n_cam, pb.n_time, pb.n_marker
    cams = pb.params[:6 * C].reshape(C, 6)
    frames = pb.params[6 * C:6 * (C + T)].reshape(T, 6)
    markers = pb.params[6 * (C + T):].reshape(M, 6)
    pose = gpu.compose_poses(frames[pb.time_idx], markers[pb.marker_idx])          # base o marker-from-base
    corners = gpu.marker_corners(pose, side).reshape(-1, 3)
    cam_of_point = np.repeat(pb.cam_idx, 4)
    img = pb.obs8.reshape(-1, 2).astype(np.float32)
    err, rms, _ = gpu.project_points_error(corners, cam_of_point, cams, intr, img)
    assert H.rel(err, H.HONGO_COSTS[0]) < 1e-9


# ---- the rig-size path: the whole trust-region loop in one launch of one CTA (csrc/ba_rig.cuh) ------------------
def _rig_cases():
    pb, intr, side, fix0 = H.hongo()
    yield "hongo", "B", (pb, intr, side, fix0)
    pb, intr, side, fix0 = H.test2()
    yield "test2", "B", (pb, intr, side, fix0)
    pr = S.marker_rig_b(4, 8, 15, 24, perturb=(0.25, 0.07))   # rejected steps: radius halving / quartering, then recovery
    pb = F.ModelBFile(pr.n_time, pr.n_cam, pr.n_marker, pr.counts, pr.time_idx, pr.cam_idx, pr.marker_idx, pr.obs8, pr.params)
    yield "rejected", "B", (pb, pr.intr, pr.marker_side, 1)
    pa, intr = H.two_cam()
    yield "two_cam", "A", (pa, intr)


@pytest.mark.parametrize("case", list(_rig_cases()), ids=lambda c: c[0])
def test_rig_path_is_the_same_method_as_the_multi_kernel_pipeline(gpu, oracle, case):
    name, model, args = case

    def load():
        if model == "B":
            _set_b(gpu, *args)
        else:
            pa, intr = args
            gpu.set_model_a(pa.n_cam, pa.n_pt, pa.cam_idx, pa.pt_idx, pa.obs_xy, intr)
            gpu.set_parameters(pa.params)

    load()
    s1, rows1 = gpu.solve()
    x1 = gpu.get_parameters()
    assert s1.path_used == abi.PATH_RIG
    with _Env(BA_RIG=0):
        load()
        s0, rows0 = gpu.solve()
        x0 = gpu.get_parameters()
    assert s0.path_used != abi.PATH_RIG
    assert (s1.termination_type, s1.termination_reason, s1.num_iterations) == (s0.termination_type, s0.termination_reason, s0.num_iterations)
    assert (s1.num_successful_steps, s1.num_unsuccessful_steps) == (s0.num_successful_steps, s0.num_unsuccessful_steps)
    assert (s1.num_jacobian_evaluations, s1.num_linear_solves) == (s0.num_jacobian_evaluations, s0.num_linear_solves)
    assert H.rel(s1.final_cost, s0.final_cost) <= (COST_RTOL if name != "two_cam" else 1.0) or abs(s1.final_cost) < 1e-10
    for a, b in zip(rows1, rows0):
        assert (a["iteration"], a["step_is_valid"], a["step_is_successful"]) == (b["iteration"], b["step_is_valid"], b["step_is_successful"])
        if name != "two_cam" or b["cost"] > 1e-3:   # two_cam ends in an exact fit: costs ~1e-12 are round-off
            assert H.rel(a["cost"], b["cost"]) <= 1e-9, (a, b)
    assert np.abs(x1 - x0).max() < (POSE_ATOL if name != "two_cam" else 1e-6)
    # the same solve one iteration per call (ba_cuda_solve_begin / _iterate / _end): same rows, bit for bit
    load()
    gpu.solve_begin()
    n_calls = 0
    while not gpu.solve_iterate(1):
        n_calls += 1
        assert n_calls < 100
    s2, rows2 = gpu.solve_end()
    assert s2.path_used == abi.PATH_RIG and s2.num_iterations == s1.num_iterations
    for a, b in zip(rows2, rows1):
        assert a["cost"] == b["cost"] and a["trust_region_radius"] == b["trust_region_radius"]
    assert np.array_equal(gpu.get_parameters(), x1)


def test_rig_path_with_a_robust_loss_and_without_jacobi_scaling(gpu, oracle):
    pb, intr, side, fix0 = H.hongo()
    for kw in (dict(loss_function=abi.LOSS_HUBER, loss_scale=1.5), dict(jacobi_scaling=0)):
        opt = cuda.default_options()
        for k, v in kw.items():
            setattr(opt, k, v)
        _set_b(gpu, pb, intr, side, fix0)
        s1, rows1 = gpu.solve(opt)
        x1 = gpu.get_parameters()
        assert s1.path_used == abi.PATH_RIG
        with _Env(BA_RIG=0):
            _set_b(gpu, pb, intr, side, fix0)
            s0, rows0 = gpu.solve(opt)
            x0 = gpu.get_parameters()
        assert s0.path_used == abi.PATH_GENERIC and s1.num_iterations == s0.num_iterations
        for a, b in zip(rows1, rows0):
            assert H.rel(a["cost"], b["cost"]) <= 1e-9 and a["step_is_successful"] == b["step_is_successful"]
        assert np.abs(x1 - x0).max() < POSE_ATOL


@pytest.mark.parametrize("rig", [1, 0])
def test_structure_built_on_the_host_is_the_device_built_one(gpu, oracle, rig):
    # rig-size Model B problems get their lists (sorted observations, incidences, destination blocks, pair lists, chunk tables)
    # from the host in one upload (ba_structure.cuh, build_structure_host); the orders are the device build's, so every
    # sum runs in the same order: the two solves agree bit for bit, through the one-CTA kernel and through the multi-kernel pipeline
    cases = [H.hongo(), H.test2()]
    pr = S.marker_rig_b(4, 8, 15, 24, perturb=(0.25, 0.07))
    cases.append((F.ModelBFile(pr.n_time, pr.n_cam, pr.n_marker, pr.counts, pr.time_idx, pr.cam_idx, pr.marker_idx, pr.obs8, pr.params),
                  pr.intr, pr.marker_side, 1))
    pr = S.marker_rig_b(3, 5, 40, 7, visibility=0.6)   # ragged: frames that miss cameras / markers
    cases.append((F.ModelBFile(pr.n_time, pr.n_cam, pr.n_marker, pr.counts, pr.time_idx, pr.cam_idx, pr.marker_idx, pr.obs8, pr.params),
                  pr.intr, pr.marker_side, 1))
    for pb, intr, side, fix0 in cases:
        out = []
        for host_max in (4096, 0):
            with _Env(BA_HOST_BUILD_MAX=host_max, BA_RIG=rig):
                _set_b(gpu, pb, intr, side, fix0)
                s, rows = gpu.solve()
                out.append((s, rows, gpu.get_parameters(), gpu.eval()))
        (s1, rows1, x1, ev1), (s0, rows0, x0, ev0) = out
        assert s1.num_iterations == s0.num_iterations and s1.path_used == s0.path_used
        assert [r["cost"] for r in rows1] == [r["cost"] for r in rows0]
        assert [r["trust_region_radius"] for r in rows1] == [r["trust_region_radius"] for r in rows0]
        assert np.array_equal(x1, x0)
        assert ev1[0] == ev0[0] and np.array_equal(ev1[1], ev0[1]) and np.array_equal(ev1[2], ev0[2])


def test_model_a_structure_built_on_the_host(gpu, oracle):
    # Test1 / Test2 sizes of Model A: lists from the host, no tile / strip structures; the one-CTA kernel gives bit for bit
    # the rows it gives on the device-built structure, and with BA_RIG=0 the generic pipeline takes over (same oracle rows)
    pa, intr = H.two_cam()
    cases = [(pa.n_cam, pa.n_pt, pa.cam_idx, pa.pt_idx, pa.obs_xy, intr, pa.params)]
    pr = S.bal_like(8, 600, 4, 6, 3)
    cases.append((pr.n_cam, pr.n_pt, pr.cam_idx, pr.pt_idx, pr.obs_xy, pr.intr, pr.params))
    pr = S.marker_rig_a(4, 3, 20, 5)
    cases.append((pr.n_cam, pr.n_pt, pr.cam_idx, pr.pt_idx, pr.obs_xy, pr.intr, pr.params))
    for k, (n_cam, n_pt, ci, pi, ob, K, x0) in enumerate(cases):
        out = []
        for host_max in (4096, 0):
            with _Env(BA_HOST_BUILD_MAX=host_max):
                gpu.set_model_a(n_cam, n_pt, ci, pi, ob, K)
                gpu.set_parameters(x0)
                s, rows = gpu.solve()
                out.append((s, rows, gpu.get_parameters(), gpu.eval()))
        (s1, rows1, x1, ev1), (s0, rows0, x0_, ev0) = out
        assert s1.path_used == s0.path_used == abi.PATH_RIG
        assert [r["cost"] for r in rows1] == [r["cost"] for r in rows0]
        assert np.array_equal(x1, x0_)
        assert ev1[0] == ev0[0] or H.rel(ev1[0], ev0[0]) < 1e-12   # ba_cuda_eval: k_jac_a on the host-built one, k_fa_jac on the other
        if k == 0:
            continue   # the two-camera fixture ends in an exact fit (costs ~1e-12): covered by test_rig_path_*
        xo, so, rows_o = oracle.solve_model_a(n_cam, n_pt, ci, pi, ob, K, x0)
        _check_rows(rows1, rows_o)
        with _Env(BA_RIG=0):   # host-built and not on the rig kernel: the generic pipeline
            gpu.set_model_a(n_cam, n_pt, ci, pi, ob, K)
            gpu.set_parameters(x0)
            s2, rows2 = gpu.solve()
        assert s2.path_used != abi.PATH_RIG
        _check_rows(rows2, rows_o)
