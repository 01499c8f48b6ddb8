"""CPU checks of the box-constrained TVLQR oracle (oracle/box_tvlqr.py): the ADMM restatement of the
reference's bounded QP (irs_lqr/tv_lqr.py:69-137 with :113-118,:132-134) against an independent dense
QP solve, its independence from the penalty parameter, and the MPC loop of irs_lqr.py:169-184."""
import numpy as np
import pytest

from oracle import box_tvlqr as bq
from oracle import cpu_restatement as cr
from oracle import example_configs as ec


def _problem(T, u_const=(0.5, 0.6)):
    cfg = ec.bicycle(T=T)
    orc = cr.BicycleOracle(cfg["h"])
    u0 = np.tile(np.array(u_const), (T, 1))
    x_trj = cr.rollout(orc, cfg["x0"], u0)
    At, Bt, ct = cr.exact_tv_matrices(orc, x_trj, u0)
    return cfg, orc, At, Bt, ct


def test_admm_matches_dense_qp_and_respects_bounds():
    T = 10
    cfg, orc, At, Bt, ct = _problem(T)
    xlo, xhi = np.array([-1e4, -1e4, -1e4, -1e4, -0.3]), np.array([1e4, 1e4, 1e4, 1e4, 0.3])
    ulo, uhi = np.array([-1e4, -0.4]), np.array([1e4, 0.4])
    x, u, it = bq.admm_box_qp(At, Bt, ct, cfg["Q"], cfg["Qd"], cfg["R"], cfg["x0"], cfg["xd_trj"], xlo, xhi, ulo, uhi)
    xr, ur, res = bq.dense_qp_reference(At, Bt, ct, cfg["Q"], cfg["Qd"], cfg["R"], cfg["x0"], cfg["xd_trj"], xlo, xhi,
                                        ulo, uhi)
    assert res.success and it < bq.DEFAULT_MAX_ITER
    np.testing.assert_allclose(u, ur, rtol=0, atol=2e-5)
    np.testing.assert_allclose(x, xr, rtol=0, atol=2e-5)
    assert np.max(np.abs(x[1:, 4])) <= 0.3 + 1e-6 and np.max(np.abs(u[:, 1])) <= 0.4 + 1e-6
    assert np.max(np.abs(x[:, 4])) > 0.29


def test_inactive_bounds_reduce_to_the_riccati_solution():
    T = 15
    cfg, orc, At, Bt, ct = _problem(T, (0.1, 0.0))
    big = 1e6
    x, u, _ = bq.admm_box_qp(At, Bt, ct, cfg["Q"], cfg["Qd"], cfg["R"], cfg["x0"], cfg["xd_trj"], -big * np.ones(5),
                             big * np.ones(5), -big * np.ones(2), big * np.ones(2))
    xs, us = cr.solve_tvlqr(At, Bt, ct, cfg["Q"], cfg["Qd"], cfg["R"], cfg["x0"], cfg["xd_trj"])
    np.testing.assert_allclose(u, us, rtol=0, atol=1e-6)
    np.testing.assert_allclose(x, xs, rtol=0, atol=1e-6)


@pytest.mark.parametrize("rho0", [0.2, 5.0])
def test_mpc_descent_is_independent_of_the_penalty(rho0):
    """The closed-loop result is the sequence of QP minimisers, not an artefact of the ADMM penalty."""
    T = 20
    cfg = ec.bicycle(T=T)
    orc = cr.BicycleOracle(cfg["h"])
    x_trj = cr.rollout(orc, cfg["x0"], cfg["u_trj_initial"])
    At, Bt, ct = cr.exact_tv_matrices(orc, x_trj, cfg["u_trj_initial"])
    args = (orc, At, Bt, ct, cfg["Q"], cfg["Qd"], cfg["R"], cfg["x0"], cfg["xd_trj"], cfg["xbound"][0],
            cfg["xbound"][1], cfg["ubound"][0], cfg["ubound"][1])
    x1, u1, _ = bq.mpc_box_descent(*args)
    x2, u2, _ = bq.mpc_box_descent(*args, rho0=rho0)
    np.testing.assert_allclose(u1, u2, rtol=0, atol=2e-6)
    np.testing.assert_allclose(x1, x2, rtol=0, atol=2e-6)
    assert np.max(np.abs(x1[:, 4])) <= np.pi / 4 + 1e-6
    assert cr.evaluate_cost(x1, u1, cfg["xd_trj"], cfg["Q"], cfg["R"]) < \
        cr.evaluate_cost(x_trj, cfg["u_trj_initial"], cfg["xd_trj"], cfg["Q"], cfg["R"])
