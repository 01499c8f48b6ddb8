"""GPU: learned dynamics (the reference's examples/pendulum/pendulum_nn.py) through the
same DynamicalSystem / IrsLqr call surface as the analytic systems.  Checked against torch's own outputs for
the committed network, against the REFERENCE's zero-order linearization of it (tests/golden/mlp_pendulum.npz,
oracle/make_mlp_fixture.py) and against the oracle on identical noise.  Tolerance 1e-4 relative for what
passes through the fp32 sample path, 1e-5 for single float32 network evaluations (sums in another order)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import cpu_restatement as cr          # noqa: E402
from oracle import example_configs as ec          # noqa: E402
from oracle.mlp_ref import MlpOracle              # noqa: E402

FP32_RTOL = 1e-4
KEYS = ("W1", "b1", "W2", "b2", "W3", "b3")


def rel_err(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(float(np.max(np.abs(b))), 1e-300))


@pytest.fixture(scope="module")
def api():
    import torch
    assert torch.cuda.is_available()
    import irs_mpc_b200.all as m
    return m


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "mlp_pendulum.npz"))


@pytest.fixture(scope="module")
def system(api, gold):
    w = [gold[k] for k in KEYS]
    return api.MlpDynamics([(w[0], w[1]), (w[2], w[3]), (w[4], w[5])])


def make_params(api, cfg, T, u_trj=None, x0=None):
    p = api.IrsLqrParameters()
    p.Q, p.Qd, p.R = cfg["Q"], cfg["Qd"], cfg["R"]
    p.x0 = cfg["x0"] if x0 is None else x0
    p.xd_trj = cfg["xd_trj"][:T + 1]
    p.u_trj_initial = cfg["u_trj_initial"][:T] if u_trj is None else u_trj
    p.xbound, p.ubound = cfg["xbound"], cfg["ubound"]
    return p


def test_dynamics_match_torch(system, gold):
    out = system.dynamics_batch(gold["pts"][:, :2], gold["pts"][:, 2:])
    assert out.shape == (64, 2) and out.dtype == np.float64
    assert rel_err(out, gold["out"]) < 1e-5
    np.testing.assert_array_equal(system.dynamics(gold["pts"][7, :2], gold["pts"][7, 2:]), out[7])


def test_jacobian_matches_autograd(system, gold):
    J = system.jacobian_xu_batch(gold["pts"][:, :2], gold["pts"][:, 2:])
    assert J.shape == (64, 2, 3)
    assert rel_err(J, gold["jac"]) < 1e-5
    np.testing.assert_array_equal(system.jacobian_xu(gold["pts"][9, :2], gold["pts"][9, 2:]), J[9])


def test_rollout_by_the_warp_equals_stepping_the_per_thread_functor(api, system):
    """IrsLqr.rollout (irs_lqr.py:105-119) runs the network with one warp per trajectory; `dynamics` with one thread
    per point: every hidden unit is summed in the same order, so the trajectories are bit-identical."""
    T = 40
    cfg = ec.pendulum_nn(T=T)
    u_trj = cfg["u_trj_initial"] + np.random.default_rng(2).standard_normal((T, 1))
    solver = api.IrsLqrExact(system, make_params(api, cfg, T, u_trj=u_trj))
    x = np.zeros((T + 1, 2))
    x[0] = cfg["x0"]
    for t in range(T):
        x[t + 1] = system.dynamics(x[t], u_trj[t])
    np.testing.assert_array_equal(solver.rollout(cfg["x0"], u_trj), x)


def test_from_a_torch_module_and_refuses_other_architectures(api, gold):
    import torch
    import torch.nn as nn

    class DynamicsNLP(nn.Module):                 # as written in pendulum_nn.py:19-33
        def __init__(self, act):
            super().__init__()
            self.dynamics_mlp = nn.Sequential(nn.Linear(3, 100), act(), nn.Linear(100, 100), act(), nn.Linear(100, 2))

        def forward(self, x):
            return self.dynamics_mlp(x)

    net = DynamicsNLP(nn.ReLU)
    with torch.no_grad():
        for layer, (wk, bk) in zip((0, 2, 4), (("W1", "b1"), ("W2", "b2"), ("W3", "b3"))):
            net.dynamics_mlp[layer].weight.copy_(torch.from_numpy(gold[wk]))
            net.dynamics_mlp[layer].bias.copy_(torch.from_numpy(gold[bk]))
    s = api.MlpDynamics(net.eval())
    assert rel_err(s.dynamics_batch(gold["pts"][:, :2], gold["pts"][:, 2:]), gold["out"]) < 1e-5
    with pytest.raises(RuntimeError, match="Linear-ReLU-Linear-ReLU-Linear"):
        api.MlpDynamics(DynamicsNLP(nn.Tanh).eval())
    with pytest.raises(RuntimeError, match="three linear layers"):
        api.MlpDynamics(nn.Sequential(nn.Linear(3, 8), nn.ReLU(), nn.Linear(8, 2)))
    with pytest.raises(RuntimeError, match="dim_x = 2"):
        api.MlpDynamics(net, dim_x=3, dim_u=1)


def test_zero_order_replay_matches_the_reference(api, system, gold):
    """The reference's IrsLqrZeroOrder.get_TV_matrices on its PendulumNN wrapper, same replayed noise."""
    T, N, d = gold["noise_shape"]
    noise = np.random.default_rng(int(gold["noise_seed"])).standard_normal((T, N, d)).astype(np.float32)
    state = {"t": 0}

    def sampling(xbar, ubar, it):
        e = noise[state["t"] % T].astype(np.float64)
        state["t"] += 1
        return e[:, :2], e[:, 2:]

    cfg = ec.pendulum_nn(T=int(T))
    solver = api.IrsLqrZeroOrder(system, make_params(api, cfg, int(T)), sampling)
    assert rel_err(solver.x_trj, gold["rollout_x"]) < 1e-5
    assert abs(solver.cost - float(gold["initial_cost"])) < 1e-5 * float(gold["initial_cost"])
    At, Bt, ct = solver.get_TV_matrices(gold["x_trj"], gold["u_trj"])
    assert rel_err(At, gold["At"]) < FP32_RTOL
    assert rel_err(Bt, gold["Bt"]) < FP32_RTOL
    assert rel_err(ct, gold["ct"]) < FP32_RTOL


@pytest.mark.parametrize("antithetic", [True, False])
@pytest.mark.parametrize("N", [1000, 4097])
def test_zero_order_philox_matches_oracle_on_same_deltas(api, system, gold, N, antithetic):
    T = 5
    cfg = ec.pendulum_nn(T=T)
    rng = np.random.default_rng(3)
    u_trj = cfg["u_trj_initial"] + rng.standard_normal((T, 1))
    sampler = api.GaussianSampling(cfg["sigma"][:2], cfg["sigma"][2:], N, power=cfg["power"], seed=17, antithetic=antithetic)
    solver = api.IrsLqrZeroOrder(system, make_params(api, cfg, T, u_trj=u_trj), sampler)
    solver.iter = 2
    At, Bt, ct = solver.get_TV_matrices(solver.x_trj, solver.u_trj)
    deltas = sampler.deltas(T, solver.iter).astype(np.float64)
    orc = MlpOracle([gold[k] for k in KEYS])
    At_o, Bt_o, ct_o = cr.zero_order_tv_matrices(orc, solver.x_trj, solver.u_trj, deltas)
    assert rel_err(At, At_o) < FP32_RTOL
    assert rel_err(Bt, Bt_o) < FP32_RTOL
    assert rel_err(ct, ct_o) < FP32_RTOL


@pytest.mark.parametrize("N", [37, 128, 1000, 5000])
def test_tensor_core_engine_matches_the_per_thread_functor(api, system, gold, N, monkeypatch):
    """The hidden layer on the tensor cores (bf16 split, smooth_mlp.cuh) against the generic kernel that evaluates
    the network per thread in float32 (IRS_MLP_ENGINE=0), same Philox noise; ragged tiles and chunks included."""
    from irs_mpc_b200 import _graph
    monkeypatch.setattr(_graph, "USE_GRAPHS", False)
    T = 7
    cfg = ec.pendulum_nn(T=T)
    u_trj = cfg["u_trj_initial"] + np.random.default_rng(11).standard_normal((T, 1))
    res = {}
    for engine in ("1", "0"):
        monkeypatch.setenv("IRS_MLP_ENGINE", engine)
        sampler = api.GaussianSampling(cfg["sigma"][:2], cfg["sigma"][2:], N, power=cfg["power"], seed=31)
        solver = api.IrsLqrZeroOrder(system, make_params(api, cfg, T, u_trj=u_trj), sampler)
        res[engine] = solver.get_TV_matrices(solver.x_trj, solver.u_trj)
    for got, want in zip(res["1"], res["0"]):
        assert rel_err(got, want) < FP32_RTOL


@pytest.mark.parametrize("H1,H2", [(16, 16), (64, 32), (100, 37), (111, 112), (128, 128), (5, 1)])
def test_other_hidden_widths(api, H1, H2, monkeypatch):
    """Random networks of other widths (operand padding, TMEM columns beyond 128 for H1 = 128, several tile groups
    per block): tensor-core engine against the oracle on the same Philox deltas and against the per-thread engine."""
    from irs_mpc_b200 import _graph
    monkeypatch.setattr(_graph, "USE_GRAPHS", False)
    rng = np.random.default_rng(100 * H1 + H2)
    layers = [((rng.standard_normal((H1, 3)) / np.sqrt(3)).astype(np.float32), (0.3 * rng.standard_normal(H1)).astype(np.float32)),
              ((rng.standard_normal((H2, H1)) / np.sqrt(H1)).astype(np.float32), (0.3 * rng.standard_normal(H2)).astype(np.float32)),
              ((rng.standard_normal((2, H2)) / np.sqrt(H2)).astype(np.float32), (0.3 * rng.standard_normal(2)).astype(np.float32))]
    s = api.MlpDynamics(layers)
    orc = MlpOracle([a for pair in layers for a in pair])
    pts = rng.standard_normal((33, 3))
    assert rel_err(s.dynamics_batch(pts[:, :2], pts[:, 2:]), orc.dynamics_batch(pts[:, :2], pts[:, 2:])) < 1e-5
    assert rel_err(s.jacobian_xu_batch(pts[:, :2], pts[:, 2:]), orc.jacobian_xu_batch(pts[:, :2], pts[:, 2:])) < 1e-5
    T, N = 4, 3000
    cfg = ec.pendulum_nn(T=T)
    u_trj = cfg["u_trj_initial"] + rng.standard_normal((T, 1))
    res = {}
    for engine in ("1", "0"):
        monkeypatch.setenv("IRS_MLP_ENGINE", engine)
        sampler = api.GaussianSampling(cfg["sigma"][:2], cfg["sigma"][2:], N, power=cfg["power"], seed=41)
        solver = api.IrsLqrZeroOrder(s, make_params(api, cfg, T, u_trj=u_trj), sampler)
        res[engine] = solver.get_TV_matrices(solver.x_trj, solver.u_trj)
        deltas = sampler.deltas(T, solver.iter).astype(np.float64)
        x_trj = solver.x_trj
    want = cr.zero_order_tv_matrices(orc, x_trj, u_trj, deltas)
    for got, w, other in zip(res["1"], want, res["0"]):
        assert rel_err(got, w) < FP32_RTOL
        assert rel_err(got, other) < FP32_RTOL
    with pytest.raises(Exception, match="hidden widths"):
        api.MlpDynamics([(np.zeros((129, 3), np.float32), np.zeros(129, np.float32)),
                         (np.zeros((8, 129), np.float32), np.zeros(8, np.float32)),
                         (np.zeros((2, 8), np.float32), np.zeros(2, np.float32))])


def test_first_order_and_exact_match_oracle(api, system, gold):
    T, N = 5, 1500
    cfg = ec.pendulum_nn(T=T)
    orc = MlpOracle([gold[k] for k in KEYS])
    sampler = api.GaussianSampling(cfg["sigma"][:2], cfg["sigma"][2:], N, power=cfg["power"], seed=23)
    solver = api.IrsLqrFirstOrder(system, make_params(api, cfg, T), sampler)
    At, Bt, ct = solver.get_TV_matrices(solver.x_trj, solver.u_trj)
    deltas = sampler.deltas(T, solver.iter).astype(np.float64)
    At_o, Bt_o, ct_o = cr.first_order_tv_matrices(orc, solver.x_trj, solver.u_trj, deltas)
    assert rel_err(At, At_o) < FP32_RTOL and rel_err(Bt, Bt_o) < FP32_RTOL and rel_err(ct, ct_o) < FP32_RTOL
    exact = api.IrsLqrExact(system, make_params(api, cfg, T))
    At, Bt, ct = exact.get_TV_matrices(exact.x_trj, exact.u_trj)
    At_o, Bt_o, ct_o = cr.exact_tv_matrices(orc, exact.x_trj, exact.u_trj)
    assert rel_err(At, At_o) < 1e-5 and rel_err(Bt, Bt_o) < 1e-5 and rel_err(ct, ct_o) < 1e-5


def test_descent_matches_oracle_and_iterations_lower_the_cost(api, system, gold):
    """pendulum_nn.py:141-156: IrsLqrZeroOrder on the learned system.  One descent from the same trajectory with
    the same deltas against the oracle; then the script's loop (fewer iterations) must lower the cost."""
    T, N = 200, 10000
    cfg = ec.pendulum_nn(T=T)
    sampler = api.GaussianSampling(cfg["sigma"][:2], cfg["sigma"][2:], N, power=cfg["power"], seed=5)
    solver = api.IrsLqrZeroOrder(system, make_params(api, cfg, T), sampler)
    x_new, u_new = solver.local_descent(solver.x_trj, solver.u_trj)
    cost = solver.evaluate_cost(x_new, u_new)
    orc = MlpOracle([gold[k] for k in KEYS])
    deltas = sampler.deltas(T, solver.iter).astype(np.float64)
    At, Bt, ct = cr.zero_order_tv_matrices(orc, solver.x_trj, solver.u_trj, deltas)
    K, k = cr.tvlqr_riccati(At, Bt, ct, cfg["Q"], cfg["Qd"], cfg["R"], cfg["xd_trj"])
    x_o, u_o = cr.closed_loop_descent(orc, K, k, solver.x_trj[0])
    cost_o = cr.evaluate_cost(x_o, u_o, cfg["xd_trj"], cfg["Q"], cfg["R"])
    assert rel_err(x_new, x_o) < 5 * FP32_RTOL
    assert abs(cost - cost_o) / abs(cost_o) < 5 * FP32_RTOL
    first = solver.cost
    solver.iterate(6)
    assert len(solver.cost_lst) == 8 and solver.cost < 0.6 * first


def test_batched_instances_and_cem_run_on_the_learned_system(api, system, gold):
    """The instance-batched path (one warp per instance in the Riccati and rollout kernels) and the CEM baseline
    (one warp per candidate) on the learned system: every instance equals its own single-instance solver bit for bit."""
    T, N, I = 30, 2000, 5
    cfg = ec.pendulum_nn(T=T)
    rng = np.random.default_rng(8)
    x0 = 0.3 * rng.standard_normal((I, 2))
    smp = api.GaussianSampling(cfg["sigma"][:2], cfg["sigma"][2:], N, power=cfg["power"], seed=77)
    bat = api.BatchedIrsLqrZeroOrder(system, cfg["Q"], cfg["Qd"], cfg["R"], x0, cfg["xd_trj"], cfg["u_trj_initial"], smp)
    xb, ub, cb = bat.iterate(1)
    for b in (0, 3):
        smp1 = api.GaussianSampling(cfg["sigma"][:2], cfg["sigma"][2:], N, power=cfg["power"], seed=77, stream_id=0)
        one = api.BatchedIrsLqrZeroOrder(system, cfg["Q"], cfg["Qd"], cfg["R"], x0[b:b + 1], cfg["xd_trj"],
                                         cfg["u_trj_initial"], smp1, instance_offset=b)
        x1, u1, c1 = one.iterate(1)
        np.testing.assert_array_equal(x1[0], xb[b])
        np.testing.assert_array_equal(u1[0], ub[b])
    assert np.all(np.isfinite(cb))
    prm = api.CemParameters()
    prm.Q, prm.Qd, prm.R, prm.x0, prm.xd_trj = cfg["Q"], cfg["Qd"], cfg["R"], cfg["x0"], cfg["xd_trj"]
    prm.u_trj_initial = cfg["u_trj_initial"]
    prm.n_elite, prm.batch_size, prm.initial_std = 10, 64, np.array([1.0])
    np.random.seed(3)
    cem = api.CrossEntropyMethod(system, prm)
    first = cem.cost
    cem.iterate(4, verbose=False)
    assert len(cem.cost_lst) == 6 and min(cem.cost_lst[1:]) < first
    orc = MlpOracle([gold[k] for k in KEYS])
    x_o = cr.rollout(orc, cfg["x0"], cem.u_trj)
    assert rel_err(cem.x_trj, x_o) < 1e-4
