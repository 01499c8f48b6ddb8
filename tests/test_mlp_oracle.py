"""CPU: the restatement of the reference's learned dynamics (oracle/mlp_ref.py) against what torch — the
reference's own evaluator of that system, examples/pendulum/pendulum_nn.py:66-90 — returned for the committed
network (tests/golden/mlp_pendulum.npz, written by oracle/make_mlp_fixture.py), and against the
REFERENCE's IrsLqrZeroOrder.get_TV_matrices driven by that system."""
import os

import numpy as np

from oracle import cpu_restatement as cr
from oracle import example_configs as ec
from oracle.mlp_ref import MlpOracle

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mlp_pendulum.npz")


def load():
    g = np.load(GOLD)
    return g, MlpOracle([g[k] for k in ("W1", "b1", "W2", "b2", "W3", "b3")])


def test_network_shape_is_the_references():
    g, _ = load()
    assert g["W1"].shape == (100, 3) and g["W2"].shape == (100, 100) and g["W3"].shape == (2, 100)   # pendulum_nn.py:23-29
    assert all(g[k].dtype == np.float32 for k in ("W1", "b1", "W2", "b2", "W3", "b3"))
    assert float(g["train_loss"]) < 0.05          # it did learn the pendulum (targets are O(10))


def test_forward_matches_torch():
    g, o = load()
    out = o.dynamics_batch(g["pts"][:, :2], g["pts"][:, 2:])
    assert out.dtype == np.float64
    assert np.max(np.abs(out - g["out"])) < 2e-6 * np.max(np.abs(g["out"]))      # float32 sums in another order
    one = o.dynamics(g["pts"][3, :2], g["pts"][3, 2:])
    np.testing.assert_array_equal(one, out[3])


def test_jacobian_matches_autograd():
    g, o = load()
    J = o.jacobian_xu_batch(g["pts"][:, :2], g["pts"][:, 2:])
    assert J.shape == (64, 2, 3)
    assert np.max(np.abs(J - g["jac"])) < 2e-6 * np.max(np.abs(g["jac"]))
    np.testing.assert_array_equal(o.jacobian_xu(g["pts"][5, :2], g["pts"][5, 2:]), J[5])


def test_zero_order_linearization_matches_the_reference():
    g, o = load()
    T, N, d = g["noise_shape"]
    noise = np.random.default_rng(int(g["noise_seed"])).standard_normal((T, N, d)).astype(np.float32).astype(np.float64)
    At, Bt, ct = cr.zero_order_tv_matrices(o, g["x_trj"], g["u_trj"], noise)
    for got, key in ((At, "At"), (Bt, "Bt"), (ct, "ct")):
        assert np.max(np.abs(got - g[key])) < 1e-5 * np.max(np.abs(g[key])), key


def test_initial_rollout_and_cost_match_the_reference():
    g, o = load()
    T = g["u_trj"].shape[0]
    cfg = ec.pendulum_nn(T=T)
    x = cr.rollout(o, cfg["x0"], cfg["u_trj_initial"])
    assert np.max(np.abs(x - g["rollout_x"])) < 1e-5 * max(1.0, np.max(np.abs(g["rollout_x"])))
    cost = cr.evaluate_cost(x, cfg["u_trj_initial"], cfg["xd_trj"], cfg["Q"], cfg["R"])
    assert abs(cost - float(g["initial_cost"])) < 1e-5 * float(g["initial_cost"])
