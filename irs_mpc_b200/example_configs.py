"""Problem data of the reference example scripts (L4 of the reference: examples/<system>/*.py).

Each entry cites the script it mirrors (paths relative to /root/reference/examples).  Returned as a
plain dict so that both the oracle (float64 numpy) and the CUDA path can be built from it.
"""
import numpy as np


def pendulum(T=200):
    # pendulum/pendulum_zero_order.py:11-36, pendulum/pendulum_exact.py:11-31
    return dict(system="pendulum", h=0.05, T=T,
                Q=np.diag([1., 1.]), Qd=np.diag([20., 20.]), R=np.diag([1.]),
                x0=np.array([0., 0.]), xd_trj=np.tile(np.array([np.pi, 0.]), (T + 1, 1)),
                u_trj_initial=np.tile(np.array([0.1]), (T, 1)),
                xbound=[-np.array([1e4, 1e4]), np.array([1e4, 1e4])],
                ubound=np.array([-np.array([1e4]), np.array([1e4])]),
                sigma=np.array([1.0, 1.0, 1.0]), power=0.5, num_samples=1000, projection=False)


def bicycle(T=100):
    # bicycle/bicycle_first_order.py:11-36 ("easy" goal), bicycle/bicycle_exact.py:12-33
    return dict(system="bicycle", h=0.1, T=T,
                Q=np.diag([5, 5, 3, 0.1, 0.1]), Qd=np.diag([50., 50, 30, 1, 1]),
                R=np.diag([1, 0.1]), x0=np.zeros(5),
                xd_trj=np.tile(np.array([3.0, 1.0, np.pi / 2, 0, 0]), (T + 1, 1)),
                u_trj_initial=np.tile(np.array([0.1, 0.0]), (T, 1)),
                xbound=[-np.array([1e4, 1e4, 1e4, 1e4, np.pi / 4]),
                        np.array([1e4, 1e4, 1e4, 1e4, np.pi / 4])],
                ubound=np.array([-np.array([1e4, 1e4]), np.array([1e4, 1e4])]),
                sigma=np.array([2.0, 2.0, 1.0, 2.0, 0.01, 2.0, 1.0]), power=0.5,
                num_samples=10000, projection=False)


def quadrotor(T=200):
    # quadrotor/quadrotor_zero_order.py:12-52, quadrotor/quadrotor_exact.py:12-41
    i = np.arange(T + 1, dtype=np.float64)
    xd = np.zeros((T + 1, 12))
    xd[:, 0] = 1.5 * np.cos(0.05 * i)
    xd[:, 1] = 1.5 * np.sin(0.05 * i)
    xd[:, 2] = 0.02 * i
    big = np.array([1e5, 1e5, 1e5, 2.0 * np.pi, np.pi / 2, 2.0 * np.pi, 1e5, 1e5, 1e5, 1e5, 1e5, 1e5])
    return dict(system="quadrotor", h=0.05, T=T,
                Q=1.0 * np.diag([10., 10, 10, 10, 10, 10, 0, 0, 0, 0, 0, 0]),
                Qd=10.0 * np.diag([10., 10, 10, 10, 10, 10, 1, 1, 1, 1, 1, 1]),
                R=1.0 * np.diag([1., 1, 1, 1]), x0=np.zeros(12), xd_trj=xd,
                u_trj_initial=np.tile(np.array([2.0, 2.0, 2.0, 2.0]), (T, 1)),
                xbound=[-big, big],
                ubound=np.array([-1e5 * np.ones(4), 1e5 * np.ones(4)]),
                sigma=0.1 * np.ones(16), power=0.5, num_samples=1000, projection=False)


def three_cart(T=100):
    # three_cart/three_cart_zero_order.py:11-43
    return dict(system="three_cart", h=0.05, T=T,
                Q=0.01 * np.diag([50., 50, 50, 20, 100, 20]), Qd=np.diag([50., 50, 50, 20, 100, 20]),
                R=0.01 * np.diag([1., 1]), x0=np.array([0., 1, 2, 0, 0, 0]),
                xd_trj=np.tile(np.array([2., 3, 4, 0, 0, 0]), (T + 1, 1)),
                u_trj_initial=np.tile(np.array([0.1, -0.1]), (T, 1)),
                xbound=[-1e4 * np.ones(6), 1e4 * np.ones(6)],
                ubound=np.array([-1000. * np.ones(2), 1000. * np.ones(2)]),
                sigma=np.array([4.0, 4, 4, 4, 4, 4, 0.5, 0.5]), power=0.2, num_samples=1000,
                projection=True)


def pendulum_nn(T=200):
    # pendulum/pendulum_nn.py:92-110 (problem), :141-150 (sampling: sigma 1 / iter^0.5, 10000 samples); the
    # system is the learned network of :19-62, no step size of its own
    cfg = pendulum(T)
    cfg.update(system="pendulum_nn", h=0.0, num_samples=10000)
    return cfg


def quadrotor_batch(lo, hi, T=100, total=4096):
    """BASELINE.json configs[4]: instances lo..hi-1 of `total` independent quadrotor MPC problems (the
    reference runs one IrsLqr object per problem, irs_lqr/irs_lqr.py:34-71).  Instance b tracks the helix
    of quadrotor/quadrotor_zero_order.py:23-27 with phase 2 pi b / total and starts on it plus
    N(0, diag(scale^2)) from default_rng(5000 + b): 0.1 on positions / velocities, 0.01 on angles /
    angular rates (larger tilt noise makes the uncontrolled initial rollout tumble).
    Returns (x0 [I,12], xd_trj [I,T+1,12])."""
    ph = 2.0 * np.pi * np.arange(lo, hi) / total
    tt = np.arange(T + 1, dtype=np.float64)
    xd = np.zeros((hi - lo, T + 1, 12))
    xd[:, :, 0] = 1.5 * np.cos(0.05 * tt[None, :] + ph[:, None])
    xd[:, :, 1] = 1.5 * np.sin(0.05 * tt[None, :] + ph[:, None])
    xd[:, :, 2] = 0.02 * tt[None, :]
    scale = np.array([0.1] * 3 + [0.01] * 3 + [0.1] * 3 + [0.01] * 3)
    x0 = xd[:, 0, :] + np.stack([scale * np.random.default_rng(5000 + b).standard_normal(12)
                                 for b in range(lo, hi)])
    return x0, xd


CONFIGS = {"pendulum": pendulum, "bicycle": bicycle, "quadrotor": quadrotor,
           "three_cart": three_cart}
