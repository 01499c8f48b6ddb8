"""Writes tests/golden/mlp_pendulum.npz (run in the BUILD container, where /root/reference and torch exist).

Follows examples/pendulum/pendulum_nn.py:19-62 of the reference: the same network class, the same
artificial data (20000 uniform points, targets from the reference's own PendulumDynamics.dynamics_batch —
the script's `dynamics_batch_np` no longer exists in pendulum_dynamics.py), Adam lr 1e-3 with StepLR(500),
600 full-batch iterations, seeded here.  Stores the trained float32 weights together with what torch — the
reference's evaluator of this system — returns for them:
  out [64, 2]      net(xu)                       (pendulum_nn.py:72-81)
  jac [64, 2, 3]   torch.autograd.grad per output (pendulum_nn.py:83-90)
and the reference's IrsLqrZeroOrder.get_TV_matrices (irs_lqr_zero_order.py:38-63, unmodified code) driven by
the PendulumNN wrapper on replayed noise: At, Bt, ct for T = 8 nominal points, N = 4000 samples.
"""
import os
import sys

import numpy as np
import torch
import torch.nn as nn
import torch.optim as optim

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_import      # noqa: E402


def main():
    ref_import.install()
    from pendulum_dynamics import PendulumDynamics                     # the reference's
    from irs_lqr.dynamical_system import DynamicalSystem
    from irs_lqr.irs_lqr import IrsLqrParameters
    from irs_lqr.irs_lqr_zero_order import IrsLqrZeroOrder

    np.random.seed(12345)
    torch.manual_seed(12345)
    torch.set_num_threads(4)

    class DynamicsNLP(nn.Module):                                      # pendulum_nn.py:19-33
        def __init__(self):
            super().__init__()
            self.dynamics_mlp = nn.Sequential(nn.Linear(3, 100), nn.ReLU(), nn.Linear(100, 100), nn.ReLU(),
                                              nn.Linear(100, 2))

        def forward(self, x):
            return self.dynamics_mlp(x)

    pendulum = PendulumDynamics(0.05)
    num_data = 20000                                                   # pendulum_nn.py:40-47
    xu = np.random.rand(num_data, 3)
    xu[:, 0] = 6 * np.pi * (xu[:, 0] - 0.5)
    xu[:, 1] = 30.0 * (xu[:, 1] - 0.5)
    xu[:, 2] = 30.0 * (xu[:, 2] - 0.5)
    xtarget = pendulum.dynamics_batch(xu[:, 0:2], xu[:, 2, None])

    net = DynamicsNLP()                                                # pendulum_nn.py:49-62
    net.train()
    optimizer = optim.Adam(net.parameters(), lr=0.001)
    scheduler = optim.lr_scheduler.StepLR(optimizer, step_size=500)
    criterion = nn.MSELoss()
    for it in range(600):
        optimizer.zero_grad()
        loss = criterion(net(torch.Tensor(xu)), torch.Tensor(xtarget))
        loss.backward()
        optimizer.step()
        scheduler.step()
    net.eval()
    print("training loss", float(loss))

    class PendulumNN(DynamicalSystem):                                 # pendulum_nn.py:66-90 (methods given `self`)
        def __init__(self):
            super().__init__()
            self.dim_x = 2
            self.dim_u = 1

        def dynamics(self, x, u):
            xu = torch.Tensor(np.concatenate((x, u))).unsqueeze(0)
            return net(xu).detach().numpy()[0]

        def dynamics_batch(self, x, u):
            xu = torch.Tensor(np.hstack((x, u)))
            return net(xu).detach().numpy()

        def jacobian_xu(self, x, u):
            xu = torch.Tensor(np.concatenate((x, u))).unsqueeze(0)
            xu.requires_grad = True
            xnext = net(xu)
            d0 = torch.autograd.grad(xnext[0, 0], xu, retain_graph=True)[0].numpy()
            d1 = torch.autograd.grad(xnext[0, 1], xu)[0].numpy()
            return np.vstack((d0, d1))[0:2]

    system = PendulumNN()
    rng = np.random.default_rng(7)
    pts = np.column_stack((rng.uniform(-3 * np.pi, 3 * np.pi, 64), rng.uniform(-15, 15, 64), rng.uniform(-15, 15, 64)))
    out = system.dynamics_batch(pts[:, :2], pts[:, 2:])
    jac = np.stack([system.jacobian_xu(p[:2], p[2:]) for p in pts])

    # the reference's zero-order linearization of this system on replayed noise
    T, N = 8, 4000
    params = IrsLqrParameters()                                        # pendulum_nn.py:95-110
    params.Q = np.diag([1, 1]);  params.Qd = np.diag([20., 20.]);  params.R = np.diag([1])
    params.x0 = np.array([0, 0]);  params.xd_trj = np.tile(np.array([np.pi, 0]), (T + 1, 1))
    params.xbound = [-np.array([1e4, 1e4]), np.array([1e4, 1e4])]
    params.ubound = np.array([-np.array([1e4]), np.array([1e4])])
    params.u_trj_initial = np.tile(np.array([0.1]), (T, 1))
    noise = np.random.default_rng(1004).standard_normal((T, N, 3)).astype(np.float32)      # tests regenerate it
    state = {"t": 0}

    def sampling(xbar, ubar, it):
        e = noise[state["t"] % T].astype(np.float64)
        state["t"] += 1
        return e[:, :2], e[:, 2:]

    solver = IrsLqrZeroOrder(system, params, sampling)
    x_trj = np.column_stack((np.linspace(0.0, 2.5, T + 1), np.linspace(0.0, 3.0, T + 1)))
    u_trj = np.linspace(-2.0, 2.0, T)[:, None]
    At, Bt, ct = solver.get_TV_matrices(x_trj, u_trj)

    sd = net.dynamics_mlp
    np.savez_compressed(
        os.path.join(ROOT, "tests", "golden", "mlp_pendulum.npz"),
        W1=sd[0].weight.detach().numpy(), b1=sd[0].bias.detach().numpy(),
        W2=sd[2].weight.detach().numpy(), b2=sd[2].bias.detach().numpy(),
        W3=sd[4].weight.detach().numpy(), b3=sd[4].bias.detach().numpy(),
        pts=pts, out=out, jac=jac, noise_seed=np.int64(1004), noise_shape=np.array([T, N, 3]), x_trj=x_trj, u_trj=u_trj, At=At, Bt=Bt, ct=ct,
        initial_cost=np.float64(solver.cost), rollout_x=solver.x_trj, train_loss=np.float64(float(loss)))
    print("wrote tests/golden/mlp_pendulum.npz")


if __name__ == "__main__":
    main()
