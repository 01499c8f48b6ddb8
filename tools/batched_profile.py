import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from irs_mpc_b200.all import GaussianSampling, QuadrotorDynamics, BatchedIrsLqrZeroOrder
from irs_mpc_b200 import example_configs as ec
cfg = ec.CONFIGS["quadrotor"](T=100)
system = QuadrotorDynamics(cfg["h"])
I, T = 4096, 100
ph = 2.0 * np.pi * np.arange(I) / I
tt = np.arange(T + 1, dtype=np.float64)
xdb = np.zeros((I, T + 1, 12))
xdb[:, :, 0] = 1.5 * np.cos(0.05 * tt[None, :] + ph[:, None]); xdb[:, :, 1] = 1.5 * np.sin(0.05 * tt[None, :] + ph[:, None]); xdb[:, :, 2] = 0.02 * tt[None, :]
X0 = np.array([0.1] * 3 + [0.01] * 3 + [0.1] * 3 + [0.01] * 3)
x0b = xdb[:, 0, :] + X0 * np.random.default_rng(5).standard_normal((I, 12))
smp = GaussianSampling(cfg["sigma"][:12], cfg["sigma"][12:], 1000, power=cfg["power"], seed=77)
bat = BatchedIrsLqrZeroOrder(system, cfg["Q"], cfg["Qd"], cfg["R"], x0b, xdb, cfg["u_trj_initial"], smp)
for k in range(2):
    bat.local_descent()
torch.cuda.synchronize()
bat.check()
