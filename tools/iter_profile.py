"""Run a few iRS-LQR iterations (quadrotor cfg3) — target for `ncu --metrics gpu__time_duration.sum`.

    python tools/iter_profile.py [iterations]
"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from irs_mpc_b200 import example_configs as ec                                      # noqa: E402
from irs_mpc_b200.all import (GaussianSampling, IrsLqrParameters, IrsLqrZeroOrder,  # noqa: E402
                              QuadrotorDynamics)

cfg = ec.quadrotor(T=100)
system = QuadrotorDynamics(cfg["h"])
params = IrsLqrParameters()
for key in ("Q", "Qd", "R", "x0", "xd_trj", "u_trj_initial", "xbound", "ubound"):
    setattr(params, key, cfg[key])
sampler = GaussianSampling(cfg["sigma"][:12], cfg["sigma"][12:], 100000, power=cfg["power"], seed=1)
solver = IrsLqrZeroOrder(system, params, sampler)
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 5
for k in range(2):
    solver.local_descent(solver.x_trj, solver.u_trj)
torch.cuda.synchronize()
t = time.perf_counter()
for k in range(iters):
    xn, un = solver.local_descent(solver.x_trj, solver.u_trj)
    c = solver.evaluate_cost(xn, un)
torch.cuda.synchronize()
print("%.3f ms per iteration (local_descent + evaluate_cost), cost %.6f" % (1e3 * (time.perf_counter() - t) / iters, c))
