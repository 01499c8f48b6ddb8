"""BASELINE.json configs[4] (4096 quadrotor MPC instances, T=100, N=1000): ms per batch iteration and per
kernel class (tuning aid)."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from irs_mpc_b200 import _device, example_configs as ec                                    # noqa: E402
from irs_mpc_b200.all import BatchedIrsLqrZeroOrder, GaussianSampling, QuadrotorDynamics   # noqa: E402
from irs_mpc_b200.tv_lqr import riccati_device                                             # noqa: E402

I, T = int(os.environ.get("BB_I", "4096")), 100
cfg = ec.quadrotor(T=T)
system = QuadrotorDynamics(cfg["h"])
x0, xd = ec.quadrotor_batch(0, I, T=T, total=I)
smp = GaussianSampling(cfg["sigma"][:12], cfg["sigma"][12:], 1000, power=cfg["power"], seed=77)
bat = BatchedIrsLqrZeroOrder(system, cfg["Q"], cfg["Qd"], cfg["R"], x0, xd, cfg["u_trj_initial"], smp)


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


print("batch iteration        %.3f ms" % timed(bat.local_descent))
At, Bt, ct, st = bat.linearize()
At, Bt, ct = At.clone(), Bt.clone(), ct.clone()
print("linearize (smooth+fit) %.3f ms" % timed(bat.linearize))
print("riccati                %.3f ms" % timed(lambda: riccati_device(At, Bt, ct, bat._dQ, bat._dQd, bat._dR, bat._dxd, bat._xd_stride)))
bat.check()
