"""Learned dynamics (pendulum_nn.py: 3-100-100-2 ReLU network, T=200, N=1e4 samples per step): ms per
linearization with the hidden layer on the tensor cores and with the per-thread functor, ms per iRS-LQR iteration,
and the float32 numpy oracle on the host (tuning aid)."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from irs_mpc_b200 import _device, example_configs as ec, smoothing                           # noqa: E402
from irs_mpc_b200.all import GaussianSampling, IrsLqrParameters, IrsLqrZeroOrder, MlpDynamics   # noqa: E402

T = int(os.environ.get("MB_T", "200"))
N = int(os.environ.get("MB_N", "10000"))
g = np.load(os.path.join(ROOT, "tests", "golden", "mlp_pendulum.npz"))
system = MlpDynamics([(g["W1"], g["b1"]), (g["W2"], g["b2"]), (g["W3"], g["b3"])])
cfg = ec.pendulum_nn(T=T)
x = _device.to_device(np.cumsum(0.02 * np.ones((T, 2)), axis=0))
u = _device.to_device(cfg["u_trj_initial"])
flops = 2 * (3 * 100 + 100 * 100 + 100 * 2)


def timed(fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for engine in ("1", "0") if os.environ.get("MB_BOTH", "1") == "1" else ("1",):
    os.environ["IRS_MLP_ENGINE"] = engine
    ws = smoothing.Workspace(system, 0, T, N)
    seed = [0]

    def acc():
        seed[0] += 1
        smoothing.accumulate(system, 0, x, u, N, ws, sigma=cfg["sigma"], seed=seed[0], it=1, flags=8)

    def lin():
        acc()
        smoothing.finalize(system, 0, x, u, ws, N)
    ms_a, ms_l = timed(acc), timed(lin)
    print("engine %s (C=%d S=%d): accumulate %.4f ms (%.3e samples/s, %.1f TFLOP/s of network math), accumulate+fit %.4f ms"
          % ("tcgen05" if engine == "1" else "per-thread", ws.C, ws.S, ms_a, T * N / (ms_a * 1e-3),
             T * N * flops / (ms_a * 1e-3) / 1e12, ms_l), flush=True)
os.environ["IRS_MLP_ENGINE"] = "1"
params = IrsLqrParameters()
for key in ("Q", "Qd", "R", "x0", "xd_trj", "u_trj_initial", "xbound", "ubound"):
    setattr(params, key, cfg[key])
sampler = GaussianSampling(cfg["sigma"][:2], cfg["sigma"][2:], N, power=cfg["power"], seed=3)
solver = IrsLqrZeroOrder(system, params, sampler)


def iteration():
    xn, un = solver.local_descent(solver.x_trj, solver.u_trj)
    return solver.evaluate_cost(xn, un)
t0 = time.time()
for _ in range(5):
    iteration()
torch.cuda.synchronize()
t0 = time.time()
for _ in range(20):
    iteration()
torch.cuda.synchronize()
print("iRS-LQR iteration (local_descent + evaluate_cost, numpy in/out): %.3f ms" % ((time.time() - t0) / 20 * 1e3))
t0 = time.time()
for _ in range(20):
    solver.rollout(cfg["x0"], solver.u_trj)
print("open-loop rollout, T=%d: %.3f ms" % (T, (time.time() - t0) / 20 * 1e3))
t0 = time.time()
for _ in range(20):
    solver.get_TV_matrices(solver.x_trj, solver.u_trj)
print("get_TV_matrices (numpy in/out): %.3f ms" % ((time.time() - t0) / 20 * 1e3))
# host: the oracle's float32 numpy network, one timestep's N samples (what the reference's loop does T times)
sys.path.insert(0, ROOT)
from oracle.mlp_ref import MlpOracle      # noqa: E402
orc = MlpOracle([g[k] for k in ("W1", "b1", "W2", "b2", "W3", "b3")])
dx = np.random.default_rng(0).standard_normal((N, 3))
t0 = time.time()
reps = 20
for _ in range(reps):
    f = orc.dynamics_batch(dx[:, :2], dx[:, 2:])
    np.linalg.lstsq(dx, f - f[0], rcond=None)
dt = (time.time() - t0) / reps
print("host numpy (float32 network + lstsq), one timestep of N=%d: %.3f ms -> %.3e samples/s" % (N, dt * 1e3, N / dt))
