"""Kernel-only timing of the smoothing accumulate stage (tuning aid, not the official bench).

    python tools/quick_bench.py [system ...]     e.g. quadrotor three_cart
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from irs_mpc_b200 import _device, example_configs as ec, smoothing   # noqa: E402
from irs_mpc_b200.systems import SYSTEM_CLASSES                      # noqa: E402

FLOPS = {"pendulum": 37, "bicycle": 158, "quadrotor": 838, "three_cart": 220}
SHAPES = {"pendulum": (200, 100000), "bicycle": (100, 100000), "quadrotor": (100, 100000),
          "three_cart": (100, 1000000)}


def run(name, order=0, reps=10):
    T, N = SHAPES[name]
    N = int(os.environ.get("QB_N", N))
    cfg = ec.CONFIGS[name](T=T)
    system = SYSTEM_CLASSES[name](cfg["h"])
    x = np.zeros((T, system.dim_x)) + cfg["x0"]
    x_nom = _device.to_device(x)
    u_nom = _device.to_device(cfg["u_trj_initial"])
    ws = smoothing.Workspace(system, order, T, N)
    # QB_FLAGS: extra smoothing flags (8 = antithetic pairs, the GaussianSampling default)
    flags = (2 if cfg["projection"] else 0) | int(os.environ.get("QB_FLAGS", "8"))
    times = []
    inner = 10          # launches per timed batch: host launch latency overlaps with the previous kernel
    for k in range(reps + 2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for j in range(inner):
            smoothing.accumulate(system, order, x_nom, u_nom, N, ws, sigma=cfg["sigma"], seed=k * inner + j, it=1,
                                 flags=flags)
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1) / inner)
    ms = float(np.median(times[2:]))
    sps = T * N / (ms * 1e-3)
    fl = FLOPS[name] if order == 0 else 61
    print("%-10s order=%d T=%d N=%d C=%d S=%d  %.3f ms  %.3e samples/s  %.2f TFLOP/s algorithmic (%.1f%% of 74.4)"
          % (name, order, T, N, ws.C, ws.S, ms, sps, sps * fl / 1e12, 100 * sps * fl / 74.4e12))


if __name__ == "__main__":
    names = sys.argv[1:] or ["quadrotor", "three_cart", "bicycle", "pendulum"]
    for nm in names:
        run(nm)
    if "bicycle" in names:
        run("bicycle", order=1)
