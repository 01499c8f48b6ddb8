"""A tiny pass over the smoothing kernels for compute-sanitizer (racecheck / memcheck): quadrotor and
three_cart (in-kernel projection, centred) through the paired tensor-core kernel, two chunks per point,
ragged tail, then the finalize; checked against nothing here (the parity tests do that)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from irs_mpc_b200 import _device, example_configs as ec, smoothing   # noqa: E402
from irs_mpc_b200.systems import SYSTEM_CLASSES                      # noqa: E402

for name, flags in (("quadrotor", 8), ("three_cart", 2 | 8), ("quadrotor", 0)):
    T, N = 3, 4097 + 300
    cfg = ec.CONFIGS[name](T=T)
    system = SYSTEM_CLASSES[name](cfg["h"])
    x_nom = _device.to_device(np.zeros((T, system.dim_x)) + cfg["x0"])
    u_nom = _device.to_device(cfg["u_trj_initial"])
    At, Bt, ct, status, ws = smoothing.linearize(system, 0, x_nom, u_nom, N, sigma=cfg["sigma"], seed=3, it=1, flags=flags)
    print(name, flags, "status", int(status.sum().item()), "A00", float(At[0, 0, 0].item()), flush=True)
