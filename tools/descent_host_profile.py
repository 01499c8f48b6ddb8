"""Host-side anatomy of one local_descent at BASELINE.json configs[2] (tuning aid): where the wall time of
the synchronous call goes — staging, graph key, node update, launch, device time, unpacking."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from irs_mpc_b200 import _device, _lib, example_configs as ec, irs_lqr as mod, smoothing     # noqa: E402
from irs_mpc_b200.all import (GaussianSampling, IrsLqrParameters, IrsLqrZeroOrder,       # noqa: E402
                              QuadrotorDynamics)

mod._PIPELINE_SEGMENTS = int(os.environ.get("SEGS", "5"))
cfg = ec.quadrotor(T=100)
system = QuadrotorDynamics(cfg["h"])
params = IrsLqrParameters()
for key in ("Q", "Qd", "R", "x0", "xd_trj", "u_trj_initial", "xbound", "ubound"):
    setattr(params, key, cfg[key])
sampler = GaussianSampling(cfg["sigma"][:12], cfg["sigma"][12:], 100000, power=0.5, seed=7)
solver = IrsLqrZeroOrder(system, params, sampler)
x, u = solver.x_trj, solver.u_trj
for _ in range(6):
    solver.local_descent(x, u)
T, n, m = solver.T, 12, 4
db = solver._descent_buffers()
nx, nu = db["nx"], db["nu"]
slot = solver._graphs._slots["descent"]
acc = {k: 0.0 for k in ("stage", "key", "update", "launch", "sync", "unpack", "total")}
K = 300
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
dev = 0.0
for it in range(K):
    t0 = time.perf_counter()
    h = db["in_host"].numpy()
    h[:nx] = np.asarray(x, dtype=np.float64)[:T + 1].reshape(-1)
    h[nx:] = np.asarray(u, dtype=np.float64)[:T].reshape(-1)
    t1 = time.perf_counter()
    key = solver._graph_key()
    t2 = time.perf_counter()
    solver._graph_update(slot[1])
    t3 = time.perf_counter()
    e0.record()
    _lib.call("irs_graph_launch", slot[1], _device.stream_ptr())
    e1.record()
    t4 = time.perf_counter()
    torch.cuda.current_stream().synchronize()
    t5 = time.perf_counter()
    o = db["out_host"].numpy()
    smoothing.check_status(o[nx + nu + 3:].view(np.int32)[:T])
    x_out = o[:nx].reshape(T + 1, n).copy()
    u_out = o[nx:nx + nu].reshape(T, m).copy()
    ok = np.all(np.isfinite(x_out)) and np.all(np.isfinite(u_out))
    t6 = time.perf_counter()
    dev += e0.elapsed_time(e1)
    for k_, v in zip(("stage", "key", "update", "launch", "sync", "unpack", "total"),
                     (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4, t6 - t5, t6 - t0)):
        acc[k_] += v
print("segments", solver._pipeline_segments())
print({k_: round(1e6 * v / K, 1) for k_, v in acc.items()}, "device(graph) us:", round(1e3 * dev / K, 1))
t0 = time.perf_counter()
for it in range(K):
    xn, un = solver.local_descent(x, u)
    c = solver.evaluate_cost(xn, un)
print("public API: %.1f us per descent + cost" % (1e6 * (time.perf_counter() - t0) / K))
