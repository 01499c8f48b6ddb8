"""Host cost of one ShardedLinearizer.linearize_n call against its device time (tuning aid; torchrun, >= 2 ranks)."""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from irs_mpc_b200 import _device, example_configs as ec      # noqa: E402
from irs_mpc_b200.all import QuadrotorDynamics               # noqa: E402
from irs_mpc_b200.distributed import ShardedLinearizer       # noqa: E402

rank = int(os.environ["RANK"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
T, N = 100, 100000
cfg = ec.quadrotor(T=T)
s = QuadrotorDynamics(cfg["h"])
x = _device.to_device(np.zeros((T, 12)))
u = _device.to_device(cfg["u_trj_initial"])
sh = ShardedLinearizer(s, 0)
kw = dict(sigma=cfg["sigma"], it=1, flags=8)
for k in range(10):
    sh.linearize_n(x, u, N, seed=k, **kw)
dist.barrier()
torch.cuda.synchronize()
for steps in (30, 300):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0.record()
    for k in range(steps):
        sh.linearize_n(x, u, N, seed=100 + k, **kw)
    e1.record()
    t_host = time.perf_counter() - t0
    torch.cuda.synchronize()
    if rank == 0:
        print("%d steps: host %.1f us per call (enqueue only), device %.1f us per step" % (
            steps, t_host / steps * 1e6, e0.elapsed_time(e1) / steps * 1e3), flush=True)
dist.destroy_process_group()
