"""Does the fit of batch k beside the sampling of batch k + 1 pay?  (tuning aid; plain python for one GPU,
torchrun for several.)  Two workspaces alternate; the fit runs on a second stream."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from irs_mpc_b200 import _device, example_configs as ec, smoothing      # noqa: E402
from irs_mpc_b200.all import QuadrotorDynamics                          # noqa: E402


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
    T, N = 100, 100000
    cfg = ec.quadrotor(T=T)
    s = QuadrotorDynamics(cfg["h"])
    x = _device.to_device(np.zeros((T, 12)))
    u = _device.to_device(cfg["u_trj_initial"])
    kw = dict(sigma=cfg["sigma"], it=1, flags=8)
    D = int(os.environ.get("PP_DEPTH", "3"))
    ws = [smoothing.Workspace(s, 0, T, N) for _ in range(D)]
    px = None
    if world > 1:
        from irs_mpc_b200.distributed import PeerExchange
        px = PeerExchange(T, ws[0].width)
    main_s = torch.cuda.current_stream()
    side = torch.cuda.Stream(priority=-1)
    ev_acc = [torch.cuda.Event() for _ in range(D)]
    ev_fin = [torch.cuda.Event() for _ in range(D)]

    def fit(w):
        if px is None:
            smoothing.finalize(s, 0, x, u, w, N)
        else:
            px.finalize(s, 0, x, u, w, world * N)

    def serial(k):
        smoothing.accumulate(s, 0, x, u, N, ws[0], seed=k, i0=rank * N, **kw)
        fit(ws[0])

    state = {"k": 0}

    def piped(k):
        b = state["k"] % D
        if state["k"] >= D:
            main_s.wait_event(ev_fin[b])
        smoothing.accumulate(s, 0, x, u, N, ws[b], seed=k, i0=rank * N, **kw)
        ev_acc[b].record(main_s)
        with torch.cuda.stream(side):
            side.wait_event(ev_acc[b])
            fit(ws[b])
            ev_fin[b].record(side)
        state["k"] += 1

    def timed(fn, steps=60, warm=8, drain=False):
        for k in range(warm):
            fn(k)
        if drain:
            main_s.wait_stream(side)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for k in range(steps):
            fn(warm + k)
        if drain:
            main_s.wait_stream(side)
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    a = timed(serial)
    b = timed(piped, drain=True)
    ok = int(torch.stack([w.status.max() for w in ws]).max().item())
    if rank == 0:
        print("depth %d, world %d: serial %.4f ms/step, fit beside the next batch's sampling %.4f ms/step (status %d)" % (D, world, a, b, ok),
              flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
