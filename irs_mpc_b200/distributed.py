"""Multi-GPU smoothing: one process per GPU (torchrun), torch.distributed for the plumbing.

The reference's only "distributed backend" is a ZeroMQ PUSH/PULL task farm over timesteps
(irs_lqr/irs_lqr_quasistatic.py:228-273, zmq_parallel_cmp/array_io.py:6-26).  Here:

* timestep sharding (`linearize_t`): rank r owns the nominal points [r*ceil(T/W), ...); every
  rank runs the fused kernels on its slice with the GLOBAL point index in the Philox counter and
  the per-point blocks [A_t | B_t | c_t] are all-gathered (n*(n+m+1) doubles per step).  Results
  are bit-identical to the single-GPU run because the per-point reduction order is unchanged.
* sample sharding (`linearize_n`): every rank owns all T points and a slice of the samples
  (global sample index in the Philox counter); the chunk-reduced fp64 Gram blocks [T, width] are
  all-gathered and every rank sums them in rank order and solves — deterministic, identical on
  all ranks, and the exchange is T*width doubles per rank.
The sequential Riccati pass is never split (replicated on every rank).

The gather helpers are backend agnostic (NCCL on GPUs, gloo in the CPU tests).
"""
import torch
import torch.distributed as dist

from . import _device, smoothing


def shard_range(total, world, rank):
    """Contiguous ceil-split: [start, stop) of `total` items owned by `rank` (may be empty)."""
    per = (total + world - 1) // world
    start = min(rank * per, total)
    stop = min(start + per, total)
    return start, stop, per


def pack_abc(At, Bt, ct):
    """[L,n,n], [L,n,m], [L,n] -> [L, n*(n+m+1)] (row-major [A|B|c] per step)."""
    L, n = ct.shape
    m = Bt.shape[2]
    return torch.cat((At.reshape(L, n * n), Bt.reshape(L, n * m), ct.reshape(L, n)), dim=1)


def unpack_abc(packed, n, m):
    L = packed.shape[0]
    At = packed[:, :n * n].reshape(L, n, n)
    Bt = packed[:, n * n:n * n + n * m].reshape(L, n, m)
    ct = packed[:, n * n + n * m:].reshape(L, n)
    return At, Bt, ct


def gather_rows(local, total, group=None):
    """All-gather row blocks of a ceil-split [total, w] array.  `local` holds this rank's rows
    (possibly fewer than ceil(total/W), possibly zero); returns the full [total, w] tensor."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    start, stop, per = shard_range(total, world, rank)
    assert local.shape[0] == stop - start, (local.shape, start, stop)
    w = local.shape[1]
    padded = torch.zeros((per, w), dtype=local.dtype, device=local.device)
    if stop > start:
        padded[:stop - start].copy_(local)
    out = torch.empty((world * per, w), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, padded, group=group)
    return out[:total]


def gather_ranks(local, group=None):
    """All-gather equally shaped blocks: [*shape] -> [W, *shape]."""
    world = dist.get_world_size(group)
    flat = local.contiguous().view(-1)
    out = torch.empty((world * flat.numel(),), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, flat, group=group)
    return out.view((world,) + tuple(local.shape))


class ShardedLinearizer:
    def __init__(self, system, order, group=None):
        self.system, self.order, self.group = system, order, group
        self._ws = None

    def _workspace(self, P, N):
        key = (self.system.system_id, self.order, P, N)
        if self._ws is None or self._ws.key != key:
            self._ws = smoothing.Workspace(self.system, self.order, P, N)
        return self._ws

    def linearize_t(self, x_nom, u_nom, N, **kw):
        """Timestep-sharded.  x_nom [T,n], u_nom [T,m] replicated on every rank; returns the full
        (At, Bt, ct, status) on every rank."""
        T = x_nom.shape[0]
        n, m = self.system.dim_x, self.system.dim_u
        world, rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        start, stop, _ = shard_range(T, world, rank)
        width = n * (n + m + 1) + 1
        if stop > start:
            ws = self._workspace(stop - start, N)
            xs, us = x_nom[start:stop].contiguous(), u_nom[start:stop].contiguous()
            smoothing.accumulate(self.system, self.order, xs, us, N, ws, p0=start, **kw)
            At, Bt, ct, status = smoothing.finalize(self.system, self.order, xs, us, ws, N)
            local = torch.cat((pack_abc(At, Bt, ct), status.to(torch.float64).unsqueeze(1)), dim=1)
        else:
            local = torch.zeros((0, width), dtype=torch.float64, device=x_nom.device)
        full = gather_rows(local, T, self.group)
        At, Bt, ct = unpack_abc(full[:, :width - 1], n, m)
        return At, Bt, ct, full[:, width - 1].to(torch.int32)

    def linearize_n(self, x_nom, u_nom, N_local, **kw):
        """Sample-sharded.  Every rank draws N_local samples per point (global sample index
        rank*N_local + i); returns (At, Bt, ct, status) fitted on all W*N_local samples."""
        T = x_nom.shape[0]
        world, rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        ws = self._workspace(T, N_local)
        smoothing.accumulate(self.system, self.order, x_nom, u_nom, N_local, ws,
                             i0=rank * N_local, **kw)
        mine = smoothing.reduce_chunks(self.system, self.order, ws)
        everyone = gather_ranks(mine, self.group)
        return smoothing.finalize(self.system, self.order, x_nom, u_nom, ws, world * N_local,
                                  reduced=everyone, nranks=world, rank_stride=mine.numel())
