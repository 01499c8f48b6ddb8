"""Host driver of the randomized-smoothing linearization kernels.

`linearize(...)` is the device-level entry: nominal points already on the GPU in, (At, Bt, ct) on
the GPU out, no host synchronisation.  It implements the numeric core of
IrsLqrZeroOrder.get_TV_matrices (irs_lqr/irs_lqr_zero_order.py:38-63) and
IrsLqrFirstOrder.get_TV_matrices (irs_lqr/irs_lqr_first_order.py:28-54) for all nominal points at
once; the per-timestep Python loop of the reference becomes the grid's y-extent.
"""
import ctypes

import numpy as np
import torch

from . import _device, _lib

ZERO_ORDER = 0
FIRST_ORDER = 1


def plan(system_id, order, P, N):
    C = ctypes.c_int(0)
    S = ctypes.c_longlong(0)
    _lib.call("irs_smooth_plan", system_id, order, P, N, ctypes.byref(C), ctypes.byref(S))
    return C.value, S.value


class Workspace:
    """Reusable device buffers for one (system, order, P, N) shape."""

    def __init__(self, system, order, P, N):
        self.key = (system.system_id, order, P, N)
        self.C, self.S = plan(system.system_id, order, P, N)
        self.width = _lib.lib().irs_partial_width(system.system_id, order)
        n, m = system.dim_x, system.dim_u
        self.partials = _device.empty((P, self.C, self.width), torch.float32)
        self.At = _device.empty((P, n, n))
        self.Bt = _device.empty((P, n, m))
        self.ct = _device.empty((P, n))
        self.status = _device.empty((P,), torch.int32)


def accumulate(system, order, x_nom, u_nom, N, ws, sigma=None, noise=None, seed=0, it=1,
               stream_id=0, p0=0, i0=0, flags=0):
    """Launch the accumulation kernel: partial Gram blocks (order 0) or Jacobian sums (order 1)."""
    P = x_nom.shape[0]
    prm, nprm = system._params()
    if system.batch_differs_from_scalar and order == ZERO_ORDER:
        flags |= 1   # IRS_SAMPLES_BATCH_VARIANT: samples go through dynamics_batch (…zero_order.py:51)
    sig = None
    if noise is None:
        sig = np.ascontiguousarray(np.asarray(sigma, dtype=np.float32))
        if sig.shape != (system.dim_x + system.dim_u,):
            raise ValueError("sigma must have n + m = %d entries" % (system.dim_x + system.dim_u))
        sig = sig.ctypes.data_as(ctypes.c_void_p)
    fn = ("irs_smooth_zero_order_accumulate" if order == ZERO_ORDER
          else "irs_smooth_first_order_accumulate")
    _lib.call(fn, system.system_id, prm, nprm, flags, _device.ptr(x_nom), _device.ptr(u_nom), P,
              int(N), sig, _device.ptr(noise), int(seed), int(it), int(stream_id),
              int(p0), int(i0), ws.C, ws.S, _device.ptr(ws.partials), _device.stream_ptr())


def reduce_chunks(system, order, ws, out=None):
    """[P, C, width] fp32 partials -> [P, width] fp64 (the block exchanged between ranks)."""
    P = ws.partials.shape[0]
    if out is None:
        out = _device.empty((P, ws.width))
    _lib.call("irs_smooth_reduce_chunks", system.system_id, order, _device.ptr(ws.partials), P, ws.C,
              _device.ptr(out), _device.stream_ptr())
    return out


def finalize(system, order, x_nom, u_nom, ws, n_total, partials=None, reduced=None, nranks=1,
             rank_stride=0):
    """Fit from fp32 per-chunk partials (default: ws.partials) or from fp64 reduced blocks."""
    P = x_nom.shape[0]
    prm, nprm = system._params()
    part = None if reduced is not None else (ws.partials if partials is None else partials)
    _lib.call("irs_smooth_finalize", system.system_id, prm, nprm, order, _device.ptr(x_nom),
              _device.ptr(u_nom), P, ws.C, _device.ptr(part), _device.ptr(reduced), nranks,
              int(rank_stride), float(n_total), _device.ptr(ws.At), _device.ptr(ws.Bt),
              _device.ptr(ws.ct), _device.ptr(ws.status), _device.stream_ptr())
    return ws.At, ws.Bt, ws.ct, ws.status


def linearize(system, order, x_nom, u_nom, N, ws=None, **kw):
    """x_nom [P,n], u_nom [P,m] (CUDA float64) -> (At, Bt, ct, status) CUDA tensors."""
    P = x_nom.shape[0]
    if ws is None or ws.key != (system.system_id, order, P, N):
        ws = Workspace(system, order, P, N)
    accumulate(system, order, x_nom, u_nom, N, ws, **kw)
    return finalize(system, order, x_nom, u_nom, ws, N) + (ws,)


def check_status(status):
    bad = int(status.sum().item())
    if bad:
        raise np.linalg.LinAlgError(
            "smoothing fit: the sample Gram matrix [dx du]^T[dx du] is rank deficient at %d "
            "nominal point(s) (too few samples, or NaN in the dynamics)" % bad)
