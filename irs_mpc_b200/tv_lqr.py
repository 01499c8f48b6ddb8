"""solve_tvlqr / get_solver with the reference's signatures (irs_lqr/tv_lqr.py:11-145).

The reference builds a Drake MathematicalProgram QP and calls OSQP.  Here the same optimisation
problem is solved by an affine Riccati recursion on the GPU (csrc/tvlqr.cuh) — exact, not
iterative — which is the QP's minimiser whenever no box bound is active.  Active bounds and the
position-controlled / relative-bound variants are not implemented.
"""
import numpy as np
import torch

from . import _device, _lib

_SOLVER_NAMES = ("osqp", "snopt", "clp", "scs", "gurobi")
TVLQR_FAILED = "TV_LQR failed. Optimization problem is not solved."


class RiccatiSolver:
    """Stand-in for the Drake solver objects returned by the reference's get_solver."""

    def __init__(self, name):
        self.name = name

    def __repr__(self):
        return "RiccatiSolver(requested=%r)" % self.name


def get_solver(solver_name):
    # tv_lqr.py:11-27: unknown names raise ValueError("Do not recognize solver.")
    if solver_name in _SOLVER_NAMES:
        return RiccatiSolver(solver_name)
    raise ValueError("Do not recognize solver.")


def riccati_device(At, Bt, ct, Q, Qd, R, xd, xd_stride):
    """Device-level backward pass.  At [I,T,n,n] ... -> K [I,T,m,n], k [I,T,m], status [I]."""
    I, T, n, _ = At.shape
    m = Bt.shape[3]
    K = _device.empty((I, T, m, n))
    k = _device.empty((I, T, m))
    status = _device.empty((I,), torch.int32)
    _lib.call("irs_tvlqr_riccati", n, m, _device.ptr(At), _device.ptr(Bt), _device.ptr(ct),
              _device.ptr(Q), _device.ptr(Qd), _device.ptr(R), _device.ptr(xd), int(xd_stride),
              I, T, _device.ptr(K), _device.ptr(k), _device.ptr(status), _device.stream_ptr())
    return K, k, status


def _violates(v, lo, hi, tol=1e-9):
    return bool(np.any(v < np.asarray(lo) - tol) or np.any(v > np.asarray(hi) + tol))


def solve_tvlqr(At, Bt, ct, Q, Qd, R, x0, x_trj_d, solver=None, indices_u_into_x=None,
                x_bound_abs=None, u_bound_abs=None, x_bound_rel=None, u_bound_rel=None,
                xinit=None, uinit=None):
    """Same arguments and return value as the reference (tv_lqr.py:30-33, :142-145):
    numpy float64 in, (x*[T+1,n], u*[T,m]) out.  ValueError(TVLQR_FAILED) on failure (:139-140)."""
    if indices_u_into_x is not None or x_bound_rel is not None or u_bound_rel is not None:
        raise NotImplementedError(
            "position-controlled / relative-bound TVLQR is outside the analytic-dynamics hot path")
    At = np.asarray(At, dtype=np.float64)
    Bt = np.asarray(Bt, dtype=np.float64)
    T, n, m = At.shape[0], At.shape[1], Bt.shape[2]
    ct = np.asarray(ct, dtype=np.float64).reshape(T, n)
    dA = _device.to_device(At[None])
    dB = _device.to_device(Bt[None])
    dc = _device.to_device(ct[None])
    dQ, dQd, dR = _device.to_device(Q), _device.to_device(Qd), _device.to_device(R)
    dxd = _device.to_device(np.asarray(x_trj_d, dtype=np.float64)[:T + 1])
    dx0 = _device.to_device(np.asarray(x0, dtype=np.float64)[None])
    K, k, status = riccati_device(dA, dB, dc, dQ, dQd, dR, dxd, 0)
    xs = _device.empty((1, T + 1, n))
    us = _device.empty((1, T, m))
    _lib.call("irs_tvlqr_linear_rollout", n, m, _device.ptr(dA), _device.ptr(dB), _device.ptr(dc),
              _device.ptr(K), _device.ptr(k), _device.ptr(dx0), 1, T, _device.ptr(xs),
              _device.ptr(us), _device.stream_ptr())
    if int(status.item()) != 0:
        raise ValueError(TVLQR_FAILED)
    xs = _device.to_numpy(xs[0])
    us = _device.to_numpy(us[0])
    if not (np.all(np.isfinite(xs)) and np.all(np.isfinite(us))):
        raise ValueError(TVLQR_FAILED)
    if x_bound_abs is not None and _violates(xs[1:], np.asarray(x_bound_abs[0])[1:T + 1],
                                             np.asarray(x_bound_abs[1])[1:T + 1]):
        raise NotImplementedError("an absolute state bound is active: box-constrained TVLQR is not "
                                  "implemented (inactive-bound regime only)")
    if u_bound_abs is not None and _violates(us, np.asarray(u_bound_abs[0])[:T],
                                             np.asarray(u_bound_abs[1])[:T]):
        raise NotImplementedError("an absolute input bound is active: box-constrained TVLQR is not "
                                  "implemented (inactive-bound regime only)")
    return xs, us
