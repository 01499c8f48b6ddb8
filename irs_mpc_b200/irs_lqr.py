"""IrsLqrParameters / IrsLqr / IrsLqrExact / IrsLqrFirstOrder / IrsLqrZeroOrder.

Same names, constructor signatures, public methods and public state as the reference
(irs_lqr/irs_lqr.py:7-218, irs_lqr_exact.py, irs_lqr_first_order.py, irs_lqr_zero_order.py); the
arithmetic runs in the sm_100a kernels of libirs_mpc_b200.so.  numpy float64 at the API edge.

Differences a user can observe, all deliberate:
  * `local_descent` performs ONE Riccati backward pass + closed-loop rollout instead of T QP
    solves; identical results while the box bounds are inactive.  If the new
    trajectory touches `xbound`/`ubound`, NotImplementedError is raised (box-constrained TVLQR is
    the next row of SURVEY.md section 8f) — there is no silent clamping.
  * `sampling` may be a `GaussianSampling` object: the noise is then generated inside the kernel.
    Any other callable is honoured through the replay path.
  * only systems derived from `CudaDynamicalSystem` are accepted (no CPU fallback).
"""
import time

import numpy as np
import torch

from . import _device, _lib, smoothing
from .dynamical_system import CudaDynamicalSystem
from .sampling import GaussianSampling
from .tv_lqr import TVLQR_FAILED, get_solver, riccati_device


class IrsLqrParameters:
    """Attribute bag, irs_lqr/irs_lqr.py:7-31."""

    def __init__(self):
        self.Q = None
        self.Qd = None
        self.R = None
        self.x0 = None
        self.xd_trj = None
        self.u_trj_initial = None
        self.xbound = None
        self.ubound = None
        self.solver_name = "osqp"


class IrsLqr:
    def __init__(self, system, params):
        self.system = system
        self.params = params
        self.check_valid_system(self.system)
        self.check_valid_params(self.params, self.system)

        self.Q = params.Q
        self.Qd = params.Qd
        self.R = params.R
        self.x0 = params.x0
        self.xd_trj = params.xd_trj
        self.u_trj = params.u_trj_initial
        self.xbound = params.xbound
        self.ubound = params.ubound
        self.solver = get_solver(params.solver_name)

        self.T = self.u_trj.shape[0]
        self.dim_x = self.system.dim_x
        self.dim_u = self.system.dim_u

        # device-resident problem data (float64)
        self._dQ = _device.to_device(np.asarray(self.Q, dtype=np.float64))
        self._dQd = _device.to_device(np.asarray(self.Qd, dtype=np.float64))
        self._dR = _device.to_device(np.asarray(self.R, dtype=np.float64))
        self._dxd = _device.to_device(np.asarray(self.xd_trj, dtype=np.float64)[:self.T + 1])
        self._ws = None
        self.timings = {}

        self.x_trj = self.rollout(self.x0, self.u_trj)
        self.cost = self.evaluate_cost(self.x_trj, self.u_trj)

        self.x_trj_lst = [self.x_trj]
        self.u_trj_lst = [self.u_trj]
        self.cost_lst = [self.cost]

        self.start_time = time.time()
        self.iter = 1

    # -- validation (irs_lqr.py:73-103, same messages) ---------------------------------------
    def check_valid_system(self, system):
        if system.dim_x == 0:
            raise RuntimeError("System has zero states. Did you forget to set dim_x?")
        elif system.dim_u == 0:
            raise RuntimeError("System has zero inputs. Did you forget to set dim_u?")
        if not isinstance(system, CudaDynamicalSystem):
            raise RuntimeError(
                "Could not evaluate dynamics. irs_mpc_b200 runs the dynamics as CUDA functors: the "
                "system must derive from CudaDynamicalSystem (pendulum, bicycle, quadrotor, "
                "three_cart); there is no CPU fallback for arbitrary Python dynamics.")
        try:
            system.dynamics(np.zeros(system.dim_x), np.zeros(system.dim_u))
        except _lib.IrsCudaError:
            raise
        except Exception:
            raise RuntimeError("Could not evaluate dynamics. Have you implemented it?")

    def check_valid_params(self, params, system):
        if np.asarray(params.Q).shape != (system.dim_x, system.dim_x):
            raise RuntimeError("Q matrix must be diagonal with dim_x x dim_x.")
        if np.asarray(params.Qd).shape != (system.dim_x, system.dim_x):
            raise RuntimeError("Qd matrix must be diagonal with dim_x x dim_x.")
        if np.asarray(params.R).shape != (system.dim_u, system.dim_u):
            raise RuntimeError("R matrix must be diagonal with dim_u x dim_u.")

    # -- rollout / cost (irs_lqr.py:105-137) ---------------------------------------------------
    def rollout(self, x0, u_trj):
        u = _device.to_device(np.asarray(u_trj, dtype=np.float64).reshape(1, self.T, self.dim_u))
        x0d = _device.to_device(np.asarray(x0, dtype=np.float64).reshape(1, self.dim_x))
        x_trj = _device.empty((1, self.T + 1, self.dim_x))
        cost = _device.empty((1,))
        prm, nprm = self.system._params()
        _lib.call("irs_rollout_open_loop", self.system.system_id, prm, nprm, _device.ptr(u),
                  _device.ptr(x0d), _device.ptr(self._dxd), 0, _device.ptr(self._dQ),
                  _device.ptr(self._dR), 1, self.T, _device.ptr(x_trj), _device.ptr(cost),
                  _device.stream_ptr())
        return _device.to_numpy(x_trj[0])

    def evaluate_cost(self, x_trj, u_trj):
        x = _device.to_device(np.asarray(x_trj, dtype=np.float64).reshape(1, self.T + 1, self.dim_x))
        u = _device.to_device(np.asarray(u_trj, dtype=np.float64).reshape(1, self.T, self.dim_u))
        cost = _device.empty((1,))
        _lib.call("irs_evaluate_cost", self.dim_x, self.dim_u, _device.ptr(x), _device.ptr(u),
                  _device.ptr(self._dxd), 0, _device.ptr(self._dQ), _device.ptr(self._dR), 1, self.T,
                  _device.ptr(cost), _device.stream_ptr())
        return float(cost.item())

    # -- linearization -------------------------------------------------------------------------
    def _tv_matrices_device(self, x_nom, u_nom):
        """x_nom [T,n], u_nom [T,m] CUDA float64 -> (At, Bt, ct, status) CUDA tensors."""
        raise NotImplementedError("This class is virtual.")

    def get_TV_matrices(self, x_trj, u_trj):
        """(At[T,n,n], Bt[T,n,m], ct[T,n]) numpy float64, as in the reference."""
        x_nom = _device.to_device(np.asarray(x_trj, dtype=np.float64)[:self.T])
        u_nom = _device.to_device(np.asarray(u_trj, dtype=np.float64)[:self.T])
        At, Bt, ct, status = self._tv_matrices_device(x_nom, u_nom)
        smoothing.check_status(status)
        return _device.to_numpy(At), _device.to_numpy(Bt), _device.to_numpy(ct)

    # -- descent (irs_lqr.py:148-186) ------------------------------------------------------------
    def local_descent(self, x_trj, u_trj):
        T, n, m = self.T, self.dim_x, self.dim_u
        xh = np.asarray(x_trj, dtype=np.float64)
        x_all = _device.to_device(xh)
        u_nom = _device.to_device(np.asarray(u_trj, dtype=np.float64)[:T])
        x_nom = x_all[:T]
        At, Bt, ct, status = self._tv_matrices_device(x_nom, u_nom)
        K, k, rstatus = riccati_device(At.view(1, T, n, n), Bt.view(1, T, n, m), ct.view(1, T, n),
                                       self._dQ, self._dQd, self._dR, self._dxd, 0)
        x_new = _device.empty((1, T + 1, n))
        u_new = _device.empty((1, T, m))
        cost = _device.empty((1,))
        prm, nprm = self.system._params()
        _lib.call("irs_rollout_closed_loop", self.system.system_id, prm, nprm, _device.ptr(K),
                  _device.ptr(k), _device.ptr(x_all[0:1]), _device.ptr(self._dxd), 0,
                  _device.ptr(self._dQ), _device.ptr(self._dR), 1, T, _device.ptr(x_new),
                  _device.ptr(u_new), _device.ptr(cost), _device.stream_ptr())
        # one synchronising read-back for everything the host needs
        smoothing.check_status(status)
        if int(rstatus.item()) != 0:
            raise ValueError(TVLQR_FAILED)
        x_out = _device.to_numpy(x_new[0])
        u_out = _device.to_numpy(u_new[0])
        if not (np.all(np.isfinite(x_out)) and np.all(np.isfinite(u_out))):
            raise ValueError(TVLQR_FAILED)
        self._check_bounds(x_out, u_out)
        self._last_descent_cost = float(cost.item())
        return x_out, u_out

    def _check_bounds(self, x_new, u_new, tol=1e-9):
        if self.xbound is not None:
            lo, hi = np.asarray(self.xbound[0]), np.asarray(self.xbound[1])
            if np.any(x_new[1:] < lo - tol) or np.any(x_new[1:] > hi + tol):
                raise NotImplementedError(
                    "a state bound (params.xbound) is active on the new trajectory: the "
                    "box-constrained TVLQR of the reference (tv_lqr.py:113-124) is not implemented")
        if self.ubound is not None:
            lo, hi = np.asarray(self.ubound[0]), np.asarray(self.ubound[1])
            if np.any(u_new < lo - tol) or np.any(u_new > hi + tol):
                raise NotImplementedError(
                    "an input bound (params.ubound) is active on the new trajectory: the "
                    "box-constrained TVLQR of the reference (tv_lqr.py:113-124) is not implemented")

    # -- iteration (irs_lqr.py:188-218): runs max_iterations + 1 descents ------------------------
    def iterate(self, max_iterations, verbose=True):
        while True:
            x_trj_new, u_trj_new = self.local_descent(self.x_trj, self.u_trj)
            cost_new = self.evaluate_cost(x_trj_new, u_trj_new)

            if verbose:
                print("Iteration: {:02d} ".format(self.iter) + " || " +
                      "Current Cost: {0:05f} ".format(cost_new) + " || " +
                      "Elapsed time: {0:05f} ".format(time.time() - self.start_time))

            self.x_trj_lst.append(x_trj_new)
            self.u_trj_lst.append(u_trj_new)
            self.cost_lst.append(cost_new)

            if self.iter > max_iterations:
                break

            self.cost = cost_new
            self.x_trj = x_trj_new
            self.u_trj = u_trj_new
            self.iter += 1

        return self.x_trj, self.u_trj, self.cost


class IrsLqrExact(IrsLqr):
    """irs_lqr/irs_lqr_exact.py:6-31 — Jacobian at the nominal point, evaluated in fp64."""

    def __init__(self, system, params):
        super().__init__(system, params)

    def _tv_matrices_device(self, x_nom, u_nom):
        P, n, m = x_nom.shape[0], self.dim_x, self.dim_u
        At = _device.empty((P, n, n))
        Bt = _device.empty((P, n, m))
        ct = _device.empty((P, n))
        prm, nprm = self.system._params()
        _lib.call("irs_exact_linearize", self.system.system_id, prm, nprm, _device.ptr(x_nom),
                  _device.ptr(u_nom), P, _device.ptr(At), _device.ptr(Bt), _device.ptr(ct),
                  _device.stream_ptr())
        status = torch.zeros((P,), dtype=torch.int32, device=x_nom.device)
        return At, Bt, ct, status


class _SampledIrsLqr(IrsLqr):
    order = None

    def __init__(self, system, params, sampling):
        super().__init__(system, params)
        self.sampling = sampling

    def _replay_noise(self, x_nom, u_nom):
        """Call the user's closure once per timestep (as the reference does) and upload."""
        xh = _device.to_numpy(x_nom)
        uh = _device.to_numpy(u_nom)
        rows = []
        for t in range(xh.shape[0]):
            dx, du = self.sampling(xh[t], uh[t], self.iter)
            rows.append(np.hstack((np.asarray(dx), np.asarray(du))).astype(np.float32))
        return _device.to_device(np.stack(rows), torch.float32)

    def _tv_matrices_device(self, x_nom, u_nom):
        s = self.sampling
        if isinstance(s, GaussianSampling):
            At, Bt, ct, status, self._ws = smoothing.linearize(
                self.system, self.order, x_nom, u_nom, s.num_samples, self._ws,
                sigma=s.sigma(self.iter), seed=s.seed, it=self.iter, stream_id=s.stream_id,
                flags=s.flags())
        else:
            noise = self._replay_noise(x_nom, u_nom)
            At, Bt, ct, status, self._ws = smoothing.linearize(
                self.system, self.order, x_nom, u_nom, noise.shape[1], self._ws, noise=noise)
        return At, Bt, ct, status


class IrsLqrFirstOrder(_SampledIrsLqr):
    """irs_lqr/irs_lqr_first_order.py:6-54 — averaged Jacobians at the perturbed points."""
    order = smoothing.FIRST_ORDER


class IrsLqrZeroOrder(_SampledIrsLqr):
    """irs_lqr/irs_lqr_zero_order.py:5-63 — least-squares fit of sampled dynamics."""
    order = smoothing.ZERO_ORDER

    def compute_least_squares(self, dxdu, deltaf):
        """irs_lqr_zero_order.py:27-36: (Ahat, Bhat) = lstsq(dxdu, deltaf)^T split.  Runs the same
        Gram + Cholesky kernels as the fused path (normal equations in fp64 on fp32 partials)."""
        dxdu = np.asarray(dxdu, dtype=np.float64)
        deltaf = np.asarray(deltaf, dtype=np.float64)
        n, d = self.dim_x, self.dim_x + self.dim_u
        G = _device.to_device(dxdu)
        F = _device.to_device(deltaf)
        # small dense problem: Gram on device in fp64 through torch plumbing is NOT the hot path;
        # this helper exists for API completeness only.
        gram = (G.T @ G)
        rhs = (G.T @ F)
        L = torch.linalg.cholesky(gram)
        AB = torch.cholesky_solve(rhs, L).T
        AB = _device.to_numpy(AB)
        return AB[:, :n], AB[:, n:d]
