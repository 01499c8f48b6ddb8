"""The four analytic systems of the reference behind their original constructor signatures.

PendulumDynamics(h)   examples/pendulum/pendulum_dynamics.py:8-19
BicycleDynamics(h)    examples/bicycle/bicycle_dynamics.py:8-17
QuadrotorDynamics(h)  examples/quadrotor/quadrotor_dynamics.py:15-38 (attributes m, L, g, I, kF, kM)
ThreeCartDynamics(dt) examples/three_cart/three_cart_dynamics.py:8-18 (attribute d, `projection`)

All arithmetic runs in the CUDA functors of csrc/systems.cuh.
"""
import numpy as np
import torch

from . import _device, _lib
from .dynamical_system import CudaDynamicalSystem


class PendulumDynamics(CudaDynamicalSystem):
    system_id = 0
    system_name = "pendulum"

    def __init__(self, h):
        super().__init__()
        self.h = h
        self.dim_x = 2
        self.dim_u = 1

    def device_params(self):
        return [self.h]


class BicycleDynamics(CudaDynamicalSystem):
    system_id = 1
    system_name = "bicycle"

    def __init__(self, h):
        super().__init__()
        self.h = h
        self.dim_x = 5
        self.dim_u = 2

    def device_params(self):
        return [self.h]


class QuadrotorDynamics(CudaDynamicalSystem):
    system_id = 2
    system_name = "quadrotor"

    def __init__(self, h):
        super().__init__()
        self.h = h
        self.dim_x = 12
        self.dim_u = 4
        self.m = 0.775
        self.L = 0.15
        self.g = 9.81
        self.I = np.diag([0.0015, 0.0025, 0.0035])
        self.I_inv = np.linalg.inv(self.I)
        self.kF = 1.0
        self.kM = 0.0245

    def device_params(self):
        I = np.asarray(self.I)
        if np.any(np.abs(I - np.diag(np.diag(I))) > 0):
            raise ValueError("the CUDA quadrotor functor supports a diagonal inertia matrix only")
        return [self.h, self.m, self.L, self.g, I[0, 0], I[1, 1], I[2, 2], self.kF, self.kM]


class ThreeCartDynamics(CudaDynamicalSystem):
    system_id = 3
    system_name = "three_cart"
    batch_differs_from_scalar = True
    centered_capable = True          # absolute-point regressors can be accumulated relative to the nominal

    def __init__(self, dt):
        super().__init__()
        self.h = dt
        self.dim_x = 6
        self.dim_u = 2
        self.d = 0.2

    def device_params(self):
        return [self.h, self.d]

    def jacobian_xu(self, x, u):
        raise NotImplementedError(
            "Non differentiable simulation, does not support Jacobian computations.")

    def jacobian_xu_batch(self, x, u):
        raise NotImplementedError(
            "Non differentiable simulation, does not support Jacobian computations.")

    def projection(self, x, dx, u, du):
        """three_cart_dynamics.py:196-264: returns ABSOLUTE (x+dx projected, u+du)."""
        from .sampling import project_samples
        return project_samples(self, x, dx, u, du)


SYSTEM_CLASSES = {
    "pendulum": PendulumDynamics,
    "bicycle": BicycleDynamics,
    "quadrotor": QuadrotorDynamics,
    "three_cart": ThreeCartDynamics,
}
