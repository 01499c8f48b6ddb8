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


class MlpDynamics(CudaDynamicalSystem):
    """Learned dynamics x+ = net([x, u]): the reference's PendulumNN wrapper
    (examples/pendulum/pendulum_nn.py:66-90) around its DynamicsNLP network (:19-33),
    Linear(dim_x + dim_u, H1) ReLU Linear(H1, H2) ReLU Linear(H2, dim_x), evaluated in float32 like the
    reference's torch module and returned as float64 numpy.  `jacobian_xu` is the exact derivative of the
    piecewise-linear network — what torch.autograd.grad gives in pendulum_nn.py:78-85.

    network: a torch.nn.Module holding exactly three nn.Linear layers with ReLU between them (the reference's
    `DynamicsNLP()` or its `.dynamics_mlp`), or a sequence [(W1, b1), (W2, b2), (W3, b3)] of arrays in
    torch.nn.Linear layout (W [out, in]).  A module is checked against this class's own forward pass on
    random inputs, so any other architecture is refused instead of silently mis-evaluated.  The weights are
    copied at construction (re-create the object after further training).

    Built for dim_x = 2, dim_u = 1 (the reference's learned pendulum) and hidden widths up to 128."""
    system_id = 4
    system_name = "mlp_2_1"

    def __init__(self, network, dim_x=2, dim_u=1, h=0.0):
        super().__init__()
        if (dim_x, dim_u) != (2, 1):
            raise RuntimeError("learned dynamics are built for dim_x = 2, dim_u = 1 (pendulum_nn.py:66-71)")
        self.h = h
        self.dim_x = dim_x
        self.dim_u = dim_u
        layers = self._layers_of(network)
        if len(layers) != 3:
            raise RuntimeError("expected three linear layers (pendulum_nn.py:23-29), got %d" % len(layers))
        d = dim_x + dim_u
        (W1, b1), (W2, b2), (W3, b3) = [(np.ascontiguousarray(W, dtype=np.float32), np.ascontiguousarray(b, dtype=np.float32))
                                        for W, b in layers]
        H1, H2 = W1.shape[0], W2.shape[0]
        if W1.shape != (H1, d) or W2.shape != (H2, H1) or W3.shape != (dim_x, H2) or b1.shape != (H1,) \
                or b2.shape != (H2,) or b3.shape != (dim_x,):
            raise RuntimeError("layer shapes do not chain to %d -> H1 -> H2 -> %d" % (d, dim_x))
        self.weights = (W1, b1, W2, b2, W3, b3)
        import ctypes
        handle = ctypes.c_int(-1)
        _lib.call("irs_mlp_register", dim_x, dim_u, H1, H2, *[a.ctypes.data for a in self.weights], ctypes.byref(handle))
        self._handle = handle.value
        if not isinstance(network, (list, tuple)):
            self._check_module(network)

    @staticmethod
    def _layers_of(network):
        if isinstance(network, (list, tuple)):
            return [(np.asarray(W), np.asarray(b)) for W, b in network]
        linear = [mod for mod in network.modules() if isinstance(mod, torch.nn.Linear)]
        return [(mod.weight.detach().cpu().numpy(), mod.bias.detach().cpu().numpy()) for mod in linear]

    def forward_numpy(self, xu):
        """float32 forward pass of the registered weights (host; used to validate a module, not on the data path)."""
        W1, b1, W2, b2, W3, b3 = self.weights
        a = np.maximum(np.asarray(xu, dtype=np.float32) @ W1.T + b1, 0)
        a = np.maximum(a @ W2.T + b2, 0)
        return a @ W3.T + b3

    def _check_module(self, network):
        rng = np.random.default_rng(0)
        xu = (4.0 * rng.standard_normal((64, self.dim_x + self.dim_u))).astype(np.float32)
        with torch.no_grad():
            want = network(torch.from_numpy(xu).to(next(network.parameters()).device)).cpu().numpy()
        got = self.forward_numpy(xu)
        if want.shape != got.shape or not np.allclose(want, got, rtol=1e-4, atol=1e-4 * max(1.0, float(np.abs(want).max()))):
            raise RuntimeError("the module is not Linear-ReLU-Linear-ReLU-Linear (its output differs from that "
                               "network's forward pass); only that architecture is supported")

    def device_params(self):
        return [self.h, float(self._handle)]

    def __del__(self):
        try:
            if getattr(self, "_handle", -1) >= 0:
                _lib.call("irs_mlp_release", self._handle)
        except Exception:      # interpreter shutdown: the library may be gone already
            pass


SYSTEM_CLASSES = {
    "pendulum": PendulumDynamics,
    "bicycle": BicycleDynamics,
    "quadrotor": QuadrotorDynamics,
    "three_cart": ThreeCartDynamics,
}
