"""TEST INFRASTRUCTURE ONLY — not part of the product path.

Imports the UNMODIFIED reference (hjsuh94/irs_mpc at /root/reference) in THIS container so that
`oracle/make_golden.py` can generate golden vectors and `tests/test_oracle_vs_reference.py`
can pin the float64 restatement (`oracle/cpu_restatement.py`) against the reference's own code.

The reference imports `pydrake` at module scope (irs_lqr/tv_lqr.py:2-8,
examples/pendulum/pendulum_dynamics.py:2, examples/quadrotor/quadrotor_dynamics.py:6-11) and
builds symbolic expressions inside constructors (pendulum_dynamics.py:21-26,
bicycle_dynamics.py:20-24).  pydrake is not installed and cannot be (no network), so we register
stub modules whose attributes swallow attribute access, calls and arithmetic.  Everything that is
*numerically* evaluated by the reference on the smoothing path (dynamics, dynamics_batch,
projection, rollout, evaluate_cost, IrsLqrZeroOrder.get_TV_matrices, compute_least_squares) is
plain numpy and runs unmodified.  `jacobian_xu*` and `solve_tvlqr` need the real pydrake and do
NOT run; they are restated in oracle/cpu_restatement.py and pinned on the reference's stored
result files instead.

/root/reference does not exist on the GPU box: nothing under `-m gpu`, smoke() or bench.py
imports this module.
"""
import os
import sys
import types
import warnings

import numpy as np

REFERENCE_ROOT = os.environ.get("IRS_MPC_REFERENCE", "/root/reference")


class _Absorb:
    """Object that swallows calls, attribute access and arithmetic (pydrake stand-in)."""

    def __call__(self, *a, **k):
        return _Absorb()

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return _Absorb()

    def _binop(self, other):
        return _Absorb()

    __add__ = __radd__ = __sub__ = __rsub__ = _binop
    __mul__ = __rmul__ = __truediv__ = __rtruediv__ = __pow__ = __rpow__ = _binop

    def __neg__(self):
        return _Absorb()

    def __iter__(self):
        return iter(())


def _stub_module(name):
    mod = types.ModuleType(name)
    mod.__file__ = "<pydrake-stub:%s>" % name
    mod.__path__ = []

    def _getattr(attr):
        if attr.startswith("__") and attr.endswith("__"):
            raise AttributeError(attr)
        return _Absorb()

    mod.__getattr__ = _getattr
    return mod


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "irs_lqr"))


_installed = False


def install():
    """Make `import irs_lqr...` and the example dynamics modules resolve to the reference."""
    global _installed
    if _installed:
        return
    if not reference_available():
        raise RuntimeError("reference tree not found at %s" % REFERENCE_ROOT)
    import torch  # noqa: F401  (must be imported before the stubs shadow anything)

    for name in ("pydrake", "pydrake.all", "pydrake.symbolic", "pydrake.examples",
                 "pydrake.examples.quadrotor", "pydrake.forwarddiff"):
        if name not in sys.modules:
            sys.modules[name] = _stub_module(name)
    # numpy>=1.24 removed these aliases (quadrotor_dynamics.py:41,94 use them).
    if not hasattr(np, "float"):
        np.float = float
    if not hasattr(np, "object"):
        np.object = object
    warnings.filterwarnings("ignore", category=SyntaxWarning)
    paths = [REFERENCE_ROOT] + [os.path.join(REFERENCE_ROOT, "examples", s)
                                for s in ("pendulum", "bicycle", "quadrotor", "three_cart")]
    for p in reversed(paths):
        if p not in sys.path:
            sys.path.insert(0, p)
    # Our own drop-in shim package is also called `irs_lqr`; make sure the reference wins here.
    for k in [k for k in sys.modules if k == "irs_lqr" or k.startswith("irs_lqr.")]:
        del sys.modules[k]
    _installed = True


def load():
    """Return a namespace with the reference classes used on the hot path."""
    install()
    ns = types.SimpleNamespace()
    from irs_lqr.irs_lqr import IrsLqr, IrsLqrParameters
    from irs_lqr.irs_lqr_zero_order import IrsLqrZeroOrder
    from irs_lqr.irs_lqr_first_order import IrsLqrFirstOrder
    from irs_lqr.irs_lqr_exact import IrsLqrExact
    from pendulum_dynamics import PendulumDynamics
    from bicycle_dynamics import BicycleDynamics
    from quadrotor_dynamics import QuadrotorDynamics
    from three_cart_dynamics import ThreeCartDynamics
    ns.IrsLqr, ns.IrsLqrParameters = IrsLqr, IrsLqrParameters
    ns.IrsLqrZeroOrder, ns.IrsLqrFirstOrder, ns.IrsLqrExact = (
        IrsLqrZeroOrder, IrsLqrFirstOrder, IrsLqrExact)
    ns.PendulumDynamics, ns.BicycleDynamics = PendulumDynamics, BicycleDynamics
    ns.QuadrotorDynamics, ns.ThreeCartDynamics = QuadrotorDynamics, ThreeCartDynamics
    return ns
