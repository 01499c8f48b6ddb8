"""Drop-in import path of the reference's example scripts: `from three_cart_dynamics import ThreeCartDynamics`
(examples/three_cart/three_cart_*.py import their system from the sibling module
examples/three_cart/three_cart_dynamics.py).  With this repository on sys.path the same statement resolves
to the CUDA-backed class; the reference module needs pydrake."""
from irs_mpc_b200.systems import ThreeCartDynamics  # noqa: F401
