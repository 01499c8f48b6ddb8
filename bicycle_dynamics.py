"""Drop-in import path of the reference's example scripts: `from bicycle_dynamics import BicycleDynamics`
(examples/bicycle/bicycle_*.py import their system from the sibling module
examples/bicycle/bicycle_dynamics.py).  With this repository on sys.path the same statement resolves
to the CUDA-backed class; the reference module needs pydrake."""
from irs_mpc_b200.systems import BicycleDynamics  # noqa: F401
