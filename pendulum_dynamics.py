"""Drop-in import path of the reference's example scripts: `from pendulum_dynamics import PendulumDynamics`
(examples/pendulum/pendulum_*.py import their system from the sibling module
examples/pendulum/pendulum_dynamics.py).  With this repository on sys.path the same statement resolves
to the CUDA-backed class; the reference module needs pydrake."""
from irs_mpc_b200.systems import PendulumDynamics  # noqa: F401
