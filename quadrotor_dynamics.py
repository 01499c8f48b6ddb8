"""Drop-in import path of the reference's example scripts: `from quadrotor_dynamics import QuadrotorDynamics`
(examples/quadrotor/quadrotor_*.py import their system from the sibling module
examples/quadrotor/quadrotor_dynamics.py).  With this repository on sys.path the same statement resolves
to the CUDA-backed class; the reference module needs pydrake."""
from irs_mpc_b200.systems import QuadrotorDynamics  # noqa: F401
