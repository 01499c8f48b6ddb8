from irs_mpc_b200.dynamical_system import *  # noqa: F401,F403
