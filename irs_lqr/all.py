from irs_mpc_b200.all import *  # noqa: F401,F403
