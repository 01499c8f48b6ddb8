from irs_mpc_b200.tv_lqr import *  # noqa: F401,F403
from irs_mpc_b200.tv_lqr import get_solver, solve_tvlqr  # noqa: F401,E402
