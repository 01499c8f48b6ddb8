from irs_mpc_b200.irs_lqr import *  # noqa: F401,F403
from irs_mpc_b200.irs_lqr import IrsLqr, IrsLqrParameters  # noqa: F401,E402
