from irs_mpc_b200.irs_lqr import IrsLqrZeroOrder  # noqa: F401
