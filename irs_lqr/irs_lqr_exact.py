from irs_mpc_b200.irs_lqr import IrsLqrExact  # noqa: F401
