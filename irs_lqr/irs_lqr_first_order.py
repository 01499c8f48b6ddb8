from irs_mpc_b200.irs_lqr import IrsLqrFirstOrder  # noqa: F401
