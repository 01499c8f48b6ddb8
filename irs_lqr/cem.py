from irs_mpc_b200.cem import CemParameters, CrossEntropyMethod  # noqa: F401
