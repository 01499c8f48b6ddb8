"""Drop-in import path: `from irs_lqr.all import ...` resolves to irs_mpc_b200 (reference: irs_lqr/)."""
