# diagnostic: A/B/x errors vs oracle, iid vs antithetic, quadrotor T=24
import sys, numpy as np
sys.path.insert(0, ".")
import irs_mpc_b200.all as api
from oracle import cpu_restatement as cr, example_configs as ec
def rel(a, b): return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))
T = 24
cfg = ec.quadrotor(T=T)
orc = cr.QuadrotorOracle(cfg["h"])
for anti, N in ((False, 1500), (True, 1500), (True, 3000), (False, 8192), (True, 8192), (True, 32768)):
    s = api.QuadrotorDynamics(cfg["h"])
    smp = api.GaussianSampling(cfg["sigma"][:12], cfg["sigma"][12:], N, seed=77, antithetic=anti)
    p = api.IrsLqrParameters()
    for k in ("Q", "Qd", "R", "x0", "xd_trj", "u_trj_initial", "xbound", "ubound"): setattr(p, k, cfg[k])
    sol = api.IrsLqrZeroOrder(s, p, smp)
    At, Bt, ct = sol.get_TV_matrices(sol.x_trj, sol.u_trj)
    xn, un = sol.local_descent(sol.x_trj, sol.u_trj)
    d = smp.deltas(T, 1).astype(np.float64)
    Ao, Bo, co = cr.zero_order_tv_matrices(orc, sol.x_trj, sol.u_trj, d)
    K, k = cr.tvlqr_riccati(Ao, Bo, co, cfg["Q"], cfg["Qd"], cfg["R"], cfg["xd_trj"])
    xo, uo = cr.closed_loop_descent(orc, K, k, sol.x_trj[0])
    K2, k2 = cr.tvlqr_riccati(At, Bt, ct, cfg["Q"], cfg["Qd"], cfg["R"], cfg["xd_trj"])
    x2, u2 = cr.closed_loop_descent(orc, K2, k2, sol.x_trj[0])
    print("anti=%s N=%d  A %.2e B %.2e c %.2e | x %.2e u %.2e | x(device fit, oracle solve) %.2e" % (
        anti, N, rel(At, Ao), rel(Bt, Bo), np.max(np.abs(ct - co)), rel(xn, xo), rel(un, uo), rel(xn, x2)))
