"""TEST INFRASTRUCTURE ONLY — regenerates tests/golden/ from the UNMODIFIED reference.

Run in the build container (needs /root/reference; cannot run on the GPU box):

    python -m oracle.make_golden

Writes
  tests/golden/dynamics_<system>.npz     inputs + outputs of the reference's dynamics,
                                         dynamics_batch (and three_cart projection)
  tests/golden/zero_order_<system>.npz   nominal trajectory, float32 sampled deltas, and the
                                         (At, Bt, ct) the reference's
                                         IrsLqrZeroOrder.get_TV_matrices returns when its
                                         `sampling` closure replays exactly those deltas
  tests/golden/reference_costs.json      the stored cost curves the reference ships
                                         (examples/*/analysis/*.csv) and the initial-guess costs
                                         computed by the reference's own rollout+evaluate_cost
All arrays are produced by reference code (irs_lqr/irs_lqr.py, irs_lqr/irs_lqr_zero_order.py,
examples/*/…_dynamics.py); nothing from oracle/cpu_restatement.py is involved.
"""
import json
import os

import numpy as np

from oracle import example_configs, ref_import

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                          "tests", "golden")

SYSTEM_CLASS = {"pendulum": "PendulumDynamics", "bicycle": "BicycleDynamics",
                "quadrotor": "QuadrotorDynamics", "three_cart": "ThreeCartDynamics"}
# spread of the random evaluation points per system (quadrotor: stay away from pitch = pi/2)
POINT_SCALE = {"pendulum": 2.0, "bicycle": 1.0, "quadrotor": 0.5, "three_cart": 2.0}


def make_params(ns, cfg, T=None, x0=None, u_trj=None):
    p = ns.IrsLqrParameters()
    T = cfg["T"] if T is None else T
    p.Q, p.Qd, p.R = cfg["Q"], cfg["Qd"], cfg["R"]
    p.x0 = cfg["x0"] if x0 is None else x0
    p.xd_trj = cfg["xd_trj"][:T + 1]
    p.u_trj_initial = cfg["u_trj_initial"][:T] if u_trj is None else u_trj
    p.xbound, p.ubound = cfg["xbound"], cfg["ubound"]
    return p


def golden_dynamics(ns, name, rng):
    cfg = example_configs.CONFIGS[name]()
    ref = getattr(ns, SYSTEM_CLASS[name])(cfg["h"])
    B = 512
    s = POINT_SCALE[name]
    x = s * rng.standard_normal((B, ref.dim_x))
    u = s * rng.standard_normal((B, ref.dim_u))
    if name == "quadrotor":
        u += 2.0
    out = dict(x=x, u=u,
               f_batch=ref.dynamics_batch(x.copy(), u.copy()),
               f_scalar=np.stack([ref.dynamics(x[i].copy(), u[i].copy()) for i in range(B)]))
    if name == "three_cart":
        xbar, ubar = cfg["x0"], cfg["u_trj_initial"][0]
        dx = 4.0 * rng.standard_normal((B, 6))
        du = 0.5 * rng.standard_normal((B, 2))
        xp, up = ref.projection(xbar, dx.copy(), ubar, du.copy())
        out.update(proj_xbar=xbar, proj_ubar=ubar, proj_dx=dx, proj_du=du, proj_x=xp, proj_u=up)
    np.savez(os.path.join(GOLDEN_DIR, "dynamics_%s.npz" % name), **out)


def golden_zero_order(ns, name, rng, T=5, N=1000):
    cfg = example_configs.CONFIGS[name]()
    ref = getattr(ns, SYSTEM_CLASS[name])(cfg["h"])
    n, m = ref.dim_x, ref.dim_u
    # a short problem starting somewhere interesting on the example's initial rollout
    full = ns.IrsLqrZeroOrder(ref, make_params(ns, cfg), None)
    t0 = cfg["T"] // 3
    x0 = full.x_trj[t0] + 0.05 * rng.standard_normal(n)
    u_trj = cfg["u_trj_initial"][:T] + 0.05 * rng.standard_normal((T, m))
    sigma = cfg["sigma"]
    eps = rng.standard_normal((T, N, n + m)).astype(np.float32)
    deltas = (eps * sigma.astype(np.float32)).astype(np.float32)     # float32 on purpose
    state = {"t": 0}

    def sampling(xbar, ubar, it):
        t = state["t"]
        state["t"] += 1
        dx = deltas[t][:, :n].astype(np.float64)
        du = deltas[t][:, n:].astype(np.float64)
        if cfg["projection"]:
            # three_cart/three_cart_zero_order.py:38-43: the closure returns projection(...)
            return ref.projection(xbar, dx, ubar, du)
        return dx, du

    solver = ns.IrsLqrZeroOrder(ref, make_params(ns, cfg, T=T, x0=x0, u_trj=u_trj), sampling)
    At, Bt, ct = solver.get_TV_matrices(solver.x_trj, solver.u_trj)
    np.savez(os.path.join(GOLDEN_DIR, "zero_order_%s.npz" % name),
             x_trj=solver.x_trj, u_trj=solver.u_trj, deltas=deltas, sigma=sigma,
             At=At, Bt=Bt, ct=ct, projection=np.array(cfg["projection"]),
             initial_cost=np.array(solver.cost))


def golden_costs(ns):
    ex = os.path.join(ref_import.REFERENCE_ROOT, "examples")
    files = {
        "pendulum_exact": "pendulum/analysis/pendulum_exact.csv",
        "pendulum_first_order": "pendulum/analysis/pendulum_first_order.csv",
        "pendulum_zero_order": "pendulum/analysis/pendulum_zero_order.csv",
        "quadrotor_exact": "quadrotor/analysis/quadrotor_exact.csv",
        "quadrotor_first": "quadrotor/analysis/quadrotor_first.csv",
        "quadrotor_zero": "quadrotor/analysis/quadrotor_zero.csv",
        "bicycle_easy_exact": "bicycle/analysis/bicycle_easy_exact.csv",
        "bicycle_easy_first": "bicycle/analysis/bicycle_easy_first.csv",
        "bicycle_easy_zero": "bicycle/analysis/bicycle_easy_zero.csv",
        "bicycle_hard_exact": "bicycle/analysis/bicycle_hard_exact.csv",
    }
    out = {"stored_cost_curves": {}, "initial_cost_from_reference_code": {}}
    for key, rel in files.items():
        out["stored_cost_curves"][key] = {
            "source": "examples/" + rel,
            "values": [float(v) for v in np.loadtxt(os.path.join(ex, rel)).ravel()]}
    for name in ("pendulum", "bicycle", "quadrotor", "three_cart"):
        cfg = example_configs.CONFIGS[name]()
        ref = getattr(ns, SYSTEM_CLASS[name])(cfg["h"])
        base = ns.IrsLqr(ref, make_params(ns, cfg))      # rollout + evaluate_cost, irs_lqr.py:61-62
        out["initial_cost_from_reference_code"][name] = {
            "cost": float(base.cost), "x_final": [float(v) for v in base.x_trj[-1]]}
    with open(os.path.join(GOLDEN_DIR, "reference_costs.json"), "w") as f:
        json.dump(out, f, indent=1)


def main():
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    ns = ref_import.load()
    rng = np.random.default_rng(20211018)
    for name in ("pendulum", "bicycle", "quadrotor", "three_cart"):
        golden_dynamics(ns, name, rng)
        golden_zero_order(ns, name, rng)
    golden_costs(ns)
    print("golden vectors written to", GOLDEN_DIR)


if __name__ == "__main__":
    main()
