#!/usr/bin/env python
"""Run one of the analytic iRS-LQR examples of the reference on the GPU.

    python examples/run_example.py --system quadrotor --order zero --iters 10
    python examples/run_example.py --system bicycle --order first --samples 10000
    python examples/run_example.py --system pendulum --order exact
    python examples/run_example.py --system three_cart --order zero --numpy-sampling

The problem data are those of the reference's scripts (examples/<system>/<system>_<order>_order.py,
see irs_mpc_b200/example_configs.py for the line references); the solver classes are imported from
the drop-in path `irs_lqr.all`, exactly as the reference's scripts do.  `--numpy-sampling` keeps
the reference's own sampling closure (np.random.normal on the host, called once per timestep and
replayed through the kernels); the default draws the noise inside the kernel (GaussianSampling).
`--csv` writes the cost list the way the reference's scripts do (np.savetxt of solver.cost_lst).
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from irs_lqr.all import (IrsLqrExact, IrsLqrFirstOrder, IrsLqrParameters, IrsLqrZeroOrder,  # noqa: E402
                         GaussianSampling)
from irs_mpc_b200 import example_configs as ec                                              # noqa: E402
from irs_mpc_b200.systems import SYSTEM_CLASSES                                             # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--system", default="pendulum", choices=sorted(ec.CONFIGS))
    ap.add_argument("--order", default="zero", choices=["zero", "first", "exact"])
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--samples", type=int, default=None, help="samples per timestep (default: the script's)")
    ap.add_argument("--horizon", type=int, default=None)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--numpy-sampling", action="store_true")
    ap.add_argument("--projection", default="delta", choices=["delta", "absolute"],
                    help="three_cart only: 'absolute' reproduces the reference literally (its sampling closure "
                         "returns projected ABSOLUTE points which the solver then uses as regressors, "
                         "three_cart_zero_order.py:43); 'delta' is the corrected variant (projected point minus nominal)")
    ap.add_argument("--csv", default=None)
    args = ap.parse_args()

    cfg = ec.CONFIGS[args.system](**({"T": args.horizon} if args.horizon else {}))
    system = SYSTEM_CLASSES[args.system](cfg["h"])
    params = IrsLqrParameters()
    for key in ("Q", "Qd", "R", "x0", "xd_trj", "u_trj_initial", "xbound", "ubound"):
        setattr(params, key, cfg[key])
    n = system.dim_x
    N = args.samples or cfg["num_samples"]
    sig_x, sig_u, power = cfg["sigma"][:n], cfg["sigma"][n:], cfg["power"]

    if args.numpy_sampling:
        np.random.seed(args.seed)

        def sampling(xbar, ubar, it):          # the reference's closure, e.g. pendulum_zero_order.py:38-43
            dx = np.random.normal(0.0, sig_x / (it ** power), size=(N, n))
            du = np.random.normal(0.0, sig_u / (it ** power), size=(N, system.dim_u))
            if cfg["projection"]:              # three_cart_zero_order.py:43 returns the projected ABSOLUTE points
                xp, up = system.projection(xbar, dx, ubar, du)
                return (xp, up) if args.projection == "absolute" else (xp - xbar, du)
            return dx, du
    else:
        sampling = GaussianSampling(sig_x, sig_u, N, power=power, seed=args.seed,
                                    projection=args.projection if cfg["projection"] else None)

    if args.order == "exact":
        solver = IrsLqrExact(system, params)
    elif args.order == "first":
        solver = IrsLqrFirstOrder(system, params, sampling)
    else:
        solver = IrsLqrZeroOrder(system, params, sampling)
    print("%s, %s order, T=%d, N=%d, initial cost %.6f" % (args.system, args.order, solver.T, N, solver.cost))
    t = time.time()
    solver.iterate(args.iters)
    print("final cost %.6f after %d iterations, %.3f s" % (solver.cost, args.iters, time.time() - t))
    if args.csv:
        np.savetxt(args.csv, solver.cost_lst, delimiter=",")


if __name__ == "__main__":
    main()
