"""TEST INFRASTRUCTURE ONLY — fixtures that need the reference tree (build container only).

    python -m oracle.make_script_fixtures

Writes
  tests/golden/scripts/<script>.py.txt   the COMPUTATIONAL body of four reference example scripts
                                         (imports ... solver.iterate(...) and the two prints), cut
                                         before the matplotlib section.  Verbatim reference text,
                                         kept as a test FIXTURE (never imported by the product):
                                         tests/test_dropin_scripts.py exec()s them against this
                                         repository's `irs_lqr` / `<system>_dynamics` shims to
                                         show that the reference's own scripts run unchanged.
  tests/golden/reference_timing.json     wall-clock of the UNMODIFIED reference's
                                         IrsLqrZeroOrder.get_TV_matrices in this container
                                         (pendulum at full example size, quadrotor at N=100: its
                                         dynamics_batch is a per-sample Python loop), anchoring the
                                         CPU baseline on reference code.
"""
import json
import os
import platform
import time

import numpy as np

from oracle import example_configs, ref_import
from oracle.make_golden import GOLDEN_DIR, SYSTEM_CLASS, make_params

SCRIPTS = {
    "pendulum_zero_order": "pendulum/pendulum_zero_order.py",
    "bicycle_first_order": "bicycle/bicycle_first_order.py",
    "quadrotor_zero_order": "quadrotor/quadrotor_zero_order.py",
    "three_cart_zero_order": "three_cart/three_cart_zero_order.py",
    "pendulum_exact": "pendulum/pendulum_exact.py",
}
HEADER = ("# TEST FIXTURE: verbatim computational body of /root/reference/examples/%s (hjsuh94/irs_mpc),\n"
          "# cut before its matplotlib section by oracle/make_script_fixtures.py.  Executed by\n"
          "# tests/test_dropin_scripts.py against this repository's drop-in modules.  Not product code.\n")


def script_bodies():
    out_dir = os.path.join(GOLDEN_DIR, "scripts")
    os.makedirs(out_dir, exist_ok=True)
    for key, rel in SCRIPTS.items():
        lines = open(os.path.join(ref_import.REFERENCE_ROOT, "examples", rel)).read().splitlines()
        cut = max(i for i, l in enumerate(lines) if l.startswith("print(\"Elapsed time"))
        body = "\n".join(lines[:cut + 1]) + "\n"
        with open(os.path.join(out_dir, key + ".py.txt"), "w") as f:
            f.write(HEADER % rel + body)
        print("wrote", key)


def reference_timing():
    ns = ref_import.load()
    out = {"host": platform.processor() or platform.machine(), "cpu_count": os.cpu_count(),
           "numpy": np.__version__, "what": "unmodified reference IrsLqrZeroOrder.get_TV_matrices "
           "(irs_lqr/irs_lqr_zero_order.py:38-63) under the pydrake stub, single process", "cases": {}}
    for name, T, N in (("pendulum", 200, 1000), ("quadrotor", 100, 100), ("three_cart", 100, 10000),
                       ("bicycle", 100, 10000)):
        cfg = example_configs.CONFIGS[name](T=T)
        ref = getattr(ns, SYSTEM_CLASS[name])(cfg["h"])
        n, m = ref.dim_x, ref.dim_u
        sig = cfg["sigma"]

        def sampling(xbar, ubar, it, sig=sig, n=n, m=m, N=N, ref=ref, proj=cfg["projection"]):
            dx = np.random.normal(0.0, sig[:n] / (it ** 0.5), size=(N, n))
            du = np.random.normal(0.0, sig[n:] / (it ** 0.5), size=(N, m))
            if proj:
                return ref.projection(xbar, dx, ubar, du)
            return dx, du

        solver = ns.IrsLqrZeroOrder(ref, make_params(ns, cfg, T=T), sampling)
        solver.get_TV_matrices(solver.x_trj, solver.u_trj)      # warm
        best = 1e30
        for _ in range(3):
            t = time.perf_counter()
            solver.get_TV_matrices(solver.x_trj, solver.u_trj)
            best = min(best, time.perf_counter() - t)
        out["cases"][name] = {"T": T, "N": N, "seconds": best, "samples_per_s": T * N / best,
                              "mode": "zero_order"}
        print(name, out["cases"][name])
    with open(os.path.join(GOLDEN_DIR, "reference_timing.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    script_bodies()
    reference_timing()
