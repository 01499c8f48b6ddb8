"""Drop-in proof: the reference's own example scripts run UNCHANGED against this repository.

tests/golden/scripts/*.py.txt hold the verbatim computational bodies of five reference scripts
(examples/<system>/<script>.py up to their matplotlib section; written by
oracle/make_script_fixtures.py).  Each is exec()'d here with this repository on sys.path, so that

    from pendulum_dynamics import PendulumDynamics          -> ./pendulum_dynamics.py (CUDA functor)
    from irs_lqr.all import IrsLqrParameters, IrsLqrZeroOrder -> ./irs_lqr/ -> irs_mpc_b200

resolve to the drop-in modules; matplotlib (absent in this image, unused before the cut) is stubbed.
The only edit is the iteration count of `solver.iterate(k)` (test time).  Checked: the initial cost
the constructor computes (rollout + evaluate_cost) equals the one the REFERENCE's code computes
(tests/golden/reference_costs.json), the bookkeeping of `iterate` (k+2 entries, irs_lqr.py:196-218),
the deterministic exact script reproduces the reference's stored curve, and the result files the
scripts write are readable by the reference's plotting loaders.
"""
import json
import os
import re
import sys
import types

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCRIPTS = os.path.join(ROOT, "tests", "golden", "scripts")


class _Anything:
    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Anything()


def _stub(name):
    mod = types.ModuleType(name)
    mod.__path__ = []
    mod.__getattr__ = lambda attr: _Anything()
    return mod


@pytest.fixture()
def script_env(monkeypatch):
    for name in ("matplotlib", "matplotlib.pyplot", "mpl_toolkits", "mpl_toolkits.mplot3d"):
        if name not in sys.modules:
            monkeypatch.setitem(sys.modules, name, _stub(name))
    monkeypatch.syspath_prepend(ROOT)
    # the oracle's reference importer may have put the reference's own modules first (CPU tests)
    for k in [k for k in sys.modules if k == "irs_lqr" or k.startswith("irs_lqr.") or k.endswith("_dynamics")]:
        monkeypatch.delitem(sys.modules, k)
    np.random.seed(20211018)


def run_script(key, iterations):
    text = open(os.path.join(SCRIPTS, key + ".py.txt")).read()
    text, count = re.subn(r"solver\.iterate\(\d+\)", "solver.iterate(%d)" % iterations, text)
    assert count == 1
    ns = {"__name__": "__reference_script__"}
    exec(compile(text, key + ".py", "exec"), ns)
    return ns


@pytest.fixture(scope="module")
def costs():
    with open(os.path.join(ROOT, "tests", "golden", "reference_costs.json")) as f:
        return json.load(f)


# improves: the first descent must lower the initial cost.  Not asserted for the bicycle first-order script:
# with its sigma (1 rad on the heading, 2 m/s on the speed) the averaged Jacobians give a first descent
# ABOVE the initial cost — the float64 oracle with the bounded QP loop gives 3547 from 3302 on its own
# samples, this path 3685 — although the reference's stored bicycle_easy_first.csv (unpinned: stochastic,
# produced by an unknown revision of the script) shows 1867.  Nor for three_cart: its closure returns
# ABSOLUTE points, the literal fit is a linearization without meaning and the first
# descent lands at ~8e6 from 631 (the float64 oracle agrees on identical noise:
# tests/test_full_size_parity.py::test_cfg4_three_cart_zero_order_T100_N1e4[absolute]); what the test shows
# there is that the SECOND descent, from a trajectory of size |x| ~ 1e4 sigma, still fits (centred Gram).
@pytest.mark.parametrize("key,system,iterations,improves", [("pendulum_zero_order", "pendulum", 1, True),
                                                            ("bicycle_first_order", "bicycle", 1, False),
                                                            ("quadrotor_zero_order", "quadrotor", 1, True),
                                                            ("three_cart_zero_order", "three_cart", 2, False)])
def test_reference_script_body_runs_unchanged(script_env, costs, key, system, iterations, improves):
    ns = run_script(key, iterations)
    solver = ns["solver"]
    import irs_mpc_b200.irs_lqr as ours
    assert isinstance(solver, ours.IrsLqr)                       # the script got the CUDA-backed classes
    ref0 = costs["initial_cost_from_reference_code"][system]["cost"]
    assert abs(solver.cost_lst[0] - ref0) <= 1e-12 * abs(ref0)
    # iterate(k) performs k + 1 descents and logs k + 2 costs (irs_lqr.py:196-218)
    assert len(solver.cost_lst) == iterations + 2
    assert len(solver.x_trj_lst) == iterations + 2 and len(solver.u_trj_lst) == iterations + 2
    assert all(np.isfinite(c) for c in solver.cost_lst)
    if improves:
        assert solver.cost_lst[1] < solver.cost_lst[0]           # the first descent improves the initial guess
    T = ns["timesteps"]
    assert solver.x_trj_lst[-1].shape == (T + 1, solver.dim_x) and solver.x_trj_lst[-1].dtype == np.float64
    assert solver.cost == solver.cost_lst[iterations]            # the state keeps the k-th descent


def test_exact_script_reproduces_stored_curve_and_result_files_load(script_env, costs, tmp_path):
    """pendulum_exact.py end to end (7 iterations, the length of the stored curve) against examples/pendulum/analysis/pendulum_exact.csv,
    then the result-file round trip: the reference scripts write `cost_lst` with np.savetxt(...,
    delimiter=",") (quadrotor_cem.py:60, bicycle_cem_easy.py:49) or np.save (pendulum_cem.py:54) and
    examples/plot_iterations.py:14-17,33-42 reads them back with np.load / np.loadtxt(delimiter=",")."""
    gold = np.array(costs["stored_cost_curves"]["pendulum_exact"]["values"])
    ns = run_script("pendulum_exact", len(gold) - 2)      # the stored curve has k + 2 = 9 entries (k = 7)
    solver = ns["solver"]
    assert len(solver.cost_lst) == len(gold)
    np.testing.assert_allclose(np.array(solver.cost_lst), gold, rtol=1e-10)
    csv = str(tmp_path / "pendulum_exact.csv")
    np.savetxt(csv, solver.cost_lst, delimiter=",")
    back = np.loadtxt(csv, delimiter=",")
    assert back.shape == gold.shape and back.dtype == np.float64
    np.testing.assert_allclose(back, gold, rtol=1e-10)
    # byte-level format: one "%.18e" number per line, as in the reference's own csv
    first = open(csv).readline().strip()
    assert re.fullmatch(r"-?\d\.\d{18}e[+-]\d{2}", first)
    npy = str(tmp_path / "exact_cost.npy")
    np.save(npy, solver.cost_lst)
    back = np.load(npy)
    assert back.shape == gold.shape and back.dtype == np.float64
    # trajectories as the plotting code of the scripts consumes them: x_trj_lst[i][:, j]
    for x_trj in solver.x_trj_lst:
        assert isinstance(x_trj, np.ndarray) and x_trj.shape == (ns["timesteps"] + 1, 2)
