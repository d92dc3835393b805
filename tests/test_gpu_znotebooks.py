"""GPU: the reference's example notebooks replayed through the public API on the device, against the numbers those
notebooks hold (tests/golden/notebook_runs.json, written by tests/golden/make_golden.py from the reference tree):
the quoted minimum costs of the real-weighted example graphs and the depth-1 energies qiskit-aer produced (1024 shots),
which the exact engine must meet within shot noise.  (Runs last among the GPU tests on purpose: file name.)"""
import json
import math
import os

import networkx as nx
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _gold():
    with open(os.path.join(GOLD, "notebook_runs.json")) as f:
        return json.load(f)


def _graph(gold, name):
    g = gold["graphs"][name]
    G = nx.Graph()
    G.add_nodes_from(g["nodes"])
    for u, v, w in g["edges"]:
        G.add_edge(u, v, weight=w)
    return G


def _qaoa(G, dtype, **kw):
    from qaoa_b200 import QAOA, initialstates, mixers, problems

    fix = kw.pop("fix_one_node", False)
    return QAOA(problem=problems.MaxKCutBinaryPowerOfTwo(G=G, k_cuts=2, fix_one_node=fix), mixer=mixers.X(),
                initialstate=initialstates.Plus(), dtype=dtype, **kw)


@pytest.mark.parametrize("dtype", ["complex128", "complex64"])
def test_quoted_minimum_costs(dtype):
    """ComparisonOptimizers.ipynb cell 4 / CVaR.ipynb cell 4 (`mincost = ...  # precalculated`), FixOneQubit.ipynb cell 6
    (`computeMinMaxCosts()` printed -12 0): extremes of the device cost diagonal (F64 in complex128, 32-bit fixed point in
    complex64)."""
    gold = _gold()
    rel = 1e-12 if dtype == "complex128" else 1e-8
    for name, want in gold["precalculated_mincost"].items():
        q = _qaoa(_graph(gold, name), dtype)
        lo, hi = q._get_engine().cost_minmax()
        assert -hi == pytest.approx(want, rel=rel) and abs(lo) <= rel * abs(want), (name, lo, hi)
    q = _qaoa(_graph(gold, "er_n10_k4_0"), dtype)
    lo, hi = q._get_engine().cost_minmax()
    assert [-hi, -lo] == gold["printed_minmax"]["er_n10_k4_0"]            # integer weights: exact
    if dtype == "complex128":       # the reference's own call (FixOneQubit.ipynb cell 6), and beyond its brute-force reach
        assert list(q.problem.computeMinMaxCosts(engine=q._get_engine())) == gold["printed_minmax"]["er_n10_k4_0"]
        big = _qaoa(_graph(gold, "w_ba_n21_k4_0"), dtype).problem
        lo21, hi21 = big.computeMinMaxCosts()                              # 21 qubits: device reduction
        assert lo21 == pytest.approx(gold["precalculated_mincost"]["w_ba_n21_k4_0"], rel=1e-12) and hi21 == 0.0


def test_depth1_energies_meet_aer_within_shot_noise():
    """optimize(depth=1) on the device (exact energies, COBYLA from the landscape seed) against the `cost(depth 1 = ...)`
    values of FixOneQubit.ipynb (both circuits), ComparisonOptimizers.ipynb (four optimisers) and CVaR.ipynb (21 qubits)."""
    gold = _gold()

    def exact_depth1(name, **kw):
        q = _qaoa(_graph(gold, name), "complex128", **kw)
        q.optimize(depth=1, angles={"gamma": [0, np.pi, 24], "beta": [0, np.pi / 2, 12]})
        return q.get_Exp(1), math.sqrt(q.get_Var(1) / 1024.0)

    run = gold["runs"]["FixOneQubit"]
    for fix, costs in zip((False, True), run["cost"]):
        got, sigma = exact_depth1(run["graph"], fix_one_node=fix)
        assert abs(costs[0] - got) < 4 * sigma, (fix, costs[0], got, sigma)
    run = gold["runs"]["ComparisonOptimizers"]
    got, sigma = exact_depth1(run["graph"])
    for name, costs in zip(run["order"], run["cost"]):
        assert abs(costs[0] - got) < 4 * sigma, (name, costs[0], got, sigma)
    run = gold["runs"]["CVaR"]
    got, sigma = exact_depth1(run["graph"])
    assert abs(run["cost"][0][0] - got) < 4 * sigma, (run["cost"][0][0], got, sigma)


@pytest.mark.parametrize("dtype", ["complex128", "complex64"])
def test_exactcover_grover_cvar_run(dtype):
    """examples/ExactCover/CompareGroverXY.ipynb cells 5-7 on the device: ExactCover (scale_problem=True), Dicke(2) +
    Grover(Dicke(2)), cvar = 0.7, default landscape, COBYLA, INTERP; qiskit-aer printed 0.178078, 0.110158, 0.082881, 0.082268
    for depths 1-4 (1024 shots)."""
    from qaoa_b200 import QAOA, initialstates, mixers, problems

    gold = _gold()
    run = gold["runs"]["CompareGroverXY"]
    with open(os.path.join(GOLD, run["problem"])) as f:
        pr = json.load(f)["problem"]
    ec = problems.ExactCover(columns=np.array(pr["columns"]), weights=np.array(pr["weights"]),
                             hamming_weight=run["hamming_weight"], scale_problem=True)
    k = run["hamming_weight"]
    q = QAOA(problem=ec, mixer=mixers.Grover(initialstates.Dicke(k)), initialstate=initialstates.Dicke(k), cvar=run["cvar"],
             dtype=dtype)
    q.optimize(depth=4)
    for d, aer in zip((1, 2, 3, 4), run["cost"][0]):
        assert q.get_Exp(d) == pytest.approx(aer, abs=1.5e-3), (d, q.get_Exp(d), aer)


def test_cvar_of_real_valued_costs_at_21_qubits():
    """examples/MaxCut/CVaR.ipynb: MaxCut on the real-weighted 21-node graph with cvar = 0.1.  Real-valued costs: the exact
    CVaR comes from the device's weighted quantile select over the kept final state (qaoa_cvar; 32-bit fixed-point cost
    keys, four digit sweeps).  qiskit-aer printed cost(depth 1) = -23.080 (the best 102 of 1024 shots); the exact depth-1
    optimum of CVaR(0.1) is -23.01."""
    gold = _gold()
    run = gold["runs"]["CVaR"]
    q = _qaoa(_graph(gold, run["graph"]), "complex64", cvar=0.1)
    q.optimize(depth=1, angles={"gamma": [0, np.pi, 12], "beta": [0, np.pi / 2, 8]})
    assert hasattr(q._get_engine(), "cvar") and getattr(q, "_cvar_on_host", None) is None    # the device path, no read-back
    assert abs(q.get_Exp(1) - run["cost"][1][0]) < 0.25, (q.get_Exp(1), run["cost"][1][0])
    assert q.get_Exp(1) > gold["precalculated_mincost"][run["graph"]]


@pytest.mark.parametrize("name", sorted(__import__("parity_helpers").reference_validate_cases()))
def test_validate_circuit_cases_of_the_reference_on_the_device(name):
    """unittests/test_validate_circuits.py:11-147 with the device's cost diagonal standing where the reference has the
    phase-separator circuit's unitary: exp(i t c_dev(x)) equals exp(i t cost(string)) up to a global phase."""
    import parity_helpers as H
    from qaoa_b200 import QAOA, initialstates, mixers

    make_problem, t = H.reference_validate_cases()[name]
    q = QAOA(problem=make_problem(), mixer=mixers.X(), initialstate=initialstates.Plus())
    ok, report = q.validate_circuit(t=t)
    assert ok, report
    ok2, report2 = make_problem().validate_circuit(t=t)
    assert ok2 and report2["max_cost_error"] < 1e-9, report2
