"""Generates tests/golden/*.json from the REFERENCE tree (run in the build container only).

The reference package cannot be imported as a whole (structlog / qiskit / qiskit-aer are not
installable here), but two of its modules are dependency-light and are loaded straight from
/root/reference by file path: qaoa/utils/graphutils.py (networkx) and qaoa/utils/statistic.py
(numpy).  Their OUTPUTS, plus the data fields of the reference's ExactCover result fixture, are
stored as golden vectors; no reference source is copied.

    python tests/golden/make_golden.py
"""
import importlib.util
import json
import os

import networkx as nx
import numpy as np

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def load(path, name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, path))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    graphutils = load("qaoa/utils/graphutils.py", "ref_graphutils")
    statistic = load("qaoa/utils/statistic.py", "ref_statistic")

    # ---- GraphHandler relabelling on a family of graphs ------------------------------------------
    graphs = {}
    cases = {
        "regular3_n8_seed42": nx.random_regular_graph(3, 8, seed=42),
        "regular3_n12_seed7": nx.random_regular_graph(3, 12, seed=7),
        "regular3_n16_seed42": nx.random_regular_graph(3, 16, seed=42),
        "barbell": nx.Graph([(0, 1)]),
        "path3": nx.path_graph(3),
        "house": nx.house_graph(),
        "star_relabel": nx.relabel_nodes(nx.star_graph(4), {0: "hub", 1: "a", 2: "b", 3: "c", 4: "d"}),
        "erdos_n10": nx.gnp_random_graph(10, 0.4, seed=3),
    }
    wg = nx.random_regular_graph(3, 10, seed=5)
    rng = np.random.RandomState(0)
    for u, v in wg.edges():
        wg[u][v]["weight"] = float(np.round(rng.rand() * 3, 3))
    cases["weighted_regular3_n10"] = wg
    for name, G in cases.items():
        gh = graphutils.GraphHandler(G)
        graphs[name] = {
            "input_nodes": [str(x) for x in G.nodes()],
            "input_edges": [[str(u), str(v), G[u][v].get("weight", 1)] for u, v in G.edges()],
            "relabelled_edges": [[int(u), int(v), gh.G[u][v].get("weight", 1)] for u, v in gh.G.edges()],
            "num_nodes": gh.num_nodes,
        }
    with open(os.path.join(OUT, "graphhandler.json"), "w") as f:
        json.dump(graphs, f, indent=1)

    # ---- Statistic on deterministic sample streams -------------------------------------------------
    stats = []
    rng = np.random.RandomState(1)
    for cvar in (1, 0.5, 0.25):
        for _ in range(3):
            k = int(rng.randint(3, 9))
            values = [float(x) for x in rng.randint(0, 12, size=k)]
            weights = [int(x) for x in rng.randint(1, 20, size=k)]
            st = statistic.Statistic(cvar=cvar)
            for i, (val, w) in enumerate(zip(values, weights)):
                st.add_sample(val, w, "s%d" % i)
            stats.append({"cvar": cvar, "values": values, "weights": weights, "E": float(st.get_E()),
                          "Variance": float(st.get_Variance()), "max": st.get_max(), "min": st.get_min(),
                          "CVaR": float(st.get_CVaR()), "max_sols": st.get_max_sols(), "min_sols": st.get_min_sols()})
    with open(os.path.join(OUT, "statistic.json"), "w") as f:
        json.dump(stats, f, indent=1)

    # ---- ExactCover end-to-end fixture (Dicke + Grover, 8192-shot histograms) ---------------------
    with open(os.path.join(REF, "examples/ExactCover/data/exact_cover_path_problem_example.json")) as f:
        ec = json.load(f)
    keep = {"columns": ec["problem"]["columns"], "weights": ec["problem"]["weights"],
            "hamming_weight": ec["problem"]["hamming_weight"], "solution": ec["problem"]["solution"],
            "cvar": ec["qaoa_params"]["cvar"], "depths": {}}
    for d in ("1", "3", "10"):
        keep["depths"][d] = {"optimal_angles": ec["qaoa_params"]["depths"][d]["optimal_angles"],
                             "histogram": ec["qaoa_params"]["depths"][d]["histogram"]}
    with open(os.path.join(OUT, "exactcover_fixture.json"), "w") as f:
        json.dump(keep, f, indent=1)
    print("golden vectors written to", OUT)


def result_file():
    """The reference's ExactCover result file (examples/ExactCover/data/exact_cover_path_problem_example.json), a DATA
    file written by the reference's qaoaIO, trimmed to depths 1-3: the on-disk schema qaoa_b200.utils.qaoaIO must read."""
    with open(os.path.join(REF, "examples/ExactCover/data/exact_cover_path_problem_example.json")) as f:
        d = json.load(f)
    d["qaoa_params"]["depths"] = {k: v for k, v in d["qaoa_params"]["depths"].items() if k in ("1", "2", "3")}
    with open(os.path.join(OUT, "qaoaio_result_example.json"), "w") as f:
        json.dump(d, f, indent=1)


def notebook_runs():
    """The example graphs (DATA files examples/MaxCut/data/*.gml, stored as node / weighted-edge lists in file order) and
    the numbers the reference's real pipeline (qiskit-aer, 1024 shots) left in its example notebooks' outputs: the
    `cost(depth d = ...)` log lines of `optimize`, the printed `computeMinMaxCosts()` and the "precalculated" minimum
    costs quoted in the notebooks' first cells."""
    import re

    def log_lines(nb, cell):
        with open(os.path.join(REF, nb)) as f:
            d = json.load(f)
        text = "".join("".join(o.get("text", [])) for o in d["cells"][cell].get("outputs", []))
        runs, cur = [], None
        for m in re.finditer(r"cost\(depth\s*(\d+)\s*=\s*(-?[0-9.]+)", text):
            depth, val = int(m.group(1)), float(m.group(2))
            if depth == 1:
                cur = []
                runs.append(cur)
            cur.append(val)
        return runs

    out = {"graphs": {}, "runs": {}}
    for name in ("er_n10_k4_0", "w_ba_n10_k4_0", "w_ba_n21_k4_0"):
        G = nx.read_gml(os.path.join(REF, "examples/MaxCut/data", name + ".gml"))
        out["graphs"][name] = {"nodes": [str(x) for x in G.nodes()],
                               "edges": [[str(u), str(v), G[u][v].get("weight", 1)] for u, v in G.edges()]}
    out["precalculated_mincost"] = {"w_ba_n10_k4_0": -8.657714089848158,      # ComparisonOptimizers.ipynb cell 4
                                    "w_ba_n21_k4_0": -25.23404480588015}      # CVaR.ipynb cell 4, WithFlip.ipynb cell 3
    out["printed_minmax"] = {"er_n10_k4_0": [-12, 0]}                         # FixOneQubit.ipynb cell 6 output
    # per notebook: the optimise runs in execution order, each a list of cost(depth 1), cost(depth 2), ...
    out["runs"]["FixOneQubit"] = {"graph": "er_n10_k4_0", "order": ["fix_one_node=False", "fix_one_node=True"],
                                  "cost": log_lines("examples/MaxCut/FixOneQubit.ipynb", 11)}
    out["runs"]["ComparisonOptimizers"] = {"graph": "w_ba_n10_k4_0", "order": ["spsa", "qnspsa", "neldermead", "cobyla"],
                                           "cost": log_lines("examples/MaxCut/ComparisonOptimizers.ipynb", 10)}
    out["runs"]["CVaR"] = {"graph": "w_ba_n21_k4_0", "order": ["cvar=1", "cvar=0.1"],
                           "cost": log_lines("examples/MaxCut/CVaR.ipynb", 8)}
    out["runs"]["MixerComparison"] = {"graph": "house", "order": ["vanilla", "free", "orbit"],
                                      "cost": log_lines("examples/MaxCut/MixerComparison.ipynb", 15)}
    # ExactCover/CompareGroverXY.ipynb: the problem of the result file (tests/golden/qaoaio_result_example.json) rebuilt with
    # TODAY's scale_problem=True, Dicke(2) + Grover(Dicke(2)), cvar = 0.7 (from the file), COBYLA, optimize(depth=10)
    with open(os.path.join(REF, "examples/ExactCover/CompareGroverXY.ipynb")) as f:
        cell5 = "".join("".join(o.get("text", [])) for o in json.load(f)["cells"][5].get("outputs", []))
    m = re.search(r"Optimal solution \(brute force\):\s*([01]+)\s*Cost:\s*(-?[0-9.]+)", cell5)
    out["runs"]["CompareGroverXY"] = {"problem": "qaoaio_result_example.json", "order": ["grover"], "cvar": 0.7,
                                      "hamming_weight": 2, "optimal_solution": m.group(1),
                                      "optimal_cost_printed_6_digits": float(m.group(2)),
                                      "cost": log_lines("examples/ExactCover/CompareGroverXY.ipynb", 7)}
    # ExactCover/ExactCover 6 3.ipynb: tail-assignment instance data/FRCR_6_24_3.txt (a DATA file, read with the format
    # the notebook's loader documents: "weight:row,row,..." per column, then the best state), Plus + X, interpolate=False,
    # landscape {"gamma": [0, 6 pi, 60], "beta": [0, 2 pi, 20]}; cell 7 printed computeMinMaxCosts() = (6.0, 10.0)
    rows = [ln.strip() for ln in open(os.path.join(REF, "examples/ExactCover/data/FRCR_6_24_3.txt")) if ln.strip()]
    R, F = 6, 24
    cols = [[0] * R for _ in range(F)]
    weights = []
    for r in range(R):
        w, idx = rows[r].split(":")
        weights.append(float(w))
        for i in idx.split(","):
            cols[int(i)][r] = 1
    out["runs"]["ExactCover_6_3"] = {"columns": cols, "weights": weights, "best": [int(x) for x in rows[R].split(",")],
                                     "printed_minmax": [6.0, 10.0], "order": ["plus_x_gridsearch"],
                                     "cost": log_lines("examples/ExactCover/ExactCover 6 3.ipynb", 6)}
    # PortfolioOptimization/PortOpt.ipynb: 12 assets, budget 4, gamma_scale 50.  The asset data come from qiskit_finance's
    # RandomDataProvider ([3p], not installable) seeded from data/qiskit_finance_seeds.npz (a DATA file); its published
    # algorithm -- per ticker cumsum(default_rng(seed).standard_normal(days)) + integers(1, 101), period returns
    # x[t+1]/x[t] - 1, their mean vector and np.cov -- is restated here and CHECKED against the exp_return vector the
    # notebook printed (cell 4 output, 9 digits).  Cell 9 printed the brute-force optimum, cell 20 computeMinMaxCosts(),
    # cell 19 the optimize log of six runs: cvar in (0.1, 1) x (penalty + X, XY ring [interpolate=False], Grover).
    seeds = np.load(os.path.join(REF, "examples/PortfolioOptimization/data/qiskit_finance_seeds.npz"))
    n_assets, days = 12, 101
    rng = np.random.default_rng(int(seeds[str(n_assets)][132]))
    series = np.array([np.maximum(np.cumsum(rng.standard_normal(days)) + rng.integers(1, 101), 0) for _ in range(n_assets)])
    returns = series[:, 1:] / series[:, :-1] - 1
    printed = [2.51714041e-03, -6.60209726e-04, -1.70699934e-03, 7.30592102e-04, -7.68369448e-05, 8.56893545e-04,
               -6.84333999e-03, -1.94220999e-03, 4.44702531e-03, -4.76239370e-04, -3.90803496e-03, 9.03679818e-04]
    assert np.abs(returns.mean(axis=1) - np.array(printed)).max() < 5e-12, "RandomDataProvider restatement is off"
    out["runs"]["PortOpt"] = {
        "n_assets": n_assets, "budget": 4, "gamma_scale": 50, "penalty_of_the_penalty_runs": 200,
        "exp_return": returns.mean(axis=1).tolist(), "cov_matrix": np.cov(returns).tolist(),
        "exp_return_printed": printed, "best_solution_printed": "100001001001",
        "printed_minmax": [-0.29191609262381646, 0.9214724310800932],
        "order": ["penalty cvar=0.1", "xy_ring cvar=0.1", "grover cvar=0.1", "penalty cvar=1", "xy_ring cvar=1", "grover cvar=1"],
        "cost": log_lines("examples/PortfolioOptimization/PortOpt.ipynb", 19)}
    with open(os.path.join(OUT, "notebook_runs.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
    result_file()
    notebook_runs()
