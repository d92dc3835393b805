"""Shared builders for the parity tests (oracle side + engine side from the same inputs)."""
import networkx as nx
import numpy as np

from oracle import qaoa_oracle as orc


def regular_maxcut(n, seed=42, degree=3):
    if (n * degree) % 2:
        degree += 1  # odd n: a 3-regular graph does not exist
    G = nx.random_regular_graph(degree, n, seed=seed)
    Gr = orc.relabel_graph(G)
    edges = orc.graph_edges(Gr)
    colors, table = orc.colour_table(2)
    return edges, colors, table


def maxcut_diag(n, edges, colors, table):
    return orc.maxkcut_diag(edges, n, 1, table, colors)


def random_qubo(n, seed=42):
    rng = np.random.RandomState(seed)
    Q = rng.rand(n, n)
    c = 3 * (rng.rand(n) - 0.5)
    return Q, c, 0.0


def lut_from_table(table, bits):
    """device LUT: integer whose bit j is qubit v*b+j -> colour id (oracle: string char j = bit j)."""
    names = sorted(set(table.values()), key=lambda s: int(s[5:]))
    cid = {name: i for i, name in enumerate(names)}
    lut = np.zeros(1 << bits, dtype=np.uint8)
    for val in range(1 << bits):
        lab = "".join("1" if (val >> j) & 1 else "0" for j in range(bits))
        lut[val] = cid[table[lab]]
    return lut, cid


def rel_err(a, b):
    return abs(a - b) / max(1e-300, abs(b))


def reference_validate_cases():
    """The problems of the reference's unittests/test_validate_circuits.py (:11-147), by name -> (problem factory, t).
    (The Portfolio cases of :150-223 need qiskit_finance data and are covered by the randomised Portfolio parity tests.)"""
    import json
    import os

    from qaoa_b200 import problems

    def graph(edges, n):
        G = nx.Graph()
        G.add_nodes_from(range(n))
        G.add_weighted_edges_from(edges)
        return G

    def gml():
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "notebook_runs.json")
        with open(path) as f:
            g = json.load(f)["graphs"]["w_ba_n10_k4_0"]
        G = nx.Graph()
        G.add_nodes_from(g["nodes"])
        for u, v, w in g["edges"]:
            G.add_edge(u, v, weight=w)
        return G

    bowtie = [(0, 1, 1.0), (0, 2, 1.0), (1, 2, 1.0), (3, 2, 1.0), (3, 4, 1.0), (4, 2, 1.0)]
    maxcut = lambda G: problems.MaxKCutBinaryPowerOfTwo(G=G, k_cuts=2)      # noqa: E731
    rs = lambda: np.random.RandomState(42)                                   # noqa: E731

    def qubo_c():
        rng = rs()
        Q = rng.rand(8, 8)
        return problems.QUBO(Q, 3 * (rng.rand(8) - 0.5))

    return {
        "small_maxcut": (lambda: maxcut(graph([(0, 1, 1.0)], 2)), 1),
        "small_maxcut_t1.5": (lambda: maxcut(graph([(0, 1, 1.0)], 2)), 1.5),
        "larger_maxcut": (lambda: maxcut(graph(bowtie, 5)), 1),
        "weighted_maxcut_gml": (lambda: maxcut(gml()), 1),
        "qubo_linear_offset": (lambda: problems.QUBO(-np.array([[-3, 0, 0], [2, 1, 0], [3, 0, 3]]), -np.array([1, -2, 3]), 0.0), 1),
        "qubo_combined": (lambda: problems.QUBO(-np.array([[-2, 0, 0], [2, -1, 0], [3, 0, 6]])), 1),
        "qubo_2x2_diagonal": (lambda: problems.QUBO(-np.array([[1, 0], [0, 2]])), 1),
        "qubo_2x2_full": (lambda: problems.QUBO(np.array([[1, 3], [3, 2]])), 1),
        "qubo_large_symmetric": (lambda: problems.QUBO(10 * (rs().rand(8, 8) + rs().rand(8, 8).T)), 1),
        "qubo_large_asymmetric": (lambda: problems.QUBO(rs().rand(8, 8)), 1),
        "qubo_large_asymmetric_linear": (qubo_c, 1),
    }
