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
