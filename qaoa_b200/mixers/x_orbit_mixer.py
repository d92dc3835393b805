"""X mixer with one angle per node orbit of the graph's automorphism group
(public surface of reference qaoa/mixers/x_orbit_mixer.py:53-151, arXiv:2410.05187).

Qubits whose nodes are structurally equivalent share one beta; the device sees an ordinary
multi-angle X mixer (one RX(-2 beta_q) per qubit, x_orbit_mixer.py:146-151) and ``engine_betas``
expands the per-orbit angles into per-qubit ones (orbit parameters are ordered by orbit index,
which is the order the reference's zero-padded parameter names sort in)."""

from collections import defaultdict

import networkx as nx
from networkx.algorithms.isomorphism import GraphMatcher

from .base_mixer import Mixer
from qaoa_b200 import engine as _eng
from qaoa_b200.utils.graphutils import GraphHandler


def node_orbits(G):
    """Orbits of the nodes of G under Aut(G): union-find over the automorphisms GraphMatcher enumerates,
    stopping early once everything has merged (x_orbit_mixer.py:97-124).  Orbits are listed in order of
    their union-find representative, nodes inside an orbit in sorted order."""
    nodes = sorted(G.nodes())
    index = {v: i for i, v in enumerate(nodes)}
    parent = list(range(len(nodes)))

    def find(x):
        while parent[x] != x:
            parent[x] = parent[parent[x]]
            x = parent[x]
        return x

    for auto in GraphMatcher(G, G).isomorphisms_iter():
        for i, v in enumerate(nodes):
            ri, rj = find(i), find(index[auto[v]])
            if ri != rj:
                parent[ri] = rj
        if len({find(i) for i in range(len(nodes))}) == 1:
            break
    groups = defaultdict(list)
    for i, v in enumerate(nodes):
        groups[find(i)].append(v)
    return list(groups.values())


class XOrbit(Mixer):
    def __init__(self, G: nx.Graph) -> None:
        self.label = "XOrbit"
        self.circuit = None
        self._canonical_G = GraphHandler(G).G
        self.node_orbits = node_orbits(self._canonical_G)
        self.node_to_orbit = {v: k for k, orbit in enumerate(self.node_orbits) for v in orbit}

    def get_num_parameters(self):
        return len(self.node_orbits)

    def engine_betas(self, betas):
        """per-orbit angles of one layer -> per-qubit angles (qubit i = i-th node in sorted order)"""
        nodes = sorted(self._canonical_G.nodes())
        return [betas[self.node_to_orbit[v]] for v in nodes]

    def lower_mixer(self, engine, trotter=1):
        engine.set_mixer(_eng.MIXER_XMULTI, trotter=trotter)
