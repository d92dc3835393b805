"""MaxCut with one gamma per EDGE ORBIT of the graph's automorphism group
(public surface of reference qaoa/problems/maxcut_orbit_problem.py:51-181, arXiv:2410.05187).

Edges that an automorphism maps onto each other share an angle; orbits are numbered by their first edge in
``G.edges()`` order (the order of the reference's ``orbit_groups`` dictionary, :96-107)."""

from collections import defaultdict

from networkx.algorithms.isomorphism import GraphMatcher

from .maxcut_free_problem import MaxCutFree


class MaxCutOrbit(MaxCutFree):
    def __init__(self, G, method: str = "Diffusion", fix_one_node: bool = False) -> None:
        super().__init__(G, method=method, fix_one_node=fix_one_node)
        self._compute_edge_orbits()

    def _compute_edge_orbits(self) -> None:
        G = self.graph_handler.G
        edges = list(G.edges())
        index = {}
        for i, (u, v) in enumerate(edges):
            index[(u, v)] = i
            index[(v, u)] = i
        parent = list(range(len(edges)))

        def find(x):
            while parent[x] != x:
                parent[x] = parent[parent[x]]
                x = parent[x]
            return x

        for auto in GraphMatcher(G, G).isomorphisms_iter():
            for i, (u, v) in enumerate(edges):
                j = index.get((auto[u], auto[v]))
                if j is not None:
                    ri, rj = find(i), find(j)
                    if ri != rj:
                        parent[ri] = rj
            if len({find(i) for i in range(len(edges))}) == 1:
                break
        groups = defaultdict(list)
        for i, e in enumerate(edges):
            groups[find(i)].append(e)
        self.edge_orbits = list(groups.values())
        self.edge_to_orbit = {}
        for k, orbit in enumerate(self.edge_orbits):
            for u, v in orbit:
                self.edge_to_orbit[(u, v)] = k
                self.edge_to_orbit[(v, u)] = k

    def term_of_edge(self):
        return [self.edge_to_orbit[e] for e in self.graph_handler.G.edges()]
