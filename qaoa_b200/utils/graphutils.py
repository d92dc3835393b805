"""Graph normalisation used by every graph problem.

Behaviourally equivalent to the node handling of the reference's GraphHandler
(reference qaoa/utils/graphutils.py:38-47, 49-72, 115-137): nodes become 0..n-1 and the first
node of maximum degree trades places with node n-1 (that is the node ``fix_one_node`` removes).
The reference additionally computes 100 random greedy edge colourings (:139-181) that only
decide the ORDER of commuting diagonal gates; a diagonal kernel has no use for them, so they
are computed lazily and only on request.
"""

import networkx as nx


class GraphHandler:
    def __init__(self, G, check_isomorphism=None):
        if isinstance(G, nx.DiGraph):
            raise Exception("Graph should be undirected.")
        if any(deg <= 1 for _node, deg in G.degree()):
            print(
                "Graph contains nodes with one or zero edges. These can be removed to reduce the size of the problem."
            )
        self.num_nodes = G.number_of_nodes()
        self.num_edges = G.number_of_edges()
        self.G = self._max_degree_last(self._integer_labels(G))
        # the reference always runs the (expensive) isomorphism check; by default only for small graphs
        if check_isomorphism is None:
            check_isomorphism = self.num_nodes <= 16
        if check_isomorphism and not nx.is_isomorphic(G, self.G):
            raise Exception("Something went wrong.")
        self._parallel_edges = None

    def _integer_labels(self, G):
        if set(G.nodes) == set(range(self.num_nodes)):
            return G
        return nx.relabel_nodes(G, {node: i for i, node in enumerate(G.nodes)}, copy=True)

    def _max_degree_last(self, G):
        last = self.num_nodes - 1
        top = max(G.degree(), key=lambda nd: nd[1])[0]  # first node attaining the maximum degree
        if top == last:
            return G
        return nx.relabel_nodes(G, {top: last, last: top}, copy=True)

    def edge_list(self):
        """[(u, v, weight)] in ``G.edges()`` order — the order GraphProblem.cost sums in."""
        return [(int(u), int(v), self.G[u][v].get("weight", 1)) for u, v in self.G.edges()]

    @property
    def parallel_edges(self):
        if self._parallel_edges is None:
            self._parallel_edges = self._edge_colouring()
        return self._parallel_edges

    def _edge_colouring(self, repetitions=100):
        line = nx.line_graph(self.G)
        best = None
        for _ in range(repetitions):
            colouring = nx.coloring.greedy_color(line, strategy="random_sequential")
            groups = {}
            for edge, colour in colouring.items():
                groups.setdefault(colour, []).append(edge)
            if best is None or len(groups) < len(best):
                best = groups
        return best or {}
