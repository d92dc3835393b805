"""MaxCut with one gamma per EDGE (public surface of reference qaoa/problems/maxcut_free_problem.py:31-99).

Edge e (in ``G.edges()`` order of the canonical graph, the order the reference's zero-padded parameter
names sort in) gets its own angle: layer d multiplies by ``exp(i sum_e gamma_{d,e} w_e [cut_e])``.  Lowered as
one cost term per edge (``qaoa_set_cost_maxkcut_terms``); registers beyond one CTA run layer by layer on the tile passes (X-type mixers)."""

from .maxcut_problem import MaxCut


class MaxCutFree(MaxCut):
    def term_of_edge(self):
        return list(range(self.graph_handler.num_edges))

    def get_num_parameters(self) -> int:
        return max(1, len(set(self.term_of_edge())))

    def lower_cost(self, engine):
        lut, fixed = self.colour_lut()
        engine.set_cost_maxkcut_terms(self.num_V, self.graph_handler.edge_list(), self.term_of_edge(),
                                      self.get_num_parameters(), self.N_qubits_per_node, lut,
                                      fixed_node=self.fix_one_node, fixed_colour=fixed)
