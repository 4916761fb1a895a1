"""Drop-in for reference Static/inductive/models.py (GNNLayer, RED_GNN_induc)."""
from ..layers import GNNLayer, RedGNN

__all__ = ["GNNLayer", "RED_GNN_induc"]


class RED_GNN_induc(RedGNN):
    def forward(self, subs, rels, mode='transductive'):
        """models.py:65-89: 'transductive' -> training graph / n_ent, else unseen-entity graph / n_ent_ind."""
        graph = self.loader.graph_for(mode, self.W_final.weight.device)
        return self._run(subs, rels, graph, self.loader.n_ent_for(mode))

    def _graph_and_width(self, mode):
        mode = 'transductive' if mode is None else mode
        return self.loader.graph_for(mode, self.W_final.weight.device), self.loader.n_ent_for(mode)
