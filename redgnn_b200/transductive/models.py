"""Drop-in for reference Static/transductive/models.py (GNNLayer, RED_GNN_trans)."""
from ..layers import GNNLayer, RedGNN

__all__ = ["GNNLayer", "RED_GNN_trans"]


class RED_GNN_trans(RedGNN):
    def forward(self, subs, rels, mode='train'):
        """models.py:65-89: scores (n, loader.n_ent) fp32; mode 'train' uses KG, anything else tKG."""
        graph = self.loader.graph_for(mode, self.W_final.weight.device)
        return self._run(subs, rels, graph, self.loader.n_ent)

    def _graph_and_width(self, mode):
        mode = 'train' if mode is None else mode
        return self.loader.graph_for(mode, self.W_final.weight.device), self.loader.n_ent
