"""GMF / MLP / NeuMF with the reference's interface (pytorch_models/NeuMF.py:9-138) and ``NeuMF.init`` for the
pre-train -> fuse -> train schedule of ``main.main_NeuMF`` (main.py:289-340).  Id-embedding models: every
lookup, small linear layer and gradient scatter runs in the same kernels as the MF family."""
import torch
import torch.nn as nn

from .. import ops
from .common_pytorch_models import SmallLinear
from .MF import IdEmbedding


class _Biases(nn.Module):
    """What the three classes share (NeuMF.py:14-16, 43-45, 79-81): user / item bias vectors with
    ``total_users + 1`` / ``total_items + 1`` entries initialised to 0.1 and a global bias of 4.0, gathered per
    rating with ``r4r_rows_gather`` (dense gradient scatter in the backward, like ``Tensor.gather``)."""

    def __init__(self, hyper_params):
        super().__init__()
        self.hyper_params = hyper_params
        self.user_bias = nn.Parameter(torch.full((hyper_params["total_users"] + 1,), 0.1))
        self.item_bias = nn.Parameter(torch.full((hyper_params["total_items"] + 1,), 0.1))
        self.global_bias = nn.Parameter(torch.full((1,), 4.0))

    def _bias_sum(self, user_id, item_id):
        shape = tuple(user_id.shape)
        ub = ops.rows_gather(self.user_bias, user_id.reshape(-1)).view(shape)
        ib = ops.rows_gather(self.item_bias, item_id.reshape(-1)).view(shape)
        return ub + ib + self.global_bias


class GMF(_Biases):
    """Generalised matrix factorisation (NeuMF.py:9-36): a learned linear read-out of the elementwise product of
    the dropped-out user and item embeddings, plus the biases.  First model of the NeuMF pre-training schedule."""

    def __init__(self, hyper_params):
        super().__init__(hyper_params)
        L = hyper_params["latent_size"]
        self.user_embedding = IdEmbedding(hyper_params["total_users"] + 1, L)
        self.item_embedding = IdEmbedding(hyper_params["total_items"] + 1, L)
        self.final = SmallLinear(L, 1)
        self.dropout = nn.Dropout(hyper_params["dropout"])

    def forward(self, data):
        _, _, _, _, _, user_id, item_id = data
        user = self.dropout(self.user_embedding(user_id.reshape(-1)))
        item = self.dropout(self.item_embedding(item_id.reshape(-1)))
        rating = self.final(user * item)[:, 0].view(tuple(user_id.shape))
        return self._bias_sum(user_id, item_id) + rating


class MLP(_Biases):
    """The MLP half (NeuMF.py:38-73): the concatenated embeddings go through Dropout -> Linear(2L, L) -> ReLU ->
    Linear(L, L) (``project``, indices 1 and 3 carry the weights) and a linear read-out, plus the biases."""

    def __init__(self, hyper_params):
        super().__init__(hyper_params)
        L, p = hyper_params["latent_size"], hyper_params["dropout"]
        self.user_embedding = IdEmbedding(hyper_params["total_users"] + 1, L)
        self.item_embedding = IdEmbedding(hyper_params["total_items"] + 1, L)
        self.project = nn.Sequential(nn.Dropout(p), SmallLinear(2 * L, L), nn.ReLU(), SmallLinear(L, L))
        self.final = SmallLinear(L, 1)
        self.dropout = nn.Dropout(p)

    def forward(self, data):
        _, _, _, _, _, user_id, item_id = data
        user = self.dropout(self.user_embedding(user_id.reshape(-1)))
        item = self.dropout(self.item_embedding(item_id.reshape(-1)))
        joint = self.project(torch.cat([user, item], dim=-1))
        rating = self.final(joint)[:, 0].view(tuple(user_id.shape))
        return self._bias_sum(user_id, item_id) + rating


class NeuMF(_Biases):
    """GMF and MLP side by side with separate embedding tables, read out jointly by ``final`` over the
    concatenation [gmf product, mlp features] (NeuMF.py:75-138).  ``init`` seeds it from a trained GMF and a
    trained MLP; ``main.main_NeuMF`` then trains it further with ``train_complete``."""

    def __init__(self, hyper_params):
        super().__init__(hyper_params)
        L, p = hyper_params["latent_size"], hyper_params["dropout"]
        U, I = hyper_params["total_users"] + 1, hyper_params["total_items"] + 1
        self.gmf_user_embedding = IdEmbedding(U, L)
        self.gmf_item_embedding = IdEmbedding(I, L)
        self.mlp_user_embedding = IdEmbedding(U, L)
        self.mlp_item_embedding = IdEmbedding(I, L)
        self.project = nn.Sequential(nn.Dropout(p), SmallLinear(2 * L, L), nn.ReLU(), SmallLinear(L, L))
        self.final = SmallLinear(2 * L, 1)
        self.dropout = nn.Dropout(p)

    def init(self, gmf_model, mlp_model):
        """NeuMF.py:93-112: start from the pre-trained GMF and MLP (embeddings copied, output layers concatenated,
        biases averaged; ``global_bias`` keeps its constructor value)."""
        with torch.no_grad():
            self.gmf_user_embedding.weight.data = gmf_model.user_embedding.weight.data
            self.gmf_item_embedding.weight.data = gmf_model.item_embedding.weight.data
            self.mlp_user_embedding.weight.data = mlp_model.user_embedding.weight.data
            self.mlp_item_embedding.weight.data = mlp_model.item_embedding.weight.data
            for i in (1, 3):                                    # the Linear layers of `project`
                self.project[i].weight.data = mlp_model.project[i].weight.data
                self.project[i].bias.data = mlp_model.project[i].bias.data
            self.final.weight.data = torch.cat([gmf_model.final.weight.data, mlp_model.final.weight.data], dim=-1)
            self.final.bias.data = 0.5 * (gmf_model.final.bias.data + mlp_model.final.bias.data)
            self.user_bias.data = 0.5 * (gmf_model.user_bias.data + mlp_model.user_bias.data)
            self.item_bias.data = 0.5 * (gmf_model.item_bias.data + mlp_model.item_bias.data)

    def forward(self, data):
        _, _, _, _, _, user_id, item_id = data
        uid, iid = user_id.reshape(-1), item_id.reshape(-1)
        gmf = self.dropout(self.gmf_user_embedding(uid)) * self.dropout(self.gmf_item_embedding(iid))
        mlp = self.project(torch.cat([self.dropout(self.mlp_user_embedding(uid)), self.dropout(self.mlp_item_embedding(iid))], dim=-1))
        rating = self.final(torch.cat([gmf, mlp], dim=-1))[:, 0].view(tuple(user_id.shape))
        return self._bias_sum(user_id, item_id) + rating
