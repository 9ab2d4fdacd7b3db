"""bias_only / MF_dot / MF with the reference's interface (pytorch_models/MF.py:9-68)."""
import torch
import torch.nn as nn

from .. import ops
from .common_pytorch_models import SmallLinear, TorchFM


class IdEmbedding(nn.Embedding):
    """nn.Embedding(sparse=False) whose lookup / dense gradient scatter run in r4r_rows_*."""

    def forward(self, ids):
        return ops.rows_gather(self.weight, ids)


class MF(nn.Module):
    """``bias_only`` / ``MF_dot`` / ``MF`` (MF.py:9-68): biases only; biases + dot product of the (dropped-out) id
    embeddings; or an MLP over the concatenated embeddings joined with their elementwise product through an FM.
    Tables have ``total_users + 1`` / ``total_items + 1`` rows (review models: + 2).  Every lookup is
    ``r4r_rows_gather`` and its backward the dense ``r4r_rows_scatter_add`` gradient that the reference's
    ``nn.Embedding(sparse=False)`` / ``Tensor.gather`` produce, so Adam updates every row (SURVEY.md finding 5)."""

    def __init__(self, hyper_params):
        super().__init__()
        self.hyper_params = hyper_params
        mt, L = hyper_params["model_type"], hyper_params["latent_size"]
        self.user_bias = nn.Parameter(torch.full((hyper_params["total_users"] + 1,), 0.1))
        self.item_bias = nn.Parameter(torch.full((hyper_params["total_items"] + 1,), 0.1))
        self.global_bias = nn.Parameter(torch.full((1,), 4.0))
        if mt in ("MF", "MF_dot"):
            self.user_embedding = IdEmbedding(hyper_params["total_users"] + 1, L)
            self.item_embedding = IdEmbedding(hyper_params["total_items"] + 1, L)
            self.dropout = nn.Dropout(hyper_params["dropout"])
        if mt == "MF":
            self.projection = nn.Sequential(nn.Dropout(hyper_params["dropout"]), SmallLinear(2 * L, L), nn.ReLU(), SmallLinear(L, L))
            self.final = TorchFM(2 * L, L)
            self.sigmoid = nn.Sigmoid()
            self.relu = nn.ReLU()

    def forward(self, data):
        _, _, _, _, _, user_id, item_id = data
        self.in_shape = user_id.shape
        shape = tuple(user_id.shape)
        ub = ops.rows_gather(self.user_bias, user_id.reshape(-1)).view(shape)
        ib = ops.rows_gather(self.item_bias, item_id.reshape(-1)).view(shape)
        mt = self.hyper_params["model_type"]
        if mt == "bias_only":
            return ub + ib + self.global_bias
        user = self.dropout(self.user_embedding(user_id.reshape(-1)))
        item = self.dropout(self.item_embedding(item_id.reshape(-1)))
        if mt == "MF_dot":
            return ub + ib + self.global_bias + torch.sum(user * item, dim=-1).view(shape)
        mlp = self.projection(torch.cat([user, item], dim=-1))
        rating = self.final(torch.cat([mlp, user * item], dim=-1))[:, 0].view(shape)
        return ub + ib + self.global_bias + rating
