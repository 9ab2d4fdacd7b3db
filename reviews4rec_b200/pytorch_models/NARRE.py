"""NARRE with the reference's interface (pytorch_models/NARRE.py:9-124)."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops
from ..utils import load_obj
from .common_pytorch_models import SmallLinear, TextCNN, WordTable
from .MF import IdEmbedding


class NARRE(nn.Module):
    """NARRE (NARRE.py:9-124): every one of the R reviews of a user (item) is its own conv document, the R review
    features are attended over with the id embedding of the item (user) each review is about, the attended sum is
    added to the user's (item's) own dropped-out id embedding, and the elementwise product of both sides goes
    through an MLP plus biases.  Shapes come from the data ([B, R, W] documents, [B, R] neighbour ids), not from
    ``narre_num_*``.  The neighbour pad ids ``total_users + 1`` / ``total_items + 1`` are hot rows of the id
    tables: their gradient scatter is combined per warp before the atomics."""

    def __init__(self, hyper_params):
        super().__init__()
        self.hyper_params = hyper_params
        L, p = hyper_params["latent_size"], hyper_params["dropout"]
        self.word2vec = WordTable.from_vectors(load_obj(hyper_params["data_dir"] + "/word2vec"),
                                                trainable=bool(hyper_params.get("train_word_table", False)))
        self.user_embedding = IdEmbedding(hyper_params["total_users"] + 2, L)
        self.item_embedding = IdEmbedding(hyper_params["total_items"] + 2, L)
        self.user_conv = TextCNN(hyper_params)
        self.item_conv = TextCNN(hyper_params)
        self.attention_scorer_user = nn.Sequential(SmallLinear(2 * L, L), nn.ReLU(), nn.Dropout(p), SmallLinear(L, 1))
        self.attention_scorer_item = nn.Sequential(SmallLinear(2 * L, L), nn.ReLU(), nn.Dropout(p), SmallLinear(L, 1))
        self.final = nn.Sequential(nn.Dropout(p), SmallLinear(L, L), nn.ReLU(), SmallLinear(L, 1))
        self.user_bias = nn.Parameter(torch.full((hyper_params["total_users"] + 2,), 0.1))
        self.item_bias = nn.Parameter(torch.full((hyper_params["total_items"] + 2,), 0.1))
        self.global_bias = nn.Parameter(torch.full((1,), 4.0))
        self.dropout = nn.Dropout(p)
        self.sigmoid = nn.Sigmoid()
        self.relu = nn.ReLU()

    def attention(self, x, other_x=None, scorer=None):
        """NARRE.py:53-64: softmax over the R reviews of an MLP score on [review feature, neighbour id emb]."""
        scores = scorer(torch.cat([x, other_x], dim=-1))[:, :, 0]
        return torch.sum(F.softmax(scores, dim=-1).unsqueeze(-1) * x, dim=1)

    def word_inputs(self, data):
        """The token-id tensors ``forward`` hands to the word table: every review is its own document [n*R, W]."""
        n = data[5].numel()
        ur, ir = data[3], data[4]
        return ur.reshape(n * ur.shape[-2], ur.shape[-1]), ir.reshape(n * ir.shape[-2], ir.shape[-1])

    def forward(self, data):
        _, users_who_reviewed, reviewed_items, user_reviews, item_reviews, user_id, item_id = data
        final_shape = tuple(user_id.shape)
        n = user_id.numel()
        R_u, W_u = user_reviews.shape[-2], user_reviews.shape[-1]
        R_i, W_i = item_reviews.shape[-2], item_reviews.shape[-1]
        users_who_reviewed = users_who_reviewed.reshape(n, users_who_reviewed.shape[-1])     # explicit widths: n may be 0
        reviewed_items = reviewed_items.reshape(n, reviewed_items.shape[-1])
        user_id, item_id = user_id.reshape(-1), item_id.reshape(-1)
        ub = ops.rows_gather(self.user_bias, user_id)
        ib = ops.rows_gather(self.item_bias, item_id)
        # every review is its own conv document: [n*R, W] token ids (NARRE.py:91-104)
        user_docs, item_docs = self.word2vec.many(*self.word_inputs(data))
        L = self.hyper_params["latent_size"]
        user = self.user_conv(user_docs).view(n, R_u, L)
        item = self.item_conv(item_docs).view(n, R_i, L)
        user = self.attention(user, self.item_embedding(reviewed_items), self.attention_scorer_user)
        item = self.attention(item, self.user_embedding(users_who_reviewed), self.attention_scorer_item)
        user = user + self.dropout(self.user_embedding(user_id))
        item = item + self.dropout(self.item_embedding(item_id))
        rating = self.final(user * item)[:, 0]
        return (rating + ub + ib + self.global_bias).view(final_shape)
