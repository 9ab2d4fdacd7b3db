"""TransNet / TransNet++ with the reference's interface (pytorch_models/TransNet.py:9-122):
``.source`` / ``.target`` / ``.source_fm`` (+ id embeddings) are the optimizer groups of
utils.init_transnet_optim, and ``source.ir`` / ``target.ir`` are stashed on the sub-modules."""
import torch
import torch.nn as nn

from ..utils import load_obj, xavier_init
from .common_pytorch_models import SmallLinear, TextCNN, TorchFM, WordTable
from .MF import IdEmbedding


class Source(nn.Module):
    def __init__(self, hyper_params):
        super().__init__()
        self.hyper_params = hyper_params
        L = hyper_params["latent_size"]
        self.user_conv = TextCNN(hyper_params)
        self.item_conv = TextCNN(hyper_params)
        self.project = nn.Sequential(SmallLinear(2 * L, L), nn.ReLU(), SmallLinear(L, L))
        self.dropout = nn.Dropout(hyper_params["dropout"])

    def forward(self, user, item):
        cat = torch.cat([self.user_conv(user), self.item_conv(item)], dim=-1)
        self.ir = self.dropout(self.project(cat))
        return None


class Target(nn.Module):
    def __init__(self, hyper_params):
        super().__init__()
        self.hyper_params = hyper_params
        self.word2vec = WordTable.from_vectors(load_obj(hyper_params["data_dir"] + "/word2vec"),
                                                trainable=bool(hyper_params.get("train_word_table", False)))
        self.conv = TextCNN(hyper_params)
        self.dropout = nn.Dropout(hyper_params["dropout"])
        self.fm = TorchFM(hyper_params["latent_size"], 8)

    def embed(self, review):
        return self.word2vec(review)

    def forward(self, this):
        self.ir = self.dropout(self.conv(this))
        return self.fm(self.ir)


class TransNet(nn.Module):
    def __init__(self, hyper_params):
        super().__init__()
        self.hyper_params = hyper_params
        L = hyper_params["latent_size"]
        self.target = Target(hyper_params)
        xavier_init(self.target)
        self.source = Source(hyper_params)
        xavier_init(self.source)
        if hyper_params["model_type"] == "transnet++":
            self.user_embedding = IdEmbedding(hyper_params["total_users"] + 2, 5)
            self.item_embedding = IdEmbedding(hyper_params["total_items"] + 2, 5)
            self.source_fm = TorchFM(10 + L, 8)
        else:
            self.source_fm = TorchFM(L, 8)
        self.dropout = nn.Dropout(hyper_params["dropout"])

    def forward(self, data):
        this_reviews, _, _, user_reviews, item_reviews, user_id, item_id = data
        final_shape = tuple(user_id.shape)
        n = user_id.numel()
        user, item, this = self.target.word2vec.many(user_reviews.reshape(n, -1), item_reviews.reshape(n, -1),
                                                     this_reviews.reshape(n, -1))
        self.source(user, item)
        if self.hyper_params["model_type"] == "transnet++":
            u = self.dropout(self.user_embedding(user_id.reshape(-1)))
            i = self.dropout(self.item_embedding(item_id.reshape(-1)))
            final = torch.cat([u, i, self.source.ir], dim=-1)
        else:
            final = self.source.ir
        source_out = self.source_fm(final)
        target_out = self.target(this)
        return [source_out[:, 0].view(final_shape), target_out[:, 0].view(final_shape),
                torch.mean(torch.sum(torch.pow(self.source.ir - self.target.ir, 2), -1))]
