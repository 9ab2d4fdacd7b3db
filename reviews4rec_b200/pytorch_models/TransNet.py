"""TransNet / TransNet++ with the reference's interface (pytorch_models/TransNet.py:9-122):
``.source`` / ``.target`` / ``.source_fm`` (+ id embeddings) are the optimizer groups of
utils.init_transnet_optim, and ``source.ir`` / ``target.ir`` are stashed on the sub-modules."""
import torch
import torch.nn as nn

from ..utils import load_obj, xavier_init
from .common_pytorch_models import SmallLinear, TextCNN, TorchFM, WordTable
from .MF import IdEmbedding


class Source(nn.Module):
    """Source network (TransNet.py:9-37): the user-document and item-document towers, concatenated and projected
    to the latent space.  Its output is not returned but left on ``self.ir`` -- the train step's transform loss
    reads it from there (main.py:35-53), and so does ``utils.init_transnet_optim``'s ``optimizer_source`` group.
    Each tower is one fused gather + conv + pool launch (the ``Docs`` handles carry token ids, not embeddings)."""

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
    """Target network (TransNet.py:39-61): owns the (frozen) word table shared by all three towers, embeds the
    review written for this very (user, item) pair and predicts the rating from it with an FM; ``self.ir`` is the
    representation the source network is trained to reproduce."""

    def __init__(self, hyper_params):
        super().__init__()
        self.hyper_params = hyper_params
        self.word2vec = WordTable.from_vectors(load_obj(hyper_params["data_dir"] + "/word2vec"),
                                                trainable=bool(hyper_params.get("train_word_table", False)))
        self.conv = TextCNN(hyper_params)
        self.dropout = nn.Dropout(hyper_params["dropout"])
        self.fm = TorchFM(hyper_params["latent_size"], 8)

    def embed(self, review):
        """Lazy lookup: returns a ``Docs`` handle (ids + table); the rows are gathered inside the conv kernel."""
        return self.word2vec(review)

    def forward(self, this):
        self.ir = self.dropout(self.conv(this))
        return self.fm(self.ir)


class TransNet(nn.Module):
    """TransNet / TransNet++ (TransNet.py:63-122).  ``forward`` returns the reference's 3-list
    ``[source rating [B], target rating [B], mean_b sum_l (source.ir - target.ir)^2]``; the ``++`` variant feeds
    5-dimensional user / item id embeddings next to ``source.ir`` into ``source_fm``.  Sub-module names, the
    per-sub-module xavier initialisation at construction (:68-72) and the state_dict layout are the contract
    ``utils.init_transnet_optim`` (utils.py:70-92) and the checkpoints depend on."""

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

    def word_inputs(self, data):
        """The token-id tensors ``forward`` hands to the word table: user document, item document, this review."""
        n = data[5].numel()
        return tuple(data[j].reshape(n, data[j].shape[-1]) for j in (3, 4, 0))

    def forward(self, data):
        this_reviews, _, _, user_reviews, item_reviews, user_id, item_id = data
        final_shape = tuple(user_id.shape)
        n = user_id.numel()
        user, item, this = self.target.word2vec.many(*self.word_inputs(data))
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
