"""DeepCoNN / DeepCoNN++ with the reference's constructor, forward(data) and state_dict
(pytorch_models/DeepCoNN.py:9-72)."""
import torch
import torch.nn as nn

from .. import ops
from ..utils import load_obj
from .common_pytorch_models import SmallLinear, TextCNN, TorchFM, WordTable


class DeepCoNN(nn.Module):
    """``deepconn``: FM over the two TextCNN latents + global bias (DeepCoNN.py:64-66); ``deepconn++``: MLP over
    them + user / item / global biases (:69-72).  All parameters of both heads exist in either mode, exactly like
    the reference (its state_dict carries ``final.*``, ``fm.*`` and both bias vectors regardless), and the ones the
    active head does not touch never receive a gradient, so Adam never steps them.  The two word lookups go through
    ``word2vec.many`` (one exchange when the table is row-sharded) and each tower is a single fused
    gather + conv + ReLU + max-pool launch followed by the small FC."""

    def __init__(self, hyper_params):
        super().__init__()
        self.hyper_params = hyper_params
        L = hyper_params["latent_size"]
        self.word2vec = WordTable.from_vectors(load_obj(hyper_params["data_dir"] + "/word2vec"),
                                                trainable=bool(hyper_params.get("train_word_table", False)))
        self.user_conv = TextCNN(hyper_params)
        self.item_conv = TextCNN(hyper_params)
        self.final = nn.Sequential(SmallLinear(2 * L, L), nn.ReLU(), nn.Dropout(hyper_params["dropout"]), SmallLinear(L, 1))
        self.user_bias = nn.Parameter(torch.full((hyper_params["total_users"] + 2,), 0.1))
        self.item_bias = nn.Parameter(torch.full((hyper_params["total_items"] + 2,), 0.1))
        self.global_bias = nn.Parameter(torch.full((1,), 4.0))
        self.fm = TorchFM(2 * L, 8)
        self.dropout = nn.Dropout(hyper_params["dropout"])
        self.relu = nn.ReLU()

    def word_inputs(self, data):
        """The token-id tensors ``forward`` hands to the word table, in order (a sharded table can look them up
        a step ahead: train.CapturedStep(next_data=...))."""
        first_dim = data[5].numel()                             # explicit widths: a rank's slice of a batch may be empty
        return data[3].reshape(first_dim, data[3].shape[-1]), data[4].reshape(first_dim, data[4].shape[-1])

    fused_head = True        # one kernel for fc + dropout + cat + FM / MLP + biases (ops.deepconn_head); False = the op-by-op path

    def _head(self, data, y=None, se_sum=None):
        """(rating [first_dim], se or None) through the fused head kernel (ops.deepconn_head)."""
        user_id, item_id = data[5], data[6]
        user_docs, item_docs = self.word2vec.many(*self.word_inputs(data))
        pu, pi = self.user_conv.pooled(user_docs), self.item_conv.pooled(item_docs)
        hp = self.hyper_params
        p = float(hp["dropout"]) if self.training else 0.0
        if p > 0.0 and not hasattr(self, "_r4r_drop_step"):
            self._r4r_drop_step = torch.zeros(1, device=pu.device, dtype=torch.int32)      # Philox offset, advanced once per backward
            self._r4r_seed = torch.initial_seed()
        masks = getattr(self, "_r4r_keep_masks", None)          # tests: explicit keep masks [N, 3L] (uint8), used once
        if masks is not None:
            self._r4r_keep_masks = None
        kw = dict(y=y, p=p, seed=getattr(self, "_r4r_seed", 0), step=getattr(self, "_r4r_drop_step", None) if p > 0.0 else None,
                  masks=masks, se_sum=se_sum, global_bias=self.global_bias)
        fc_u = (self.user_conv.fc.weight, self.user_conv.fc.bias)
        fc_i = (self.item_conv.fc.weight, self.item_conv.fc.bias)
        if hp["model_type"] == "deepconn":
            return ops.deepconn_head(pu, pi, fc_u, fc_i, 0, fm=(self.fm.V, self.fm.lin.weight, self.fm.lin.bias), **kw)
        ub = ops.rows_gather(self.user_bias, user_id.reshape(-1))
        ib = ops.rows_gather(self.item_bias, item_id.reshape(-1))
        final = (self.final[0].weight, self.final[0].bias, self.final[3].weight, self.final[3].bias)
        return ops.deepconn_head(pu, pi, fc_u, fc_i, 1, final=final, ub=ub, ib=ib, **kw)

    def _fused_ok(self):
        L = self.hyper_params["latent_size"]
        return self.fused_head and ops.deepconn_head_supported(L, self.user_conv.num_filters, self.fm.V.shape[1])

    def forward_with_loss(self, data, y, se_sum=None):
        """(rating, per-sample squared error) with the loss fused into the head kernel (loss.py:7-11); ``se_sum`` (device
        scalar) additionally receives the batch's sum.  train.CapturedStep uses it; ``forward`` + ``MSELoss`` give the same numbers."""
        if not self._fused_ok():
            out = self.forward(data)
            se = ops.squared_error(out, y)
            if se_sum is not None:
                se_sum += se.detach().sum()
            return out, se
        rating, se = self._head(data, y=y.reshape(-1), se_sum=se_sum)
        shape = tuple(data[5].shape)
        return rating.view(shape), se.view(shape)

    def forward(self, data):
        _, _, _, user_reviews, item_reviews, user_id, item_id = data
        final_shape = tuple(user_id.shape)                      # [B] or [B,n] (ranking, eval.py:64-92)
        if self._fused_ok():
            return self._head(data)[0].view(final_shape)
        user_docs, item_docs = self.word2vec.many(*self.word_inputs(data))
        user = self.user_conv(user_docs)
        item = self.item_conv(item_docs)
        cat = torch.cat([user, item], dim=-1)
        if self.hyper_params["model_type"] == "deepconn":
            return (self.global_bias + self.fm(cat)[:, 0]).view(final_shape)
        rating = self.final(cat)[:, 0]
        ub = ops.rows_gather(self.user_bias, user_id.reshape(-1))
        ib = ops.rows_gather(self.item_bias, item_id.reshape(-1))
        return (rating + ub + ib + self.global_bias).view(final_shape)
