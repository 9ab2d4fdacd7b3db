"""TextCNN and TorchFM with the reference's interface (common_pytorch_models.py:6-57).

The nn.Conv2d / nn.Linear children are kept only as *parameter containers* (identical state_dict
keys, shapes and default initialisation); the arithmetic is the fused gather+conv+pool kernel, the
small-linear kernel and the FM kernel of libr4r_b200.
"""
import torch
import torch.nn as nn

from .. import ops


class Docs:
    """Un-materialised result of a word-embedding lookup: token ids [N,T] + the table to read."""

    __slots__ = ("idx", "table", "shadow")

    def __init__(self, idx, table, shadow=None):
        self.idx, self.table, self.shadow = idx, table, shadow


class WordTable(nn.Embedding):
    """Frozen word table (``nn.Embedding.from_pretrained`` default freeze=True, DeepCoNN.py:15).
    Calling it returns a lazy ``Docs`` handle; the gather is fused into the conv kernel."""

    @classmethod
    def from_vectors(cls, word_vectors, trainable=False):
        """``trainable=True`` (hyper_params['train_word_table']) is an opt-in extension, not reference
        behaviour: the reference's table is frozen (SURVEY.md finding 2 / 8f-3)."""
        w = torch.as_tensor(word_vectors, dtype=torch.float32)
        m = cls(w.shape[0], w.shape[1], _weight=w, _freeze=not trainable)
        m.requires_grad = False          # same inert attribute the reference sets (DeepCoNN.py:16)
        m._shadow = ops.ShadowTable()
        return m

    def forward(self, idx):
        return Docs(idx, self.weight, self._shadow)

    def many(self, *idx_list):
        """All word lookups of one step at once (a single exchange when the table is sharded)."""
        return tuple(self.forward(i) for i in idx_list)

    def materialize(self, idx):
        return ops.word_gather(self.weight, idx)


class TextCNN(nn.Module):
    def __init__(self, hyper_params, window_sizes=[3]):
        super().__init__()
        if list(window_sizes) != [3]:
            raise ValueError("the sm_100a conv kernel implements the reference's window size 3 only")
        self.hyper_params = hyper_params
        self.num_filters = 100
        self.convs = nn.ModuleList([
            nn.Conv2d(1, self.num_filters, [w, hyper_params["word_embed_size"]], padding=(w - 1, 0))
            for w in window_sizes])
        self.fc = nn.Linear(self.num_filters * len(window_sizes), hyper_params["latent_size"])
        self.dropout = nn.Dropout(hyper_params["dropout"])

    def pooled(self, x):
        """[N,100] relu+max-pooled conv features of ``x`` (a Docs handle or a dense [N,T,E] tensor)."""
        conv = self.convs[0]
        if isinstance(x, Docs):
            return ops.conv_pool(x.idx, x.table, conv.weight, conv.bias, shadow=x.shadow)
        if x.requires_grad:
            raise RuntimeError("gradients w.r.t. a dense [N,T,E] TextCNN input are not part of this path; "
                               "a trainable word table goes through the WordTable / Docs handle")
        n, t, e = x.shape                                   # reference signature: embedded docs [N,T,E]
        idx = torch.arange(n * t, device=x.device, dtype=torch.int64).view(n, t)
        # a fresh shadow per call: activations are not a frozen table (the caching allocator hands a new tensor
        # the old address and version, which is exactly ShadowTable's cache key)
        return ops.conv_pool(idx, x.reshape(n * t, e), conv.weight, conv.bias, shadow=ops.ShadowTable())

    def forward(self, x):
        return self.dropout(ops.linear(self.pooled(x), self.fc.weight, self.fc.bias))


class TorchFM(nn.Module):
    def __init__(self, n=None, k=None):
        super().__init__()
        self.V = nn.Parameter(torch.randn(n, k), requires_grad=True)
        self.lin = nn.Linear(n, 1)

    def forward(self, x):
        return ops.fm(x, self.V, self.lin.weight.view(-1), self.lin.bias)


class SmallLinear(nn.Linear):
    """nn.Linear whose forward/backward run in r4r_linear_fwd / r4r_linear_bwd."""

    def forward(self, x):
        return ops.linear(x, self.weight, self.bias)
