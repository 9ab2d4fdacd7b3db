"""CPU oracle for the rating-prediction training hot path  --  TEST INFRASTRUCTURE ONLY.

This file is the checker, never the product: only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it.  Nothing under
``reviews4rec_b200/`` imports it and the product path fails loudly when its CUDA library is absent.

It is a *functional* restatement (plain functions over a ``{state_dict key: tensor}`` dict, torch CPU
ops, fp32 by default, fp64 on request) of what the reference's ``nn.Module`` classes compute:

  =====================  ==========================================================================
  oracle function        reference lines it follows
  =====================  ==========================================================================
  ``mse``                ``loss.py:7-11``
  ``text_cnn``           ``pytorch_models/common_pytorch_models.py:22-39`` (ctor ``:7-20``)
  ``torch_fm``           ``pytorch_models/common_pytorch_models.py:49-57``
  ``deepconn_forward``   ``pytorch_models/DeepCoNN.py:37-72``
  ``mf_forward``         ``pytorch_models/MF.py:39-68``
  ``neumf_forward``      ``pytorch_models/NeuMF.py:23-36,59-73,114-138`` (GMF / MLP / NeuMF), ``neumf_init`` ``:93-112``
  ``narre_forward``      ``pytorch_models/NARRE.py:53-124``
  ``transnet_forward``   ``pytorch_models/TransNet.py:25-37,55-61,83-122``
  ``adam_step``          ``torch.optim.Adam`` as configured at ``main.py:94-96`` / ``utils.py:70-92``
  ``train_batches``      ``main.py:8-71`` (non-TransNet branch ``:55-60``)
  ``transnet_train``     ``main.py:35-53`` restated per SURVEY.md section 8(c): the reference's own
                         loop raises on torch >= 1.5, so three ``autograd.grad`` calls on one graph
                         followed by the three optimizer steps reproduce old-torch behaviour.
  =====================  ==========================================================================

Parity pinning: the reference ships no golden vectors, so this oracle is pinned
against *outputs of the reference itself*: ``oracle/gen_golden.py`` imports the unmodified modules from
``/root/reference`` in the build container, runs them on seeded inputs and commits the results under
``tests/golden/``; ``tests/test_oracle_golden.py`` replays them through this file.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F

Params = Dict[str, torch.Tensor]

NUM_FILTERS = 100      # common_pytorch_models.py:11 (hard-coded)
WINDOW = 3             # common_pytorch_models.py:7 default window_sizes=[3]
FM_K = 8               # DeepCoNN.py:32, TransNet.py:50,77-79


# ----------------------------------------------------------------------------- small pieces
def mse(output: torch.Tensor, y: torch.Tensor, return_mean: bool = True) -> torch.Tensor:
    """loss.py:7-11."""
    se = (output - y) ** 2
    return se.mean() if return_mean else se


def _drop(x: torch.Tensor, p: float, train: bool, mask: Optional[torch.Tensor]) -> torch.Tensor:
    """nn.Dropout.  ``mask`` (0/1 keep mask, same shape) makes it deterministic for parity runs."""
    if not train or p == 0.0:
        return x
    if mask is not None:
        return x * mask.to(x.dtype) / (1.0 - p)
    return F.dropout(x, p, True)


def word_gather(table: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """nn.Embedding forward (DeepCoNN.py:53-54): out[..., :] = table[idx[...], :]."""
    return table.index_select(0, idx.reshape(-1)).reshape(*idx.shape, table.shape[1])


def conv_pool(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor):
    """common_pytorch_models.py:26-31: conv2d(pad=(w-1,0)) + relu + global max-pool.

    x [N,T,E]; w [F,1,3,E]; returns (pooled [N,F], argmax [N,F] first-max position in 0..T+1)."""
    y = F.conv2d(x.unsqueeze(1), w, b, padding=(w.shape[2] - 1, 0)).squeeze(-1)   # [N,F,T+2]
    y = F.relu(y)
    pooled, arg = F.max_pool1d(y, y.shape[2], return_indices=True)
    return pooled.squeeze(-1), arg.squeeze(-1)


def text_cnn(P: Params, prefix: str, x: torch.Tensor, p: float, train: bool,
             mask: Optional[torch.Tensor] = None) -> torch.Tensor:
    """TextCNN.forward, common_pytorch_models.py:22-39."""
    pooled, _ = conv_pool(x, P[prefix + "convs.0.weight"], P[prefix + "convs.0.bias"])
    lat = F.linear(pooled, P[prefix + "fc.weight"], P[prefix + "fc.bias"])
    return _drop(lat, p, train, mask)


def torch_fm(x: torch.Tensor, V: torch.Tensor, lin_w: torch.Tensor, lin_b: torch.Tensor) -> torch.Tensor:
    """TorchFM.forward, common_pytorch_models.py:49-57.  Returns [B,1]."""
    s1 = (x @ V).pow(2).sum(1, keepdim=True)
    s2 = (x.pow(2) @ V.pow(2)).sum(1, keepdim=True)
    return 0.5 * (s1 - s2) + F.linear(x, lin_w, lin_b)


def _flatten_ids(user_id: torch.Tensor):
    """DeepCoNN.py:40-50 / NARRE.py:69-81 / TransNet.py:86-97: optional candidates dim."""
    if user_id.dim() > 1:
        return (user_id.shape[0], user_id.shape[1]), user_id.shape[0] * user_id.shape[1]
    return (user_id.shape[0],), user_id.shape[0]


# ----------------------------------------------------------------------------- models
def deepconn_forward(P: Params, data: Sequence, hp: dict, train: bool = False,
                     masks: Optional[dict] = None) -> torch.Tensor:
    """DeepCoNN.forward, DeepCoNN.py:37-72.  ``masks`` keys: 'user_conv','item_conv','final'."""
    masks = masks or {}
    _, _, _, user_reviews, item_reviews, user_id, item_id = data
    final_shape, first = _flatten_ids(user_id)
    user_reviews = user_reviews.reshape(first, -1)
    item_reviews = item_reviews.reshape(first, -1)
    user_id, item_id = user_id.reshape(-1), item_id.reshape(-1)
    p = hp["dropout"]
    u = text_cnn(P, "user_conv.", word_gather(P["word2vec.weight"], user_reviews), p, train, masks.get("user_conv"))
    i = text_cnn(P, "item_conv.", word_gather(P["word2vec.weight"], item_reviews), p, train, masks.get("item_conv"))
    cat = torch.cat([u, i], dim=-1)
    if hp["model_type"] == "deepconn":
        rating = P["global_bias"] + torch_fm(cat, P["fm.V"], P["fm.lin.weight"], P["fm.lin.bias"])[:, 0]
        return rating.reshape(final_shape)
    h = F.relu(F.linear(cat, P["final.0.weight"], P["final.0.bias"]))
    h = _drop(h, p, train, masks.get("final"))
    rating = F.linear(h, P["final.3.weight"], P["final.3.bias"])[:, 0]
    ub = P["user_bias"].gather(0, user_id)
    ib = P["item_bias"].gather(0, item_id)
    return (rating + ub + ib + P["global_bias"]).reshape(final_shape)


def mf_forward(P: Params, data: Sequence, hp: dict, train: bool = False,
               masks: Optional[dict] = None) -> torch.Tensor:
    """MF.forward, MF.py:39-68.  ``masks`` keys: 'user','item','projection'."""
    masks = masks or {}
    user_id, item_id = data[5], data[6]
    shape = user_id.shape
    ub = P["user_bias"].gather(0, user_id.reshape(-1)).reshape(shape)
    ib = P["item_bias"].gather(0, item_id.reshape(-1)).reshape(shape)
    if hp["model_type"] == "bias_only":
        return ub + ib + P["global_bias"]
    p = hp["dropout"]
    u = _drop(P["user_embedding.weight"].index_select(0, user_id.reshape(-1)), p, train, masks.get("user"))
    i = _drop(P["item_embedding.weight"].index_select(0, item_id.reshape(-1)), p, train, masks.get("item"))
    if hp["model_type"] == "MF_dot":
        return ub + ib + P["global_bias"] + (u * i).sum(-1).reshape(shape)
    cat = _drop(torch.cat([u, i], dim=-1), p, train, masks.get("projection"))
    h = F.relu(F.linear(cat, P["projection.1.weight"], P["projection.1.bias"]))
    mlp = F.linear(h, P["projection.3.weight"], P["projection.3.bias"])
    cat2 = torch.cat([mlp, u * i], dim=-1)
    rating = torch_fm(cat2, P["final.V"], P["final.lin.weight"], P["final.lin.bias"])[:, 0].reshape(shape)
    return ub + ib + P["global_bias"] + rating


def _narre_attention(P: Params, scorer: str, x: torch.Tensor, other: torch.Tensor, p: float, train: bool,
                     mask: Optional[torch.Tensor]) -> torch.Tensor:
    """NARRE.attention, NARRE.py:53-64."""
    h = F.relu(F.linear(torch.cat([x, other], dim=-1), P[scorer + "0.weight"], P[scorer + "0.bias"]))
    h = _drop(h, p, train, mask)
    scores = F.linear(h, P[scorer + "3.weight"], P[scorer + "3.bias"])[:, :, 0]
    a = F.softmax(scores, dim=-1)
    return (a.unsqueeze(-1) * x).sum(1)


def narre_forward(P: Params, data: Sequence, hp: dict, train: bool = False,
                  masks: Optional[dict] = None) -> torch.Tensor:
    """NARRE.forward, NARRE.py:66-124.
    ``masks`` keys: 'user_conv','item_conv','att_user','att_item','user_id','item_id','final'."""
    masks = masks or {}
    _, users_who, items_rev, user_reviews, item_reviews, user_id, item_id = data
    final_shape, first = _flatten_ids(user_id)
    users_who = users_who.reshape(first, -1)
    items_rev = items_rev.reshape(first, -1)
    user_reviews = user_reviews.reshape(first, user_reviews.shape[-2], user_reviews.shape[-1])
    item_reviews = item_reviews.reshape(first, item_reviews.shape[-2], item_reviews.shape[-1])
    user_id, item_id = user_id.reshape(-1), item_id.reshape(-1)
    p = hp["dropout"]
    ub = P["user_bias"].gather(0, user_id)
    ib = P["item_bias"].gather(0, item_id)
    Bn, R, W = user_reviews.shape
    Bn2, R2, W2 = item_reviews.shape
    ux = word_gather(P["word2vec.weight"], user_reviews.reshape(Bn * R, W))
    ix = word_gather(P["word2vec.weight"], item_reviews.reshape(Bn2 * R2, W2))
    u = text_cnn(P, "user_conv.", ux, p, train, masks.get("user_conv")).reshape(Bn, R, -1)
    i = text_cnn(P, "item_conv.", ix, p, train, masks.get("item_conv")).reshape(Bn2, R2, -1)
    u = _narre_attention(P, "attention_scorer_user.", u, P["item_embedding.weight"][items_rev], p, train,
                         masks.get("att_user"))
    i = _narre_attention(P, "attention_scorer_item.", i, P["user_embedding.weight"][users_who], p, train,
                         masks.get("att_item"))
    u = u + _drop(P["user_embedding.weight"][user_id], p, train, masks.get("user_id"))
    i = i + _drop(P["item_embedding.weight"][item_id], p, train, masks.get("item_id"))
    cat = _drop(u * i, p, train, masks.get("final"))
    h = F.relu(F.linear(cat, P["final.1.weight"], P["final.1.bias"]))
    rating = F.linear(h, P["final.3.weight"], P["final.3.bias"])[:, 0]
    return (rating + ub + ib + P["global_bias"]).reshape(final_shape)


def transnet_forward(P: Params, data: Sequence, hp: dict, train: bool = False,
                     masks: Optional[dict] = None) -> List[torch.Tensor]:
    """TransNet.forward with Source/Target, TransNet.py:25-37,55-61,83-122.

    Returns [source rating, target rating, mean_b sum_l (source.ir - target.ir)^2] and stashes
    the two intermediate representations in ``transnet_forward.ir`` for tests.
    ``masks`` keys: 's_user_conv','s_item_conv','source','t_conv','target','user_id','item_id'."""
    masks = masks or {}
    this_reviews, _, _, user_reviews, item_reviews, user_id, item_id = data
    final_shape, first = _flatten_ids(user_id)
    this_reviews = this_reviews.reshape(first, -1)
    user_reviews = user_reviews.reshape(first, -1)
    item_reviews = item_reviews.reshape(first, -1)
    user_id, item_id = user_id.reshape(-1), item_id.reshape(-1)
    p = hp["dropout"]
    tab = P["target.word2vec.weight"]
    u = text_cnn(P, "source.user_conv.", word_gather(tab, user_reviews), p, train, masks.get("s_user_conv"))
    i = text_cnn(P, "source.item_conv.", word_gather(tab, item_reviews), p, train, masks.get("s_item_conv"))
    h = F.relu(F.linear(torch.cat([u, i], dim=-1), P["source.project.0.weight"], P["source.project.0.bias"]))
    s_ir = _drop(F.linear(h, P["source.project.2.weight"], P["source.project.2.bias"]), p, train, masks.get("source"))
    if hp["model_type"] == "transnet++":
        ue = _drop(P["user_embedding.weight"][user_id], p, train, masks.get("user_id"))
        ie = _drop(P["item_embedding.weight"][item_id], p, train, masks.get("item_id"))
        fin = torch.cat([ue, ie, s_ir], dim=-1)
    else:
        fin = s_ir
    src = torch_fm(fin, P["source_fm.V"], P["source_fm.lin.weight"], P["source_fm.lin.bias"])
    t = text_cnn(P, "target.conv.", word_gather(tab, this_reviews), p, train, masks.get("t_conv"))
    t_ir = _drop(t, p, train, masks.get("target"))
    tgt = torch_fm(t_ir, P["target.fm.V"], P["target.fm.lin.weight"], P["target.fm.lin.bias"])
    transnet_forward.ir = (s_ir, t_ir)
    return [src[:, 0].reshape(final_shape), tgt[:, 0].reshape(final_shape),
            ((s_ir - t_ir) ** 2).sum(-1).mean()]


def neumf_forward(P: Params, data: Sequence, hp: dict, train: bool = False,
                  masks: Optional[dict] = None) -> torch.Tensor:
    """GMF / MLP / NeuMF forward, pytorch_models/NeuMF.py:23-36, :59-73, :114-138 (``model_type`` 'GMF', 'MLP',
    'NeuMF').  ``masks`` keys: 'user','item' (GMF part), 'mlp_user','mlp_item','project'."""
    masks = masks or {}
    mt, p = hp["model_type"], hp["dropout"]
    user_id, item_id = data[5], data[6]
    shape = user_id.shape
    uid, iid = user_id.reshape(-1), item_id.reshape(-1)
    bias = P["user_bias"].gather(0, uid).reshape(shape) + P["item_bias"].gather(0, iid).reshape(shape) + P["global_bias"]
    emb = lambda k, ids, m: _drop(P[k + ".weight"].index_select(0, ids), p, train, masks.get(m))

    def project(x):
        x = _drop(x, p, train, masks.get("project"))
        return F.linear(F.relu(F.linear(x, P["project.1.weight"], P["project.1.bias"])), P["project.3.weight"], P["project.3.bias"])

    if mt == "GMF":
        joint = emb("user_embedding", uid, "user") * emb("item_embedding", iid, "item")
    elif mt == "MLP":
        joint = project(torch.cat([emb("user_embedding", uid, "user"), emb("item_embedding", iid, "item")], dim=-1))
    else:
        gmf = emb("gmf_user_embedding", uid, "user") * emb("gmf_item_embedding", iid, "item")
        mlp = project(torch.cat([emb("mlp_user_embedding", uid, "mlp_user"), emb("mlp_item_embedding", iid, "mlp_item")], dim=-1))
        joint = torch.cat([gmf, mlp], dim=-1)
    rating = F.linear(joint, P["final.weight"], P["final.bias"])[:, 0].reshape(shape)
    return bias + rating


def neumf_init(gmf: Params, mlp: Params, neumf: Params) -> Params:
    """NeuMF.init, pytorch_models/NeuMF.py:93-112."""
    out = {k: v.clone() for k, v in neumf.items()}
    out["gmf_user_embedding.weight"] = gmf["user_embedding.weight"].clone()
    out["gmf_item_embedding.weight"] = gmf["item_embedding.weight"].clone()
    out["mlp_user_embedding.weight"] = mlp["user_embedding.weight"].clone()
    out["mlp_item_embedding.weight"] = mlp["item_embedding.weight"].clone()
    for k in ("project.1.weight", "project.1.bias", "project.3.weight", "project.3.bias"):
        out[k] = mlp[k].clone()
    out["final.weight"] = torch.cat([gmf["final.weight"], mlp["final.weight"]], dim=-1)
    out["final.bias"] = 0.5 * (gmf["final.bias"] + mlp["final.bias"])
    out["user_bias"] = 0.5 * (gmf["user_bias"] + mlp["user_bias"])
    out["item_bias"] = 0.5 * (gmf["item_bias"] + mlp["item_bias"])
    return out


def forward(P: Params, data: Sequence, hp: dict, train: bool = False, masks: Optional[dict] = None):
    mt = hp["model_type"]
    if mt in ("GMF", "MLP", "NeuMF"):
        return neumf_forward(P, data, hp, train, masks)
    if mt in ("deepconn", "deepconn++"):
        return deepconn_forward(P, data, hp, train, masks)
    if mt in ("bias_only", "MF", "MF_dot"):
        return mf_forward(P, data, hp, train, masks)
    if mt == "NARRE":
        return narre_forward(P, data, hp, train, masks)
    if mt in ("transnet", "transnet++"):
        return transnet_forward(P, data, hp, train, masks)
    raise ValueError("unknown model_type %r" % (mt,))


# ----------------------------------------------------------------------------- parameters
def frozen_keys(P: Params) -> List[str]:
    """word2vec is frozen: nn.Embedding.from_pretrained default freeze=True (DeepCoNN.py:15)."""
    return [k for k in P if k.endswith("word2vec.weight")]


def init_params(hp: dict, V: int, seed: int = 0, dtype=torch.float32) -> Params:
    """Random parameters with the reference's state_dict keys/shapes and
    its effective initialisation: xavier-uniform on >=2-D params incl. word2vec (utils.py:65-68,
    main.py:377), biases 0.1 / 4.0 (DeepCoNN.py:28-30), nn.Linear default for 1-D biases."""
    g = torch.Generator().manual_seed(seed)
    mt, L, E = hp["model_type"], hp["latent_size"], hp.get("word_embed_size", 0)
    U, I = hp["total_users"], hp["total_items"]
    P: Params = {}

    def xav(*shape):
        t = torch.empty(*shape, dtype=dtype)
        if len(shape) == 2:
            fan_out, fan_in = shape
        else:
            rf = math.prod(shape[2:])
            fan_out, fan_in = shape[0] * rf, shape[1] * rf
        a = math.sqrt(6.0 / (fan_in + fan_out))
        return t.uniform_(-a, a, generator=g)

    def lin(name, out_f, in_f):
        P[name + ".weight"] = xav(out_f, in_f)
        P[name + ".bias"] = torch.empty(out_f, dtype=dtype).uniform_(-1 / math.sqrt(in_f), 1 / math.sqrt(in_f), generator=g)

    def tcnn(prefix):
        P[prefix + "convs.0.weight"] = xav(NUM_FILTERS, 1, WINDOW, E)
        P[prefix + "convs.0.bias"] = torch.empty(NUM_FILTERS, dtype=dtype).uniform_(-0.05, 0.05, generator=g)
        lin(prefix + "fc", L, NUM_FILTERS)

    def fm(prefix, n, k):
        P[prefix + "V"] = xav(n, k)
        lin(prefix + "lin", 1, n)

    if mt in ("deepconn", "deepconn++"):
        P["user_bias"] = torch.full((U + 2,), 0.1, dtype=dtype)
        P["item_bias"] = torch.full((I + 2,), 0.1, dtype=dtype)
        P["global_bias"] = torch.full((1,), 4.0, dtype=dtype)
        P["word2vec.weight"] = xav(V, E)
        tcnn("user_conv."); tcnn("item_conv.")
        lin("final.0", L, 2 * L); lin("final.3", 1, L)
        fm("fm.", 2 * L, FM_K)
    elif mt in ("bias_only", "MF", "MF_dot"):
        P["user_bias"] = torch.full((U + 1,), 0.1, dtype=dtype)
        P["item_bias"] = torch.full((I + 1,), 0.1, dtype=dtype)
        P["global_bias"] = torch.full((1,), 4.0, dtype=dtype)
        if mt != "bias_only":
            P["user_embedding.weight"] = xav(U + 1, L)
            P["item_embedding.weight"] = xav(I + 1, L)
        if mt == "MF":
            lin("projection.1", L, 2 * L); lin("projection.3", L, L)
            fm("final.", 2 * L, L)
    elif mt == "NARRE":
        P["user_bias"] = torch.full((U + 2,), 0.1, dtype=dtype)
        P["item_bias"] = torch.full((I + 2,), 0.1, dtype=dtype)
        P["global_bias"] = torch.full((1,), 4.0, dtype=dtype)
        P["word2vec.weight"] = xav(V, E)
        P["user_embedding.weight"] = xav(U + 2, L)
        P["item_embedding.weight"] = xav(I + 2, L)
        tcnn("user_conv."); tcnn("item_conv.")
        for s in ("attention_scorer_user", "attention_scorer_item"):
            lin(s + ".0", L, 2 * L); lin(s + ".3", 1, L)
        lin("final.1", L, L); lin("final.3", 1, L)
    elif mt in ("transnet", "transnet++"):
        P["target.word2vec.weight"] = xav(V, E)
        tcnn("target.conv."); fm("target.fm.", L, FM_K)
        tcnn("source.user_conv."); tcnn("source.item_conv.")
        lin("source.project.0", L, 2 * L); lin("source.project.2", L, L)
        if mt == "transnet++":
            P["user_embedding.weight"] = xav(U + 2, 5)
            P["item_embedding.weight"] = xav(I + 2, 5)
            fm("source_fm.", 10 + L, FM_K)
        else:
            fm("source_fm.", L, FM_K)
    else:
        raise ValueError(mt)
    return P


# ----------------------------------------------------------------------------- optimiser
def adam_step(p: torch.Tensor, g: torch.Tensor, m: torch.Tensor, v: torch.Tensor, step: int,
              lr: float, wd: float, b1: float = 0.9, b2: float = 0.999, eps: float = 1e-8) -> None:
    """One torch.optim.Adam update (L2 weight decay folded into the gradient, bias-corrected,
    ``denom = sqrt(v)/sqrt(1-b2^t) + eps``), in place; ``step`` is the 1-based step count."""
    if wd != 0.0:
        g = g + wd * p
    m.mul_(b1).add_(g, alpha=1 - b1)
    v.mul_(b2).addcmul_(g, g, value=1 - b2)
    bc1 = 1 - b1 ** step
    bc2 = 1 - b2 ** step
    denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
    p.addcdiv_(m, denom, value=-lr / bc1)


class AdamState:
    """Adam over a subset of ``P``; params whose grad is None are skipped like torch does."""

    def __init__(self, P: Params, keys: Sequence[str], lr: float, wd: float):
        self.P, self.keys, self.lr, self.wd = P, list(keys), lr, wd
        self.m = {k: torch.zeros_like(P[k]) for k in self.keys}
        self.v = {k: torch.zeros_like(P[k]) for k in self.keys}
        self.t = {k: 0 for k in self.keys}

    def step(self, grads: Dict[str, Optional[torch.Tensor]]) -> None:
        for k in self.keys:
            g = grads.get(k)
            if g is None:
                continue
            self.t[k] += 1
            adam_step(self.P[k], g, self.m[k], self.v[k], self.t[k], self.lr, self.wd)


def _leafify(P: Params) -> Params:
    fz = set(frozen_keys(P))
    return {k: (v.detach() if k in fz else v.detach().requires_grad_(True)) for k, v in P.items()}


def grads_of(P: Params, data: Sequence, y: torch.Tensor, hp: dict, train: bool = True,
             masks: Optional[dict] = None):
    """Forward + ``mean((out-y)^2).backward()`` (main.py:56-59).  Returns (out, per-sample SE, grads)."""
    Q = _leafify(P)
    out = forward(Q, data, hp, train, masks)
    se = mse(out, y, return_mean=False)
    se.mean().backward()
    grads = {k: (q.grad if q.requires_grad else None) for k, q in Q.items()}
    return out.detach(), se.detach(), grads


def train_batches(P: Params, batches, hp: dict, masks_per_batch=None, opt: Optional[AdamState] = None):
    """main.train() non-TransNet branch (main.py:23-66) with Adam as built at main.py:94-96.
    Mutates ``P`` in place.  Returns (metrics dict with the reference's rounding, raw SE sum, N, opt)."""
    if opt is None:
        opt = AdamState(P, [k for k in P if k not in frozen_keys(P)], hp["lr"], hp["weight_decay"])
    total, n = 0.0, 0
    for bi, (data, y) in enumerate(batches):
        masks = masks_per_batch[bi] if masks_per_batch is not None else None
        out, se, grads = grads_of(P, data, y, hp, True, masks)
        total += float(se.sum())
        n += int(out.shape[0])
        opt.step(grads)
    return {"MSE": round(total / float(n), 4)}, total, n, opt


def transnet_groups(P: Params, hp: dict):
    """Parameter groups of utils.init_transnet_optim (utils.py:70-92)."""
    fz = set(frozen_keys(P))
    src = [k for k in P if k.startswith("source.")]
    sfm = [k for k in P if k.startswith("source_fm.")]
    if hp["model_type"] == "transnet++":
        sfm += ["user_embedding.weight", "item_embedding.weight"]
    tgt = [k for k in P if k.startswith("target.") and k not in fz]
    return src, sfm, tgt


def transnet_grads(P: Params, data: Sequence, y: torch.Tensor, hp: dict, train: bool = True,
                   masks: Optional[dict] = None):
    """The three gradient sets of the restated TransNet step."""
    Q = _leafify(P)
    src, sfm, tgt = transnet_groups(P, hp)
    out = transnet_forward(Q, data, hp, train, masks)
    loss_t = mse(out[1], y)
    se_s = mse(out[0], y, return_mean=False)
    g_t = torch.autograd.grad(loss_t, [Q[k] for k in tgt], retain_graph=True, allow_unused=True)
    g_s = torch.autograd.grad(out[2], [Q[k] for k in src], retain_graph=True, allow_unused=True)
    g_f = torch.autograd.grad(se_s.mean(), [Q[k] for k in sfm], allow_unused=True)
    return ([o.detach() for o in out], se_s.detach(), loss_t.detach(),
            dict(zip(tgt, g_t)), dict(zip(src, g_s)), dict(zip(sfm, g_f)))


def transnet_train(P: Params, batches, hp: dict, masks_per_batch=None, opts=None):
    """main.train() TransNet branch (main.py:35-53,66-69), restated."""
    src, sfm, tgt = transnet_groups(P, hp)
    if opts is None:
        opts = (AdamState(P, src, hp["lr"], hp["weight_decay"]),
                AdamState(P, sfm, hp["lr"], hp["weight_decay"]),
                AdamState(P, tgt, hp["lr"], hp["weight_decay"]))
    o_src, o_sfm, o_tgt = opts
    tot, tot_t, tot_x, n, nb = 0.0, 0.0, 0.0, 0, 0
    for bi, (data, y) in enumerate(batches):
        masks = masks_per_batch[bi] if masks_per_batch is not None else None
        out, se_s, loss_t, g_t, g_s, g_f = transnet_grads(P, data, y, hp, True, masks)
        o_tgt.step(g_t)
        o_src.step(g_s)
        o_sfm.step(g_f)
        tot += float(se_s.sum()); tot_t += float(loss_t); tot_x += float(out[2])
        n += int(out[0].shape[0]); nb += 1
    metrics = {"MSE": round(tot / float(n), 4), "MSE_target": round(tot_t / float(nb), 4),
               "MSE_transform": round(tot_x / float(nb), 4)}
    return metrics, (tot, tot_t, tot_x), n, opts


# ----------------------------------------------------------------------------- closed forms used by kernels
def conv_wgrad_argmax(x: torch.Tensor, arg: torch.Tensor, g_pool: torch.Tensor, pooled: torch.Tensor):
    """Sparse weight gradient through relu + global max-pool (SURVEY.md finding 4):
    dW[f,j,:] = sum_n g[n,f]*[pooled>0] * Xpad[n, arg[n,f]+j, :],  db[f] = sum_n g[n,f]*[pooled>0].
    Equal to autograd of ``conv_pool`` (checked in tests).  x [N,T,E] -> (dW [F,1,3,E], db [F])."""
    N, T, E = x.shape
    Fn = arg.shape[1]
    gy = g_pool * (pooled > 0).to(g_pool.dtype)
    xp = F.pad(x, (0, 0, WINDOW - 1, WINDOW - 1))                       # [N,T+4,E]
    dW = torch.zeros(Fn, 1, WINDOW, E, dtype=x.dtype)
    n_idx = torch.arange(N).unsqueeze(1).expand(N, Fn)
    for j in range(WINDOW):
        rows = xp[n_idx, arg + j]                                       # [N,F,E]
        dW[:, 0, j, :] = (gy.unsqueeze(-1) * rows).sum(0)
    return dW, gy.sum(0)
