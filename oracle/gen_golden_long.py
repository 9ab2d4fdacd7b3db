"""Generate tests/golden/deepconn_long.npz by running the UNMODIFIED reference from /root/reference on LONG,
padded documents (T = 700 rows: three 256-position tiles per document and trailing padding runs of every length),
so that the multi-tile path of the fused conv kernel and its padding-run work plan (r4r_doc_plan) are pinned to the
reference itself, not only to the oracle.  Test infrastructure; run in the build container only:

    python oracle/gen_golden_long.py

Stored: inputs, initial state_dict, the reference's pooled conv features of both towers (input of TextCNN.fc,
captured with a forward hook on the unmodified module: common_pytorch_models.py:29-37), eval-mode ratings,
first-batch gradients, and the metrics / final state_dict of ``main.train`` over the batches (main.py:8-71)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import gen_golden as G                                          # noqa: E402  (shared helpers; imports nothing of the product)

DIMS = dict(E=64, T=700, L=10, V=300, U=12, I=9, B=6, NB=3, R=4, W=7, NEIGH=4)


def make_inputs(d, seed):
    g = torch.Generator().manual_seed(seed)
    B, T = d["B"], d["T"]
    ri = lambda hi, *shape: torch.randint(0, hi, shape, generator=g, dtype=torch.int64)
    lengths = [T, 0, 1, 253, 254, 255, 256, 257, 509, 510, 511, 512, 640, 697, 698, 699, 20, 333]
    batches, k = [], 0
    for _ in range(d["NB"]):
        docs = [ri(d["V"], B, T) for _ in range(2)]
        for t in docs:
            for b in range(B):
                n = lengths[k % len(lengths)]
                k += 1
                t[b, n:] = 0 if k % 5 else 7                    # pad_and_join tail (data.py:198-199); one run of a non-zero token
        uid, iid = ri(d["U"] + 1, B), ri(d["I"] + 1, B)
        y = torch.randint(1, 6, (B,), generator=g).float()
        batches.append(([None, None, None, docs[0], docs[1], uid, iid], y))
    return batches


def run(seed=321):
    sys.path.insert(0, G.REF)
    import utils as ref_utils                                   # noqa: reference modules, unmodified
    from loss import MSELoss                                    # noqa
    import main as ref_main                                     # noqa
    from pytorch_models.DeepCoNN import DeepCoNN as Model       # noqa
    import pickle
    import tempfile
    d = DIMS
    torch.manual_seed(seed)
    tmp = tempfile.mkdtemp()
    with open(os.path.join(tmp, "word2vec.pkl"), "wb") as f:
        pickle.dump(torch.randn(d["V"], d["E"]).tolist(), f, 2)
    hp = {"model_type": "deepconn", "data_dir": tmp, "latent_size": d["L"], "word_embed_size": d["E"], "dropout": 0.0,
          "total_users": d["U"], "total_items": d["I"], "lr": 0.002, "weight_decay": 1e-6, "batch_size": d["B"]}
    model = Model(hp)
    ref_utils.xavier_init(model)
    with torch.no_grad():
        # word rows and filters of O(0.3): conv values of O(1), so pooled features are a meaningful comparison
        model.word2vec.weight.mul_(8.0)
        for n_, p in model.named_parameters():
            if p.dim() == 1 and p.numel() > 1 and "bias" in n_ and "user_bias" not in n_ and "item_bias" not in n_:
                p.uniform_(-0.2, 0.2)
    out = {"init." + k: v for k, v in G.np_state(model).items()}
    batches = make_inputs(d, seed)
    for bi, (data, y) in enumerate(batches):
        G.pack_data("b%d" % bi, data, out)
        out["b%d.y" % bi] = y.numpy()
    grabbed = {}
    hooks = [model.user_conv.fc.register_forward_hook(lambda m, i, o: grabbed.__setitem__("user", i[0].detach().numpy().copy())),
             model.item_conv.fc.register_forward_hook(lambda m, i, o: grabbed.__setitem__("item", i[0].detach().numpy().copy()))]
    model.eval()
    with torch.no_grad():
        out["eval.out0"] = model(batches[0][0]).numpy()
    out["pooled.user"], out["pooled.item"] = grabbed["user"], grabbed["item"]
    for h in hooks:
        h.remove()
    model.train()
    crit = MSELoss(hp)
    data, y = batches[0]
    model.zero_grad()
    o = model(data)
    crit(o, y).backward()
    for n_, p in model.named_parameters():
        if p.grad is not None:
            out["grad." + n_] = p.grad.numpy().copy()
    out["train.out0"] = o.detach().numpy()
    model.zero_grad()
    opt = torch.optim.Adam(model.parameters(), lr=hp["lr"], weight_decay=hp["weight_decay"])
    metrics = ref_main.train(model, crit, opt, G.ListReader(batches), hp)
    out["metric.MSE"] = np.float64(metrics["MSE"])
    out["metric.N"] = np.int64(sum(int(b[1].shape[0]) for b in batches))
    for k, v in G.np_state(model).items():
        out["final." + k] = v
    out["dims"] = np.array([d[k] for k in ("E", "T", "L", "V", "U", "I", "B", "NB", "R", "W", "NEIGH")], dtype=np.int64)
    return out


if __name__ == "__main__":
    res = run()
    fn = os.path.join(G.OUT, "deepconn_long.npz")
    np.savez_compressed(fn, **res)
    print("deepconn_long ->", fn, os.path.getsize(fn), "bytes", "keys", len(res))
