"""Generate tests/golden/*.npz by running the UNMODIFIED reference from /root/reference.

Test infrastructure (see oracle/r4r_oracle.py header).  Run in the build container only:

    python oracle/gen_golden.py

The reference modules (pytorch_models/*, loss.py, utils.py, main.py) are imported as they lie under
/root/reference -- nothing is copied.  For each model_type a tiny seeded problem is pushed through

  * ``Model(hyper_params)`` + ``utils.xavier_init``            (main.py:375-377)
  * ``model.eval(); model(data)`` on 1-D ids and on [B,n] ranking-shaped ids (eval.py:64-92)
  * ``main.train(model, MSELoss, Adam, reader, hyper_params)`` for 3 batches, dropout=0.0
    (main.py:8-71, :94-96); for TransNet, whose loop raises on torch>=1.5, the
    restated three-grad step is applied to the reference *modules* with ``utils.init_transnet_optim``.

and the inputs, initial state_dict, outputs, first-batch gradients, post-training state_dict and
metrics are stored.  /root/reference does not exist on the GPU box; the .npz files travel instead.
"""
import os
import pickle
import sys
import tempfile

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")

DIMS = dict(E=12, T=20, L=4, V=40, U=12, I=9, B=5, NB=3, R=4, W=7, NEIGH=4)   # NARRE: one neighbour id per review (NARRE.py:56)


def make_inputs(mt, d, seed):
    g = torch.Generator().manual_seed(seed)
    B = d["B"]

    def ri(hi, *shape, lo=0):
        return torch.randint(lo, hi, shape, generator=g, dtype=torch.int64)

    batches = []
    for _ in range(d["NB"]):
        uid, iid = ri(d["U"] + 1, B), ri(d["I"] + 1, B)
        y = torch.randint(1, 6, (B,), generator=g).float()
        if mt in ("bias_only", "MF", "MF_dot"):
            data = [None] * 5 + [uid, iid]
        elif mt == "NARRE":
            ur, ir = ri(d["V"], B, d["R"], d["W"]), ri(d["V"], B, d["R"], d["W"])
            ur[:, -1, 3:] = 0                                   # padded tails like data.py:146-172
            who = ri(d["U"] + 2, B, d["NEIGH"]); who[:, 3:] = d["U"] + 1       # pad id, data.py:275
            rev = ri(d["I"] + 2, B, d["NEIGH"]); rev[:, 2:] = d["I"] + 1
            data = [ri(d["V"], B, d["W"]), who, rev, ur, ir, uid, iid]
        else:
            docs = [ri(d["V"], B, d["T"]) for _ in range(3)]
            for t in docs:
                t[0, 5:] = 0                                    # pad_and_join tail, data.py:198-199
                t[1, :] = 0
            data = [docs[0], ri(d["U"] + 2, B, d["NEIGH"]), ri(d["I"] + 2, B, d["NEIGH"]), docs[1], docs[2], uid, iid]
        batches.append((data, y))
    return batches


def ranking_input(mt, d, seed):
    g = torch.Generator().manual_seed(seed + 77)
    B, n = 3, 6

    def ri(hi, *shape):
        return torch.randint(0, hi, shape, generator=g, dtype=torch.int64)

    uid, iid = ri(d["U"] + 1, B, n), ri(d["I"] + 1, B, n)
    if mt in ("bias_only", "MF", "MF_dot"):
        return [None] * 5 + [uid, iid]
    if mt == "NARRE":
        return [ri(d["V"], B, n, d["W"]), ri(d["U"] + 2, B, n, d["NEIGH"]), ri(d["I"] + 2, B, n, d["NEIGH"]),
                ri(d["V"], B, n, d["R"], d["W"]), ri(d["V"], B, n, d["R"], d["W"]), uid, iid]
    return [ri(d["V"], B, n, d["T"]), ri(d["U"] + 2, B, n, d["NEIGH"]), ri(d["I"] + 2, B, n, d["NEIGH"]),
            ri(d["V"], B, n, d["T"]), ri(d["V"], B, n, d["T"]), uid, iid]


class ListReader:
    def __init__(self, batches):
        self.batches = batches

    def iter(self, eval=False):
        for b in self.batches:
            yield b

    def __len__(self):
        return len(self.batches)


def np_state(model):
    return {k: v.detach().cpu().numpy().copy() for k, v in model.state_dict().items()}


def pack_data(prefix, data, out):
    for j, t in enumerate(data):
        if t is not None:
            out["%s.d%d" % (prefix, j)] = t.numpy()


def run(mt, seed):
    sys.path.insert(0, REF)
    import utils as ref_utils                       # noqa: reference module
    from loss import MSELoss                        # noqa
    import main as ref_main                         # noqa
    if mt in ("deepconn", "deepconn++"):
        from pytorch_models.DeepCoNN import DeepCoNN as Model
    elif mt in ("transnet", "transnet++"):
        from pytorch_models.TransNet import TransNet as Model
    elif mt == "NARRE":
        from pytorch_models.NARRE import NARRE as Model
    else:
        from pytorch_models.MF import MF as Model

    d = DIMS
    torch.manual_seed(seed)
    tmp = tempfile.mkdtemp()
    w2v = torch.randn(d["V"], d["E"]).tolist()
    with open(os.path.join(tmp, "word2vec.pkl"), "wb") as f:
        pickle.dump(w2v, f, 2)
    hp = {"model_type": mt, "data_dir": tmp, "latent_size": d["L"], "word_embed_size": d["E"],
          "dropout": 0.0, "total_users": d["U"], "total_items": d["I"], "lr": 0.002, "weight_decay": 1e-6,
          "batch_size": d["B"]}
    model = Model(hp)
    ref_utils.xavier_init(model)
    # give 1-D params non-degenerate values so bias grads/updates are exercised
    with torch.no_grad():
        for n_, p in model.named_parameters():
            if p.dim() == 1 and p.numel() > 1 and "bias" in n_ and "user_bias" not in n_ and "item_bias" not in n_:
                p.uniform_(-0.2, 0.2)
    out = {}
    for k, v in np_state(model).items():
        out["init." + k] = v
    batches = make_inputs(mt, d, seed)
    for bi, (data, y) in enumerate(batches):
        pack_data("b%d" % bi, data, out)
        out["b%d.y" % bi] = y.numpy()

    # eval-mode forwards
    model.eval()
    with torch.no_grad():
        o = model(batches[0][0])
        rk = ranking_input(mt, d, seed)
        pack_data("rank", rk, out)
        o_rk = model(rk)
    if isinstance(o, list):
        for j in range(3):
            out["eval.out%d" % j] = o[j].numpy()
            out["rank.out%d" % j] = o_rk[j].numpy()
    else:
        out["eval.out0"] = o.numpy()
        out["rank.out0"] = o_rk.numpy()

    # first-batch gradients (train mode, dropout 0)
    model.train()
    crit = MSELoss(hp)
    data, y = batches[0]
    model.zero_grad()
    if mt in ("transnet", "transnet++"):
        o = model(data)
        src = list(model.source.parameters())
        sfm = list(model.source_fm.parameters())
        if mt == "transnet++":
            sfm += [model.user_embedding.weight, model.item_embedding.weight]
        tgt = [p for p in model.target.parameters() if p.requires_grad]
        g_t = torch.autograd.grad(crit(o[1], y), tgt, retain_graph=True, allow_unused=True)
        g_s = torch.autograd.grad(o[2], src, retain_graph=True, allow_unused=True)
        g_f = torch.autograd.grad(crit(o[0], y), sfm, allow_unused=True)
        names = {id(p): n_ for n_, p in model.named_parameters()}
        for plist, glist in ((tgt, g_t), (src, g_s), (sfm, g_f)):
            for p, g in zip(plist, glist):
                if g is not None:
                    out["grad." + names[id(p)]] = g.numpy()
        for j in range(3):
            out["train.out%d" % j] = o[j].detach().numpy()
    else:
        o = model(data)
        crit(o, y).backward()
        for n_, p in model.named_parameters():
            if p.grad is not None:
                out["grad." + n_] = p.grad.numpy().copy()
        out["train.out0"] = o.detach().numpy()
    model.zero_grad()

    # K training batches through the reference loop
    if mt in ("transnet", "transnet++"):
        opts = ref_utils.init_transnet_optim(hp, model)
        o_src, o_sfm, o_tgt, _ = opts
        tot = tot_t = tot_x = 0.0
        n = 0
        for data, y in batches:
            model.zero_grad()
            for oo in opts:
                oo.zero_grad()
            o = model(data)
            loss_t = crit(o[1], y)
            se = crit(o[0], y, return_mean=False)
            g_t = torch.autograd.grad(loss_t, tgt, retain_graph=True, allow_unused=True)
            g_s = torch.autograd.grad(o[2], src, retain_graph=True, allow_unused=True)
            g_f = torch.autograd.grad(se.mean(), sfm, allow_unused=True)
            for p, g in zip(tgt, g_t):
                p.grad = g
            o_tgt.step()
            for p in tgt:
                p.grad = None
            for p, g in zip(src, g_s):
                p.grad = g
            o_src.step()
            for p in src:
                p.grad = None
            for p, g in zip(sfm, g_f):
                p.grad = g
            o_sfm.step()
            for p in sfm:
                p.grad = None
            tot += float(se.sum()); tot_t += float(loss_t); tot_x += float(o[2]); n += int(y.shape[0])
        out["metric.MSE_sum"] = np.float64(tot)
        out["metric.MSE_target_sum"] = np.float64(tot_t)
        out["metric.MSE_transform_sum"] = np.float64(tot_x)
        out["metric.N"] = np.int64(n)
        # also record that the reference's own loop cannot run here
        try:
            m2 = Model(hp)
            ref_main.train(m2, crit, ref_utils.init_transnet_optim(hp, m2), ListReader(batches[:1]), hp)
            out["ref_train_raises"] = np.int64(0)
        except RuntimeError:
            out["ref_train_raises"] = np.int64(1)
    else:
        opt = torch.optim.Adam(model.parameters(), lr=hp["lr"], weight_decay=hp["weight_decay"])
        metrics = ref_main.train(model, crit, opt, ListReader(batches), hp)
        out["metric.MSE"] = np.float64(metrics["MSE"])
        out["metric.N"] = np.int64(sum(int(b[1].shape[0]) for b in batches))
    for k, v in np_state(model).items():
        out["final." + k] = v
    out["dims"] = np.array([d[k] for k in ("E", "T", "L", "V", "U", "I", "B", "NB", "R", "W", "NEIGH")], dtype=np.int64)
    return out


def main():
    os.makedirs(OUT, exist_ok=True)
    for seed, mt in enumerate(["deepconn", "deepconn++", "bias_only", "MF_dot", "MF", "NARRE", "transnet", "transnet++"]):
        res = run(mt, 100 + seed)
        fn = os.path.join(OUT, mt.replace("+", "p") + ".npz")
        np.savez_compressed(fn, **res)
        print(mt, "->", fn, os.path.getsize(fn), "bytes", "keys", len(res))


if __name__ == "__main__":
    main()
