"""Generates tests/golden/neumf.npz by running the UNMODIFIED reference GMF / MLP / NeuMF
(/root/reference/pytorch_models/NeuMF.py) through the reference's own ``main.train`` + ``torch.optim.Adam``
(main.py:8-71, :94-96) and ``NeuMF.init`` (NeuMF.py:93-112).  Build container only:

    python oracle/gen_golden_neumf.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
DIMS = dict(L=4, U=12, I=9, B=5, NB=3)


class ListReader:
    def __init__(self, batches):
        self.batches = batches

    def iter(self, eval=False):
        yield from self.batches

    def __len__(self):
        return len(self.batches)


def main():
    sys.path.insert(0, REF)
    import main as ref_main                                     # noqa: reference module
    import utils as ref_utils                                   # noqa
    from loss import MSELoss                                    # noqa
    from pytorch_models.NeuMF import GMF, MLP, NeuMF            # noqa
    d = DIMS
    g = torch.Generator().manual_seed(21)
    ri = lambda hi, *s: torch.randint(0, hi, s, generator=g, dtype=torch.int64)
    batches = [([None] * 5 + [ri(d["U"] + 1, d["B"]), ri(d["I"] + 1, d["B"])], torch.randint(1, 6, (d["B"],), generator=g).float())
               for _ in range(d["NB"])]
    rank = [None] * 5 + [ri(d["U"] + 1, 3, 6), ri(d["I"] + 1, 3, 6)]
    out = {"dims": np.array([d["L"], d["U"], d["I"], d["B"], d["NB"]], dtype=np.int64),
           "rank.d5": rank[5].numpy(), "rank.d6": rank[6].numpy()}
    for b, (data, y) in enumerate(batches):
        out["b%d.d5" % b], out["b%d.d6" % b], out["b%d.y" % b] = data[5].numpy(), data[6].numpy(), y.numpy()
    hp = {"latent_size": d["L"], "dropout": 0.0, "total_users": d["U"], "total_items": d["I"], "lr": 0.002,
          "weight_decay": 1e-6, "batch_size": d["B"]}
    trained = {}
    for name, cls in (("GMF", GMF), ("MLP", MLP), ("NeuMF", NeuMF)):
        torch.manual_seed(5)
        hp["model_type"] = name
        model = cls(hp)
        ref_utils.xavier_init(model)
        with torch.no_grad():
            for n_, p in model.named_parameters():
                if p.dim() == 1 and "bias" in n_ and p.numel() > 1:
                    p.uniform_(-0.2, 0.2) if "user_bias" not in n_ and "item_bias" not in n_ else p.add_(torch.randn(p.shape, generator=g) * 0.05)
        if name == "NeuMF":
            model.init(trained["GMF"], trained["MLP"])          # NeuMF.py:93-112 on the trained pre-models
        for k, v in model.state_dict().items():
            out["%s.init.%s" % (name, k)] = v.detach().numpy().copy()
        model.eval()
        with torch.no_grad():
            out["%s.eval.b0" % name] = model(batches[0][0]).numpy()
            out["%s.eval.rank" % name] = model(rank).numpy()
        opt = torch.optim.Adam(model.parameters(), lr=hp["lr"], weight_decay=hp["weight_decay"])
        metrics = ref_main.train(model, MSELoss(hp), opt, ListReader(batches), hp)
        out["%s.metric.MSE" % name] = np.array([metrics["MSE"]], dtype=np.float64)
        for k, v in model.state_dict().items():
            out["%s.final.%s" % (name, k)] = v.detach().numpy().copy()
        trained[name] = model
        print(name, metrics, sorted(model.state_dict().keys()))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "neumf.npz"), **out)


if __name__ == "__main__":
    main()
