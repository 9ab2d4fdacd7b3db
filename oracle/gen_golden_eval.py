"""Generates tests/golden/eval_<model_type>.npz by running the UNMODIFIED reference evaluation
(/root/reference/eval.py: evaluate :11-62, eval_ranking :64-92) on the reference model classes loaded with
the initial state and batches already stored in tests/golden/<model_type>.npz.  Build container only:

    python oracle/gen_golden_eval.py
"""
import os
import pickle
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference"
MODELS = ["deepconn", "deepconn++", "NARRE", "transnet++", "MF_dot"]


class EvalReader:
    def __init__(self, batches, rank):
        self.batches, self.rank = batches, rank
        self.data = list(range(int(rank[5].shape[0])))             # eval_ranking sizes its result by len(reader.data)

    def iter(self, eval=False):
        for b in self.batches:
            yield b

    def iter_negs(self, review):
        yield self.rank, torch.zeros(int(self.rank[5].shape[0]))


def counts_from(batches, slot):
    c = {}
    for data, _ in batches[:-1]:                                    # the last batch holds unseen ids too
        for v in data[slot].tolist():
            c[v] = c.get(v, 0) + 1
    return c


def main():
    sys.path.insert(0, REF)
    import eval as ref_eval                                         # noqa: the reference's eval.py
    from loss import MSELoss                                        # noqa
    from tests.helpers import golden_batches, golden_data, golden_hp, golden_state, load_golden
    for mt in MODELS:
        if mt in ("deepconn", "deepconn++"):
            from pytorch_models.DeepCoNN import DeepCoNN as Model
        elif mt in ("transnet", "transnet++"):
            from pytorch_models.TransNet import TransNet as Model
        elif mt == "NARRE":
            from pytorch_models.NARRE import NARRE as Model
        else:
            from pytorch_models.MF import MF as Model
        z, dims = load_golden(mt)
        hp = golden_hp(mt, dims, dropout=0.6)                       # eval() must switch dropout off
        tmp = tempfile.mkdtemp()
        with open(os.path.join(tmp, "word2vec.pkl"), "wb") as f:
            pickle.dump(np.zeros((dims["V"], dims["E"]), dtype=np.float32).tolist(), f, 2)
        hp["data_dir"] = tmp
        model = Model(hp)
        model.load_state_dict(golden_state(z, "init"))
        batches = golden_batches(z, dims)
        reader = EvalReader(batches, golden_data(z, "rank"))
        user_count, item_count = counts_from(batches, 5), counts_from(batches, 6)
        out = {"user_count_keys": np.array(sorted(user_count), dtype=np.int64),
               "user_count_vals": np.array([user_count[k] for k in sorted(user_count)], dtype=np.int64),
               "item_count_keys": np.array(sorted(item_count), dtype=np.int64),
               "item_count_vals": np.array([item_count[k] for k in sorted(item_count)], dtype=np.int64)}
        metrics, umap, imap = ref_eval.evaluate(model, MSELoss(hp), reader, hp, dict(user_count), dict(item_count), True)
        for k, v in metrics.items():
            out["metric." + k] = np.array([v], dtype=np.float64)
        for name, m in (("umap", umap), ("imap", imap)):
            keys = sorted(m)
            out[name + ".keys"] = np.array(keys, dtype=np.int64)
            out[name + ".sizes"] = np.array([len(m[k]) for k in keys], dtype=np.int64)
            out[name + ".vals"] = np.array([x for k in keys for x in m[k]], dtype=np.float64)
        rk = ref_eval.eval_ranking(model, reader, hp, True)
        out["metric.HR@1"] = np.array([rk["HR@1"]], dtype=np.float64)
        path = os.path.join(ROOT, "tests", "golden", "eval_%s.npz" % mt.replace("+", "p"))
        np.savez_compressed(path, **out)
        print(mt, metrics, rk, {k: len(v) for k, v in umap.items()})


if __name__ == "__main__":
    main()
