"""Generates tests/golden/docs_{deepconn,NARRE}.npz by running the UNMODIFIED reference reader
(/root/reference/data.py: DataLoader.iter_review -> remove_overlap / pad_and_join / pad_only) on a small
seeded review dataset.  Run in the build container only (needs /root/reference):

    python oracle/gen_golden_docs.py

`surprise` (imported by data.py:4, used only by get_surprise_format_data) is absent here and stubbed.
The dataset is stored in the CSR form reviews4rec_b200.readers.ReviewStore takes; the expected batches are
what the reference yields with simple=True (python lists, i.e. before LongTensor()).
"""
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"


def make_dataset(seed, U, I, n_train, n_eval, V, max_len):
    rng = np.random.default_rng(seed)
    pairs = set()
    while len(pairs) < n_train + n_eval:
        u = int(rng.integers(0, U)) if rng.random() < 0.6 else 0          # user 0 is a heavy reviewer
        pairs.add((u, int(rng.integers(0, I))))
    pairs = [tuple(p) for p in rng.permutation(sorted(pairs))]
    def review():
        n = int(rng.integers(0, max_len + 1))                               # empty reviews happen after stripping
        return [int(x) for x in rng.integers(1, V, n)]
    train = [[u, i, float(rng.integers(1, 6)), review()] for u, i in pairs[:n_train]]
    evals = [[u, i, float(rng.integers(1, 6)), review()] for u, i in pairs[n_train:]]
    return train, evals


def build_reference_inputs(train, evals, U, I):
    """preprocess_random_split.py:207-219 (train review lists + this_index_user_item) and test_reviews."""
    user_reviews = {u: [] for u in range(U)}
    item_reviews = {i: [] for i in range(I)}
    this_index = {}
    for u, i, _, rev in train:
        this_index.setdefault(u, {})[i] = [len(user_reviews[u]), len(item_reviews[i])]
        user_reviews[u].append(list(rev))
        item_reviews[i].append(list(rev))
    test_reviews = {}
    for u, i, _, rev in evals:
        test_reviews.setdefault(u, {})[i] = list(rev)
    return user_reviews, item_reviews, this_index, test_reviews


def csr(train, evals, U, I):
    tok = np.array([t for r in train for t in r[3]], dtype=np.int32)
    rev_off = np.concatenate([[0], np.cumsum([len(r[3]) for r in train])]).astype(np.int64)
    etok = np.array([t for r in evals for t in r[3]], dtype=np.int32)
    eoff = np.concatenate([[0], np.cumsum([len(r[3]) for r in evals])]).astype(np.int64)
    return dict(tok=tok, rev_off=rev_off,
                train_user=np.array([r[0] for r in train], dtype=np.int64), train_item=np.array([r[1] for r in train], dtype=np.int64),
                train_y=np.array([r[2] for r in train], dtype=np.float32),
                eval_user=np.array([r[0] for r in evals], dtype=np.int64), eval_item=np.array([r[1] for r in evals], dtype=np.int64),
                eval_y=np.array([r[2] for r in evals], dtype=np.float32), eval_tok=etok, eval_off=eoff)


def run(model_type, out_path):
    sys.modules.setdefault("surprise", types.ModuleType("surprise"))
    sys.path.insert(0, REF)
    import copy
    import data as refdata                                                   # the reference's reader, unmodified
    U, I, V, T, R, W, B = 7, 5, 60, 30, 4, 6, 4
    train, evals = make_dataset(11, U, I, 26, 7, V, 14)
    user_reviews, item_reviews, this_index, test_reviews = build_reference_inputs(train, evals, U, I)
    hp = {"model_type": model_type, "batch_size": B, "input_length": T, "narre_num_reviews": R, "narre_num_words": W,
          "total_users": U, "total_items": I}
    out = csr(train, evals, U, I)
    out["dims"] = np.array([U, I, V, T, R, W, B], dtype=np.int64)
    for split, rows, kw in (("train", train, dict(this_index_user_item=this_index)), ("eval", evals, dict(test_reviews=test_reviews))):
        # iter_review / pad_only mutate the review lists in place (data.py:160-170): give every loader its own copy
        ur, ir = copy.deepcopy(user_reviews), copy.deepcopy(item_reviews)
        if split == "train":
            loader = refdata.DataLoader(hp, [r[:3] for r in rows], ur, ir, None, **kw)
            train_loader = loader
        else:
            loader = refdata.DataLoader(hp, [r[:3] for r in rows], ur, ir, None, train_loader=train_loader, **kw)
        for b, (data, y) in enumerate(loader.iter_review(simple=True)):
            for j, d in enumerate(copy.deepcopy(data)):
                out["%s.b%d.d%d" % (split, b, j)] = np.array(d, dtype=np.int64)
            out["%s.b%d.y" % (split, b)] = np.array(y, dtype=np.float32)
        out["%s.nb" % split] = np.array([b + 1], dtype=np.int64)
        if split == "eval":
            # ranking candidates (data.py:375-447 iter_negs): per user the positive item + 5 sampled negatives, in the
            # format data_scripts/make_negative_sets.py writes: negs[user] = [[positive], [negatives]]
            rng = np.random.default_rng(5)
            negs = {}
            for u, i, _, _ in rows:
                if u not in negs and len(negs) < 5:
                    negs[u] = [[i], [int(x) for x in rng.choice(I, 5, replace=False)]]
            ur, ir = copy.deepcopy(user_reviews), copy.deepcopy(item_reviews)
            nl = refdata.DataLoader(hp, [r[:3] for r in rows], ur, ir, negs, train_loader=train_loader, **kw)
            out["negs.users"] = np.array(list(negs), dtype=np.int64)
            out["negs.items"] = np.array([negs[u][0] + negs[u][1] for u in negs], dtype=np.int64)
            for b, (data, y) in enumerate(nl.iter_negs(True)):
                for j, d in enumerate(data):
                    out["negs.b%d.d%d" % (b, j)] = d.numpy().astype(np.int64)
                out["negs.b%d.y" % b] = y.numpy().astype(np.float32)
            out["negs.nb"] = np.array([b + 1], dtype=np.int64)
    np.savez_compressed(out_path, **out)
    print(out_path, {k: v.shape for k, v in out.items() if k.startswith("train.b0")})


if __name__ == "__main__":
    for mt in ("deepconn", "NARRE"):
        run(mt, os.path.join(ROOT, "tests", "golden", "docs_%s.npz" % mt))
        sys.modules.pop("data", None)
