"""The document-assembly oracle (oracle/r4r_data_oracle.py) against what the unmodified reference reader
yielded for the seeded dataset of oracle/gen_golden_docs.py (tests/golden/docs_*.npz)."""
import os

import numpy as np
import pytest
import torch

from tests.helpers import GOLDEN


def load_docs_golden(mt):
    z = np.load(os.path.join(GOLDEN, "docs_%s.npz" % mt))
    U, I, V, T, R, W, B = [int(x) for x in z["dims"]]
    hp = {"model_type": mt, "batch_size": B, "input_length": T, "narre_num_reviews": R, "narre_num_words": W,
          "total_users": U, "total_items": I}
    return z, hp, (U, I, V)


def split_reviews(tok, off):
    return [tok[off[k]:off[k + 1]].tolist() for k in range(len(off) - 1)]


@pytest.mark.parametrize("mt", ["deepconn", "NARRE"])
def test_oracle_reproduces_reference_reader(mt):
    from oracle import r4r_data_oracle as D
    z, hp, (U, I, V) = load_docs_golden(mt)
    lists = D.review_lists(z["train_user"], z["train_item"], split_reviews(z["tok"], z["rev_off"]), U, I)
    for split, train in (("train", True), ("eval", False)):
        test_reviews = None if train else split_reviews(z["eval_tok"], z["eval_off"])
        got = list(D.batches(z[split + "_user"], z[split + "_item"], z[split + "_y"], lists, hp, train, test_reviews))
        assert len(got) == int(z[split + ".nb"][0])
        for b, (data, y) in enumerate(got):
            for j, d in enumerate(data):
                want = z["%s.b%d.d%d" % (split, b, j)]
                assert np.array_equal(np.array(d, dtype=np.int64), want), (split, b, j)
            assert np.array_equal(np.array(y, dtype=np.float32), z["%s.b%d.y" % (split, b)])


@pytest.mark.gpu
@pytest.mark.parametrize("mt", ["deepconn", "NARRE"])
def test_device_assembly_reproduces_reference_reader(mt):
    """ReviewStore + CsrReader (r4r_docs_assemble) yield, bit for bit, the batches the unmodified reference
    reader produced (golden), for the train split (own review left out) and an evaluation split."""
    import torch
    from reviews4rec_b200.readers import CsrReader, ReviewStore
    z, hp, (U, I, V) = load_docs_golden(mt)
    store = ReviewStore(z["tok"], z["rev_off"], z["train_user"], z["train_item"], U, I, "cuda")
    readers = {"train": CsrReader(hp, store, z["train_y"], train=True),
               "eval": CsrReader(hp, store, z["eval_y"], train=False, users=z["eval_user"], items=z["eval_item"],
                                 this_tok=z["eval_tok"], this_off=z["eval_off"])}
    for split, reader in readers.items():
        assert len(reader) == int(z[split + ".nb"][0])
        for b, (data, y) in enumerate(reader.iter()):
            for j, d in enumerate(data):
                want = z["%s.b%d.d%d" % (split, b, j)]
                assert d.dtype == torch.int64 and np.array_equal(d.cpu().numpy(), want), (split, b, j)
            assert np.array_equal(y.cpu().numpy(), z["%s.b%d.y" % (split, b)])


@pytest.mark.gpu
@pytest.mark.parametrize("mt,T", [("deepconn", 1000), ("NARRE", 0), ("transnet", 257)])
def test_device_assembly_vs_oracle_larger(mt, T):
    """Random larger dataset (heavy users whose documents overflow input_length, empty reviews, users without
    reviews): device assembly == the pinned oracle."""
    import torch
    from oracle import r4r_data_oracle as D
    from reviews4rec_b200.readers import CsrReader, ReviewStore
    rng = np.random.default_rng(3)
    U, I, V, n, B = 40, 25, 500, 600, 64
    pairs = set()
    while len(pairs) < n + 50:
        u = int(rng.integers(0, U - 2)) if rng.random() < 0.7 else 0
        pairs.add((u, int(rng.integers(0, I - 1))))
    pairs = rng.permutation(sorted(pairs))
    lens = rng.integers(0, 120, len(pairs))
    revs = [rng.integers(1, V, k).astype(np.int32) for k in lens]
    tr, ev = slice(0, n), slice(n, len(pairs))
    tok = np.concatenate(revs[tr]); off = np.concatenate([[0], np.cumsum(lens[tr])])
    etok = np.concatenate(revs[ev]); eoff = np.concatenate([[0], np.cumsum(lens[ev])])
    y = rng.integers(1, 6, len(pairs)).astype(np.float32)
    hp = {"model_type": mt, "batch_size": B, "input_length": T or 1000, "narre_num_reviews": 10, "narre_num_words": 50,
          "total_users": U, "total_items": I}
    store = ReviewStore(tok, off, pairs[tr, 0], pairs[tr, 1], U, I, "cuda")
    lists = D.review_lists(pairs[tr, 0], pairs[tr, 1], [r.tolist() for r in revs[tr]], U, I)
    cases = [(CsrReader(hp, store, y[tr], train=True), D.batches(pairs[tr, 0], pairs[tr, 1], y[tr], lists, hp, True)),
             (CsrReader(hp, store, y[ev], train=False, users=pairs[ev, 0], items=pairs[ev, 1], this_tok=etok, this_off=eoff),
              D.batches(pairs[ev, 0], pairs[ev, 1], y[ev], lists, hp, False, [r.tolist() for r in revs[ev]]))]
    for reader, want in cases:
        for (data, yy), (wd, wy) in zip(reader.iter(), want):
            for j in range(7):
                assert np.array_equal(data[j].cpu().numpy(), np.array(wd[j], dtype=np.int64)), j
            assert np.array_equal(yy.cpu().numpy(), np.array(wy, dtype=np.float32))


def test_review_list_side_matches_this_index_user_item():
    """Host logic of ReviewStore (CPU): list order, neighbour ids and ranks equal the reference's
    user_reviews / u_to_i_map / this_index_user_item construction."""
    from oracle import r4r_data_oracle as D
    from reviews4rec_b200.readers import review_list_side
    z, hp, (U, I, V) = load_docs_golden("deepconn")
    tu, ti = z["train_user"], z["train_item"]
    revs = split_reviews(z["tok"], z["rev_off"])
    user_reviews, item_reviews, this_index, u_to_i, i_to_u = D.review_lists(tu, ti, revs, U, I)
    for ids, other, n_lists, lists, nbmap, col in ((tu, ti, U, user_reviews, u_to_i, 0), (ti, tu, I, item_reviews, i_to_u, 1)):
        ptr, rev, nb, rank = review_list_side(ids, other, n_lists)
        for l in range(n_lists):
            got = [revs[r] for r in rev[ptr[l]:ptr[l + 1]]]
            assert got == lists[l] and nb[ptr[l]:ptr[l + 1]].tolist() == nbmap[l]
        for n in range(len(tu)):
            assert int(rank[n]) == this_index[int(tu[n])][int(ti[n])][col]


def _write_reference_pickles(tmp, z, U, I):
    """The files data_scripts/preprocess_random_split.py:282-299 writes, for the golden dataset."""
    import pickle
    revs = split_reviews(z["tok"], z["rev_off"])
    user_reviews = {u: [] for u in range(U)}
    item_reviews = {i: [] for i in range(I)}
    this_index = {}
    train = []
    for u, i, y, rev in zip(z["train_user"].tolist(), z["train_item"].tolist(), z["train_y"].tolist(), revs):
        this_index.setdefault(u, {})[i] = [len(user_reviews[u]), len(item_reviews[i])]
        user_reviews[u].append(rev)
        item_reviews[i].append(rev)
        train.append([u, i, y])
    erevs = split_reviews(z["eval_tok"], z["eval_off"])
    test_reviews, rows = {}, []
    for u, i, y, rev in zip(z["eval_user"].tolist(), z["eval_item"].tolist(), z["eval_y"].tolist(), erevs):
        test_reviews.setdefault(u, {})[i] = rev
        rows.append([u, i, y])
    objs = {"train": train, "test": rows[:4], "val": rows[4:], "user_reviews": user_reviews, "item_reviews": item_reviews,
            "this_index_user_item": this_index, "test_reviews": test_reviews, "num_users_items": [U, I, 59]}
    for name, obj in objs.items():
        with open(os.path.join(tmp, name + ".pkl"), "wb") as f:
            pickle.dump(obj, f, 2)
    return rows


def test_reference_pickles_to_arrays(tmp_path):
    """Host side of readers.load_data (CPU): the reference's pickles become exactly the CSR arrays the golden
    dataset was generated from."""
    from reviews4rec_b200.readers import reference_pickles_to_arrays
    z, hp, (U, I, V) = load_docs_golden("deepconn")
    rows = _write_reference_pickles(str(tmp_path), z, U, I)
    a = reference_pickles_to_arrays(str(tmp_path) + "/")
    assert (a["total_users"], a["total_items"], a["total_words"]) == (U, I, 59)
    for k in ("tok", "rev_off", "train_user", "train_item", "train_y"):
        assert np.array_equal(a[k], z[k]) and a[k].dtype == z[k].dtype, k
    got_u = np.concatenate([a["test_user"], a["val_user"]])
    got_tok = np.concatenate([a["test_tok"], a["val_tok"]])
    got_off = np.concatenate([a["test_off"], a["val_off"][1:] + a["test_off"][-1]])
    assert np.array_equal(got_u, z["eval_user"]) and np.array_equal(got_tok, z["eval_tok"]) and np.array_equal(got_off, z["eval_off"])
    assert len(a["test_y"]) == 4 and len(a["val_y"]) == len(rows) - 4


def test_load_data_wiring_cpu(tmp_path, monkeypatch):
    """readers.load_data's host wiring (what goes into ReviewStore / CsrReader), with the device classes replaced
    by recorders -- the device end is covered by tests/test_z_host_mirrors.py on the GPU."""
    from reviews4rec_b200 import readers
    z, hp, (U, I, V) = load_docs_golden("deepconn")
    _write_reference_pickles(str(tmp_path), z, U, I)
    made = {}

    class Store:
        def __init__(self, tok, rev_off, tu, ti, total_users, total_items, device):
            made["store"] = (tok, rev_off, tu, ti, total_users, total_items, device)

    class Reader:
        def __init__(self, hyper_params, store, ratings, train, users=None, items=None, this_tok=None, this_off=None, negs=None):
            made.setdefault("readers", []).append((train, len(ratings), None if users is None else len(users),
                                                    None if this_off is None else len(this_off)))
            made.setdefault("negs", []).append(negs)

    monkeypatch.setattr(readers, "ReviewStore", Store)
    monkeypatch.setattr(readers, "CsrReader", Reader)
    import pickle
    with open(os.path.join(str(tmp_path), "negs.pkl"), "wb") as f:       # make_negative_sets.py format
        pickle.dump({int(u): [[int(c[0])], [int(x) for x in c[1:]]] for u, c in zip(z["negs.users"], z["negs.items"])}, f, 2)
    hp = dict(hp, data_dir=str(tmp_path) + "/")
    train, test, val, hp2 = readers.load_data(hp, "cuda:0")
    assert made["negs"][0] is None and made["negs"][2] is None            # only the test reader ranks
    assert np.array_equal(made["negs"][1][0], z["negs.users"]) and np.array_equal(made["negs"][1][1], z["negs.items"])
    assert hp2 is hp and (hp["total_users"], hp["total_items"], hp["total_words"]) == (U, I, 59)
    tok, rev_off, tu, ti, nu, ni, dev = made["store"]
    assert np.array_equal(tok, z["tok"]) and np.array_equal(rev_off, z["rev_off"]) and (nu, ni, dev) == (U, I, "cuda:0")
    n_eval = len(z["eval_y"])
    assert made["readers"] == [(True, len(z["train_y"]), None, None), (False, 4, 4, 5), (False, n_eval - 4, n_eval - 4, n_eval - 3)]


def _held_out(z):
    revs = split_reviews(z["eval_tok"], z["eval_off"])
    table = {(int(u), int(i)): r for u, i, r in zip(z["eval_user"], z["eval_item"], revs)}
    return lambda u, i: table.get((u, i))


@pytest.mark.parametrize("mt", ["deepconn", "NARRE"])
def test_oracle_reproduces_reference_ranking_candidates(mt):
    from oracle import r4r_data_oracle as D
    z, hp, (U, I, V) = load_docs_golden(mt)
    lists = D.review_lists(z["train_user"], z["train_item"], split_reviews(z["tok"], z["rev_off"]), U, I)
    got = list(D.batches_negs(z["negs.users"], z["negs.items"], lists, hp, _held_out(z)))
    assert len(got) == int(z["negs.nb"][0])
    for b, (data, y) in enumerate(got):
        for j, d in enumerate(data):
            assert np.array_equal(np.array(d, dtype=np.int64), z["negs.b%d.d%d" % (b, j)]), (b, j)
        assert np.array_equal(np.array(y, dtype=np.float32), z["negs.b%d.y" % b])


def test_ranking_candidates_host_preparation_cpu():
    """CsrReader._prepare_negs (host): the held-out review of (user, positive item) replicated for the C candidates."""
    import types
    from reviews4rec_b200.readers import CsrReader
    z, hp, (U, I, V) = load_docs_golden("deepconn")
    store = types.SimpleNamespace(device=torch.device("cpu"))
    r = CsrReader(hp, store, z["eval_y"], train=False, users=z["eval_user"], items=z["eval_item"],
                  this_tok=z["eval_tok"], this_off=z["eval_off"], negs=(z["negs.users"], z["negs.items"]))
    ng, held = r.negs, _held_out(z)
    assert ng["C"] == 6 and tuple(ng["items"].shape) == (5, 6) and ng["this_off"].numel() == 5 * 6 + 1
    for m, (u, c) in enumerate(zip(z["negs.users"].tolist(), z["negs.items"].tolist())):
        want = held(u, c[0])
        want = [0] if want is None else want
        for k in range(6):
            lo, hi = int(ng["this_off"][m * 6 + k]), int(ng["this_off"][m * 6 + k + 1])
            assert ng["this_tok"][lo:hi].tolist() == list(want)
    with pytest.raises(RuntimeError):
        next(CsrReader(hp, store, z["eval_y"], train=False, users=z["eval_user"], items=z["eval_item"]).iter_negs(True))
