"""CPU oracle of the reference reader's document assembly  --  TEST INFRASTRUCTURE ONLY.

Pure-Python restatement (no mutation of the dataset) of how data.py turns a rating (user, item) into the
model inputs; only tests/ may import it.  Pinned: tests/golden/docs_{deepconn,NARRE}.npz hold what the
unmodified reference (`/root/reference/data.py`, run by oracle/gen_golden_docs.py) yields for a seeded dataset;
tests/test_docs_oracle.py replays them through this file.

  ======================  =========================================================================
  function                reference lines it follows
  ======================  =========================================================================
  ``review_lists``        data_scripts/preprocess_random_split.py:207-219 (user_reviews, item_reviews,
                          this_index_user_item) and data.py:36-63 ``calculate_reviewed_map``
  ``remove_overlap``      data.py:212-248
  ``pad_and_join``        data.py:174-210 (documents: concatenate, pad with 0, cut to input_length)
  ``pad_only``            data.py:146-172 (NARRE: every review padded / cut to narre_num_words, the list
                          to narre_num_reviews with all-zero reviews)
  ``batches``             data.py:251-337 ``iter_review`` (neighbour lists padded with total_users + 1 /
                          total_items + 1 to 10 and cut to 10, :277-282)
  ``batches_negs``        data.py:375-447 ``iter_negs`` (ranking candidates: the positive item + the sampled
                          negatives of every user in ``negs``; inputs shaped [bsz, 1+5, ...])
  ======================  =========================================================================
"""
from typing import Dict, List, Optional, Sequence


def review_lists(train_user: Sequence[int], train_item: Sequence[int], reviews: Sequence[List[int]], U: int, I: int):
    user_reviews = {u: [] for u in range(U)}
    item_reviews = {i: [] for i in range(I)}
    u_to_i = {u: [] for u in range(U)}
    i_to_u = {i: [] for i in range(I)}
    this_index: Dict[int, Dict[int, List[int]]] = {}
    for u, i, rev in zip(train_user, train_item, reviews):
        u, i = int(u), int(i)
        this_index.setdefault(u, {})[i] = [len(user_reviews[u]), len(item_reviews[i])]
        user_reviews[u].append(list(rev))
        item_reviews[i].append(list(rev))
        u_to_i[u].append(i)
        i_to_u[i].append(u)
    return user_reviews, item_reviews, this_index, u_to_i, i_to_u


def remove_overlap(u_r, i_r, u_to_i, i_to_u, user, item, this_index, test_review):
    """data.py:212-248.  Training (this_index given): the review the rating came from is taken out of both
    lists and becomes ``this``; evaluation: the lists are used whole and ``this`` is the held-out review."""
    if this_index is not None:
        ku, ki = this_index[user][item]
        this = [u_r[ku]]
        assert u_to_i[user][ku] == item and i_to_u[item][ki] == user
        return ([r for k, r in enumerate(u_r) if k != ku], [r for k, r in enumerate(i_r) if k != ki], this,
                [x for k, x in enumerate(i_to_u[item]) if k != ki], [x for k, x in enumerate(u_to_i[user]) if k != ku])
    this = [test_review if test_review is not None else [0]]
    return list(u_r), list(i_r), this, list(i_to_u[item]), list(u_to_i[user])


def pad_and_join(reviews_per_rating, T: int):
    out = []
    for revs in reviews_per_rating:
        doc = [t for r in revs for t in r]
        doc += [0] * max(0, T - len(doc))
        out.append(doc[:T])
    return out


def pad_only(reviews_per_rating, R: int, W: int):
    out = []
    for revs in reviews_per_rating:
        rows = [(list(r) + [0] * max(0, W - len(r)))[:W] for r in revs]
        rows += [[0] * W for _ in range(max(0, R - len(rows)))]
        out.append(rows[:R])
    return out


def batches(users, items, ratings, lists, hp: dict, train: bool, test_reviews: Optional[Sequence[List[int]]] = None):
    """Yields ([this, users_who_gave, items_reviewed, user_docs, item_docs, user, item], y) as python lists,
    like ``iter_review(simple=True)``; ``test_reviews[n]`` is the held-out review of evaluation rating n."""
    user_reviews, item_reviews, this_index, u_to_i, i_to_u = lists
    B = int(hp["batch_size"])
    narre = hp["model_type"] == "NARRE"
    join = (lambda x: pad_only(x, hp["narre_num_reviews"], hp["narre_num_words"])) if narre else (lambda x: pad_and_join(x, hp["input_length"]))
    for lo in range(0, len(users), B):
        acc = [[] for _ in range(7)]
        for n in range(lo, min(len(users), lo + B)):
            u, i = int(users[n]), int(items[n])
            u_r, i_r, this, who, what = remove_overlap(user_reviews[u], item_reviews[i], u_to_i, i_to_u, u, i,
                                                      this_index if train else None,
                                                      None if train or test_reviews is None else test_reviews[n])
            who = (who + [hp["total_users"] + 1] * max(0, 10 - len(who)))[:10]
            what = (what + [hp["total_items"] + 1] * max(0, 10 - len(what)))[:10]
            for slot, v in zip(acc, (this, who, what, u_r, i_r, u, i)):
                slot.append(v)
        yield [join(acc[0]), acc[1], acc[2], join(acc[3]), join(acc[4]), acc[5], acc[6]], [float(r) for r in ratings[lo:lo + B]]


def batches_negs(neg_users, neg_items, lists, hp: dict, test_review_of):
    """``iter_negs(review=True)`` of an evaluation reader (data.py:375-447).  ``neg_items[m]`` = [positive] +
    negatives of user ``neg_users[m]`` (make_negative_sets.py); ``test_review_of(u, i)`` = held-out review or None.
    Every candidate i2 gets the user's whole document and i2's whole document, but -- as in the reference, which
    calls ``remove_overlap(u_r, i_r, u, i)`` with the POSITIVE item -- the held-out review and the
    users-who-reviewed list are the positive item's for all candidates."""
    user_reviews, item_reviews, _, u_to_i, i_to_u = lists
    B = int(hp["batch_size"])
    narre = hp["model_type"] == "NARRE"
    join = (lambda x: pad_only(x, hp["narre_num_reviews"], hp["narre_num_words"])) if narre else (lambda x: pad_and_join(x, hp["input_length"]))
    for lo in range(0, len(neg_users), B):
        acc = [[] for _ in range(7)]
        for m in range(lo, min(len(neg_users), lo + B)):
            u, cands = int(neg_users[m]), [int(x) for x in neg_items[m]]
            i = cands[0]
            rows = [[] for _ in range(7)]
            for i2 in cands:
                u_r, i_r, this, who, what = remove_overlap(user_reviews[u], item_reviews[i2], u_to_i, i_to_u, u, i, None,
                                                          test_review_of(u, i))
                who = (who + [hp["total_users"] + 1] * max(0, 10 - len(who)))[:10]
                what = (what + [hp["total_items"] + 1] * max(0, 10 - len(what)))[:10]
                for slot, v in zip(rows, (this, who, what, u_r, i_r, u, i2)):
                    slot.append(v)
            for slot, v in zip(acc, (join(rows[0]), rows[1], rows[2], join(rows[3]), join(rows[4]), rows[5], rows[6])):
                slot.append(v)
        yield acc, [0.0] * len(acc[5])
