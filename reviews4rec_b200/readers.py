"""B200-side counterpart of the reference's fast reader (data_fast.py:14-123).

The reference keeps the eight HDF5 datasets ``a..h`` of a split (make_quick_data.py:21-44: ``a`` this
review, ``b`` users who reviewed the item, ``c`` items the user reviewed, ``d`` user document, ``e`` item
document, ``f`` user id, ``g`` item id -- all int64 -- and ``h`` the rating) fully padded in host RAM and
turns a slice of each into a device ``LongTensor`` per batch (data_fast.py:99-109): 24 KB per rating cross
PCIe although ~60 % of every document is the padding token.

``RaggedReader`` takes the same arrays, keeps each document as int32 tokens up to its trailing padding run
(pinned host memory), copies only those per batch and rebuilds the identical padded int64 tensors on the
device (``r4r_docs_expand``), so the models see exactly what the reference's reader would have produced.
It honours the reader protocol ``train()`` / ``evaluate()`` use: ``iter(eval=False)`` yielding
``([a, b, c, d, e, f, g], h)``, ``len(reader)`` = number of batches, a short last batch.
"""
import ctypes
from typing import Dict, Optional

import numpy as np
import torch

from ._lib import call

PAD_ID = 0          # data.py:198-199 pads documents with token 0


def _vp(t):
    return ctypes.c_void_p(t.data_ptr() if t is not None else 0)


class RaggedDocs:
    """One padded document array ``[N, T]`` or ``[N, R, W]`` (NARRE: every review is its own padded row)
    stored as int32 tokens + int64 row offsets in (pinned) host memory."""

    def __init__(self, padded, pad_id: int = PAD_ID, pin: bool = True):
        arr = np.ascontiguousarray(np.asarray(padded))
        if arr.ndim < 2:
            raise ValueError("documents must be [N, T] or [N, R, W]")
        self.tail = tuple(arr.shape[1:])                    # per-rating shape
        self.T = int(arr.shape[-1])
        self.rows_per_item = int(np.prod(arr.shape[1:-1])) if arr.ndim > 2 else 1
        flat = arr.reshape(-1, self.T)
        if flat.size and (flat.min() < 0 or flat.max() >= 2 ** 31):
            raise ValueError("token ids must fit in int32")
        keep = flat != pad_id
        # length = index after the last non-pad token (interior pad tokens are ordinary tokens)
        lens = np.where(keep.any(axis=1), self.T - np.argmax(keep[:, ::-1], axis=1), 0).astype(np.int64)
        mask = np.arange(self.T)[None, :] < lens[:, None]
        tokens = torch.from_numpy(flat[mask].astype(np.int32))
        offsets = torch.from_numpy(np.concatenate([[0], np.cumsum(lens)]).astype(np.int64))
        self.pad_id = int(pad_id)
        self.tokens = tokens.pin_memory() if (pin and torch.cuda.is_available()) else tokens
        self.offsets = offsets                               # [rows + 1]; sliced + rebased per batch
        self.n_items = int(arr.shape[0])

    def batch_host(self, lo: int, hi: int, off_out=None):
        """(tokens int32 view, offsets int64 [rows+1] rebased to 0) of ratings lo..hi-1; the rebased
        offsets are written into ``off_out`` (a pinned staging buffer) when given."""
        r0, r1 = lo * self.rows_per_item, hi * self.rows_per_item
        o = self.offsets[r0:r1 + 1]
        base = int(o[0])
        if off_out is None:
            reb = o - base
        else:
            reb = torch.sub(o, base, out=off_out[: o.numel()])
        return self.tokens[base:int(o[-1])], reb

    def max_batch_tokens(self, bsz: int) -> int:
        rows = bsz * self.rows_per_item
        o = self.offsets
        if o.numel() - 1 <= rows:
            return int(o[-1])
        return int((o[rows:] - o[:-rows]).max())            # batches start at multiples of bsz, this bound holds for any start

    def to_padded(self, tokens, offsets):
        """Host-side inverse (tests): numpy padded array from a ragged slice."""
        rows = offsets.numel() - 1
        out = np.full((rows, self.T), self.pad_id, dtype=np.int64)
        t, o = tokens.numpy(), offsets.numpy()
        for r in range(rows):
            out[r, : o[r + 1] - o[r]] = t[o[r]:o[r + 1]]
        return out.reshape((-1,) + self.tail)


class _Slot:
    """Device staging of one batch: ragged token buffers and the padded tensors the model reads."""

    def __init__(self, reader, device):
        bsz = reader.bsz
        self.tok, self.off, self.off_host, self.pad = {}, {}, {}, {}
        self.used, self.n = False, 0
        for k, rd in reader.docs.items():
            rows = bsz * rd.rows_per_item
            self.tok[k] = torch.empty(max(1, rd.max_batch_tokens(bsz)), device=device, dtype=torch.int32)
            self.off[k] = torch.empty(rows + 1, device=device, dtype=torch.int64)
            self.off_host[k] = torch.empty(rows + 1, dtype=torch.int64).pin_memory()
            if not reader.native:
                self.pad[k] = torch.empty((bsz,) + rd.tail, device=device, dtype=torch.int64)
        self.small = {k: torch.empty((bsz,) + tuple(v.shape[1:]), device=device, dtype=v.dtype) for k, v in reader.small.items()}


class RaggedReader:
    DOC_KEYS = ("a", "d", "e")                               # this review, user document, item document
    SMALL_KEYS = ("b", "c", "f", "g", "h")

    def __init__(self, hyper_params: dict, arrays: Dict[str, Optional[np.ndarray]], device, slots: int = 2,
                 native: bool = False):
        """``arrays``: the datasets ``a..h`` of one split as numpy arrays (``None`` for slots the model
        does not read, as ``iter_simple`` does for MF: data.py:350-358)."""
        self.hyper_params = hyper_params
        self.bsz = int(hyper_params["batch_size"])
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("RaggedReader stages batches on a CUDA device (no CPU fallback)")
        # native=False: documents are expanded to the reference's padded int64 tensors (any consumer);
        # native=True: they are handed over as ops.RaggedIdx, which this package's models read directly
        self.native = bool(native)
        self.total = int(len(arrays["h"]))
        self.docs = {k: RaggedDocs(arrays[k]) for k in self.DOC_KEYS if arrays.get(k) is not None}
        self.small = {}
        for k in self.SMALL_KEYS:
            if arrays.get(k) is not None:
                t = torch.from_numpy(np.ascontiguousarray(arrays[k]))
                t = t.float() if k == "h" else t.long()      # FloatTensor(h), LongTensor(rest): data_fast.py:102-109
                self.small[k] = t.pin_memory()
        self.slots = [_Slot(self, self.device) for _ in range(slots)]
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self._free = [torch.cuda.Event() for _ in range(slots)]
        self._ready = [torch.cuda.Event() for _ in range(slots)]
        self.h2d_bytes_last = 0

    def __len__(self):
        return self.total // self.bsz + int(self.total % self.bsz > 0)

    def stage(self, batch: int, slot: int):
        """Enqueue on the copy stream the H2D of batch ``batch``'s ragged tokens / offsets / ids / ratings into
        slot ``slot``.  Returns (data, y, n_ratings); consumers must ``wait_ready(slot)`` first (it also
        expands the documents in padded mode) and ``release(slot)`` after their last use."""
        lo = (batch % len(self)) * self.bsz
        hi = min(self.total, lo + self.bsz)
        n = hi - lo
        S = self.slots[slot]
        nbytes = 0
        if S.used:
            self._ready[slot].synchronize()                  # the previous copies out of this slot's pinned staging are done
        S.used = True
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self._free[slot])
            for k, rd in self.docs.items():
                tok, off = rd.batch_host(lo, hi, S.off_host[k])
                S.tok[k][: tok.numel()].copy_(tok, non_blocking=True)
                S.off[k][: off.numel()].copy_(off, non_blocking=True)
                nbytes += tok.numel() * 4 + off.numel() * 8
            for k, v in self.small.items():
                S.small[k][:n].copy_(v[lo:hi], non_blocking=True)
                nbytes += (hi - lo) * v[0].numel() * v.element_size()
            self._ready[slot].record(self.copy_stream)
        S.n = n
        self.h2d_bytes_last = nbytes
        get = lambda d, k: d[k][:n] if k in d else None

        def doc(k):
            if k not in self.docs:
                return None
            if not self.native:
                return S.pad[k][:n]
            from .ops import RaggedIdx
            rd = self.docs[k]
            return RaggedIdx(S.tok[k], S.off[k][: n * rd.rows_per_item + 1], (n,) + rd.tail, rd.pad_id)

        data = [doc("a"), get(S.small, "b"), get(S.small, "c"), doc("d"), doc("e"), get(S.small, "f"), get(S.small, "g")]
        return data, S.small["h"][:n], n

    def wait_ready(self, slot: int, stream=None):
        """Make ``stream`` wait for the slot's copies and -- padded mode -- rebuild the padded int64 documents
        on it (the expansion runs on the consumer's stream, right before the step, so it never competes
        with the previous step's kernels for the SMs)."""
        stream = stream or torch.cuda.current_stream()
        stream.wait_event(self._ready[slot])
        if not self.native:
            S = self.slots[slot]
            for k, rd in self.docs.items():
                call("r4r_docs_expand", _vp(S.tok[k]), _vp(S.off[k]), S.n * rd.rows_per_item, rd.T, rd.pad_id, _vp(S.pad[k]),
                     ctypes.c_void_p(stream.cuda_stream))

    def release(self, slot: int, stream=None):
        self._free[slot].record(stream or torch.cuda.current_stream())

    def iter(self, eval=False):
        """Reader protocol of main.train / eval.evaluate.  Batch k+1 is staged while batch k is consumed."""
        nb, ns = len(self), len(self.slots)
        if nb == 0:
            return
        nxt = self.stage(0, 0)
        for b in range(nb):
            cur, slot = nxt, b % ns
            if b + 1 < nb:
                nxt = self.stage(b + 1, (b + 1) % ns)
            self.wait_ready(slot)
            yield cur[0], cur[1]
            self.release(slot)


# ------------------------------------------------------------------------------------------ review-level CSR
def review_list_side(ids, other, n_lists):
    """Review lists of one side (host, numpy).  ``ids[n]`` = the list (user or item) train review n belongs to,
    ``other[n]`` = the id on the opposite side.  Returns ``ptr`` [n_lists+1], ``rev`` (review ids in list order =
    train order within a list, as preprocess_random_split.py:213-218 appends them), ``nb`` (``other`` in list
    order: u_to_i_map / i_to_u_map) and ``rank[n]`` = position of review n inside its list
    (this_index_user_item[user][item][0 or 1])."""
    ids, other = np.asarray(ids, dtype=np.int64), np.asarray(other, dtype=np.int64)
    n = ids.shape[0]
    if n and (ids.min() < 0 or ids.max() >= n_lists):
        raise ValueError("review_list_side: id outside [0, %d)" % n_lists)
    order = np.argsort(ids, kind="stable")
    ptr = np.concatenate([[0], np.cumsum(np.bincount(ids, minlength=n_lists))]).astype(np.int64)
    rank = np.empty(n, dtype=np.int32)
    rank[order] = (np.arange(n) - ptr[ids[order]]).astype(np.int32)
    return ptr, order.astype(np.int32), other[order], rank


class ReviewStore:
    """The train reviews of a dataset, resident ONCE in device memory.

    Built from the train split in rating order -- review n is the text of train rating n, exactly how
    preprocess_random_split.py:207-219 fills ``user_reviews`` / ``item_reviews`` / ``this_index_user_item`` --
    as CSR: ``tok`` (int32, all reviews back to back), ``rev_off``, and per side the review lists
    ``ptr`` / ``rev`` / ``nb`` (``nb`` = ``u_to_i_map`` / ``i_to_u_map`` of data.py:36-63).  ``rank_user[n]`` /
    ``rank_item[n]`` are ``this_index_user_item[user][item]``: where review n sits in its user's / item's list.
    """

    def __init__(self, tok, rev_off, train_user, train_item, total_users: int, total_items: int, device):
        tok = np.ascontiguousarray(np.asarray(tok, dtype=np.int32))
        rev_off = np.ascontiguousarray(np.asarray(rev_off, dtype=np.int64))
        tu = np.asarray(train_user, dtype=np.int64)
        ti = np.asarray(train_item, dtype=np.int64)
        n = tu.shape[0]
        if rev_off.shape[0] != n + 1 or ti.shape[0] != n or int(rev_off[-1]) != tok.shape[0]:
            raise ValueError("ReviewStore: one review per train rating (rev_off must have n + 1 entries ending at len(tok))")
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("ReviewStore lives in device memory (no CPU fallback)")
        self.U, self.I, self.n = int(total_users), int(total_items), n
        dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(self.device)

        self.tok, self.rev_off = dev(tok), dev(rev_off)
        (ptr, rev, nb, rank_u), (iptr, irev, inb, rank_i) = review_list_side(tu, ti, self.U), review_list_side(ti, tu, self.I)
        self.u_ptr, self.u_rev, self.u_nb = dev(ptr), dev(rev), dev(nb)
        self.i_ptr, self.i_rev, self.i_nb = dev(iptr), dev(irev), dev(inb)
        self.train_user, self.train_item = dev(tu), dev(ti)
        self.rank_user, self.rank_item = dev(rank_u), dev(rank_i)
        self.bytes = sum(t.numel() * t.element_size() for t in (self.tok, self.rev_off, self.u_ptr, self.u_rev, self.u_nb,
                                                                self.i_ptr, self.i_rev, self.i_nb))


class CsrReader:
    """Reader protocol of ``main.train`` / ``eval.evaluate`` over a ``ReviewStore``: every batch's seven
    inputs are assembled on the device (``r4r_docs_assemble``); nothing but the batch bounds comes from the
    host.  ``train=True`` iterates the store's own train ratings (each rating's own review is left out of both
    documents and becomes ``this_reviews``); otherwise ``users`` / ``items`` / ``ratings`` are an evaluation split
    and ``this_tok`` / ``this_off`` its held-out reviews as CSR (``test_reviews`` of data.py:239-241)."""

    NBW = 10                                                            # data.py:277-282

    def __init__(self, hyper_params: dict, store: ReviewStore, ratings, train: bool, users=None, items=None,
                 this_tok=None, this_off=None, negs=None):
        self.hp, self.store, self.train = hyper_params, store, bool(train)
        self.bsz = int(hyper_params["batch_size"])
        dev = store.device
        to = lambda a, dt: torch.from_numpy(np.ascontiguousarray(np.asarray(a))).to(dt).to(dev)
        if train:
            self.users, self.items = store.train_user, store.train_item
        else:
            self.users, self.items = to(users, torch.int64), to(items, torch.int64)
        self.y = to(ratings, torch.float32)
        self.total = int(self.y.shape[0])
        if self.users.shape[0] != self.total:
            raise ValueError("CsrReader: one rating per (user, item)")
        self.this_tok = to(this_tok, torch.int32) if this_tok is not None else None
        self.this_off = to(this_off, torch.int64) if this_off is not None else None
        self.negs = None
        if negs is not None:
            self._prepare_negs(negs, users, items, this_tok, this_off)
        self.narre = hyper_params["model_type"] == "NARRE"
        self.simple = hyper_params["model_type"] in ("bias_only", "MF", "MF_dot", "NeuMF")      # data.py:33-34 iter_simple
        self.T = int(hyper_params.get("input_length", 1000))
        self.R, self.W = int(hyper_params.get("narre_num_reviews", 10)), int(hyper_params.get("narre_num_words", 100))

    def __len__(self):
        return self.total // self.bsz + int(self.total % self.bsz > 0)

    def batch(self, lo: int, hi: int):
        st, n, dev = self.store, hi - lo, self.store.device
        users, items, y = self.users[lo:hi], self.items[lo:hi], self.y[lo:hi]
        if self.simple:
            return [None, None, None, None, None, users, items], y
        shape = (n, self.R, self.W) if self.narre else (n, self.T)
        udoc = torch.empty(shape, device=dev, dtype=torch.int64)
        idoc = torch.empty(shape, device=dev, dtype=torch.int64)
        this = torch.empty(shape, device=dev, dtype=torch.int64)
        items_reviewed = torch.empty(n, self.NBW, device=dev, dtype=torch.int64)
        users_who_gave = torch.empty(n, self.NBW, device=dev, dtype=torch.int64)
        stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        mode = 1 if self.narre else 0
        sk_u = st.rank_user[lo:hi] if self.train else None
        sk_i = st.rank_item[lo:hi] if self.train else None
        eval_this = (not self.train) and self.this_tok is not None
        call("r4r_docs_assemble", _vp(st.tok), _vp(st.rev_off), _vp(st.u_ptr), _vp(st.u_rev), _vp(st.u_nb), st.U, _vp(users), _vp(sk_u), n,
             mode, self.T, self.R, self.W, st.I + 1, self.NBW, _vp(udoc), _vp(items_reviewed), _vp(this),
             _vp(self.this_tok if eval_this else None), _vp(self.this_off if eval_this else None), lo, stream)
        call("r4r_docs_assemble", _vp(st.tok), _vp(st.rev_off), _vp(st.i_ptr), _vp(st.i_rev), _vp(st.i_nb), st.I, _vp(items), _vp(sk_i), n,
             mode, self.T, self.R, self.W, st.U + 1, self.NBW, _vp(idoc), _vp(users_who_gave), _vp(None),
             _vp(None), _vp(None), 0, stream)
        return [this, users_who_gave, items_reviewed, udoc, idoc, users, items], y

    def iter(self, eval=False):
        for lo in range(0, self.total, self.bsz):
            yield self.batch(lo, min(self.total, lo + self.bsz))

    # ---- ranking candidates (data.py:375-447 iter_negs)
    def _prepare_negs(self, negs, users, items, this_tok, this_off):
        """``negs`` = (users [M], items [M, C]) with items[:, 0] the positive item (make_negative_sets.py writes
        ``negs[user] = [[positive], negatives]``).  The held-out review of (user, positive) is replicated for the C
        candidates, as the reference hands it to every candidate (data.py:389-397)."""
        dev = self.store.device
        nu = np.asarray(negs[0], dtype=np.int64)
        ni = np.asarray(negs[1], dtype=np.int64).reshape(len(nu), -1)
        C = ni.shape[1]
        rows = {}
        if users is not None and this_off is not None:
            rows = {(int(u), int(i)): n for n, (u, i) in enumerate(zip(np.asarray(users), np.asarray(items)))}
        tok_h = np.asarray(this_tok, dtype=np.int32) if this_tok is not None else np.zeros(0, dtype=np.int32)
        off_h = np.asarray(this_off, dtype=np.int64) if this_off is not None else np.zeros(1, dtype=np.int64)
        toks, lens = [], []
        for u, i in zip(nu.tolist(), ni[:, 0].tolist()):
            n = rows.get((u, i))
            rev = tok_h[off_h[n]:off_h[n + 1]] if n is not None else np.zeros(1, dtype=np.int32)     # [0] when missing
            for _ in range(C):
                toks.append(rev)
                lens.append(len(rev))
        to = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a)).to(dt).to(dev)
        self.negs = {"users": to(nu, torch.int64), "items": to(ni, torch.int64), "C": C,
                     "this_tok": to(np.concatenate(toks) if toks else np.zeros(0, dtype=np.int32), torch.int32),
                     "this_off": to(np.concatenate([[0], np.cumsum(lens)]).astype(np.int64), torch.int64)}

    def iter_negs(self, review):
        if self.negs is None:
            raise RuntimeError("this reader was built without ranking negatives (negs=...)")
        st, ng, dev = self.store, self.negs, self.store.device
        C, M = ng["C"], int(ng["users"].shape[0])
        for lo in range(0, M, self.bsz):
            hi = min(M, lo + self.bsz)
            n = hi - lo
            items = ng["items"][lo:hi].contiguous()                              # [n, C]
            users = ng["users"][lo:hi].repeat_interleave(C).view(n, C).contiguous()
            y = torch.zeros(n, device=dev, dtype=torch.float32)                    # "doesn't matter, only for ranking"
            if self.simple or not review:
                yield [None, None, None, None, None, users, items], y
                continue
            nc = n * C
            ids_u, ids_i = users.view(-1), items.view(-1)
            pos_i = items[:, 0].repeat_interleave(C).contiguous()
            shape = (nc, self.R, self.W) if self.narre else (nc, self.T)
            udoc, idoc, this, scratch = (torch.empty(shape, device=dev, dtype=torch.int64) for _ in range(4))
            items_reviewed = torch.empty(nc, self.NBW, device=dev, dtype=torch.int64)
            users_who_gave = torch.empty(nc, self.NBW, device=dev, dtype=torch.int64)
            stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
            mode = 1 if self.narre else 0
            none = _vp(None)
            # user side: whole document, the items the user reviewed, the held-out review of (user, positive)
            call("r4r_docs_assemble", _vp(st.tok), _vp(st.rev_off), _vp(st.u_ptr), _vp(st.u_rev), _vp(st.u_nb), st.U, _vp(ids_u), none, nc,
                 mode, self.T, self.R, self.W, st.I + 1, self.NBW, _vp(udoc), _vp(items_reviewed), _vp(this),
                 _vp(ng["this_tok"]), _vp(ng["this_off"]), lo * C, stream)
            # item side: every candidate's own document ...
            call("r4r_docs_assemble", _vp(st.tok), _vp(st.rev_off), _vp(st.i_ptr), _vp(st.i_rev), _vp(st.i_nb), st.I, _vp(ids_i), none, nc,
                 mode, self.T, self.R, self.W, st.U + 1, self.NBW, _vp(idoc), none, none, none, none, 0, stream)
            # ... but the users-who-reviewed list of the POSITIVE item for all of them (remove_overlap(u_r, i_r, u, i))
            call("r4r_docs_assemble", _vp(st.tok), _vp(st.rev_off), _vp(st.i_ptr), _vp(st.i_rev), _vp(st.i_nb), st.I, _vp(pos_i), none, nc,
                 mode, self.T, self.R, self.W, st.U + 1, self.NBW, _vp(scratch), _vp(users_who_gave), none, none, none, 0, stream)
            tail = shape[1:]
            yield [this.view((n, C) + tail), users_who_gave.view(n, C, self.NBW), items_reviewed.view(n, C, self.NBW),
                   udoc.view((n, C) + tail), idoc.view((n, C) + tail), users, items], y


# ------------------------------------------------------------------------------------------ reference pickles
def _load_pickle(path_without_ext):
    import pickle
    with open(path_without_ext + ".pkl", "rb") as f:                   # utils.py:23-25 load_obj
        return pickle.load(f)


def reference_pickles_to_arrays(data_dir: str) -> dict:
    """Host side (numpy only) of ``load_data``: reads the pickles the reference's preprocessing writes
    (data_scripts/preprocess_random_split.py:282-299: train / test / val = lists of [user, item, rating],
    user_reviews, this_index_user_item, test_reviews, num_users_items) and returns the CSR arrays
    ``ReviewStore`` / ``CsrReader`` take.  Train review n is the text of train rating n
    (``user_reviews[u][this_index_user_item[u][i][0]]``, :213-218); held-out reviews come from
    ``test_reviews[u][i]`` (:228-243), ``[0]`` when missing (data.py:239-241)."""
    train = _load_pickle(data_dir + "train")
    user_reviews = _load_pickle(data_dir + "user_reviews")
    this_index = _load_pickle(data_dir + "this_index_user_item")
    test_reviews = _load_pickle(data_dir + "test_reviews")
    num_users, num_items, num_words = _load_pickle(data_dir + "num_users_items")
    out = {"total_users": int(num_users), "total_items": int(num_items), "total_words": int(num_words)}

    def ratings(rows):
        a = np.asarray([[r[0], r[1]] for r in rows], dtype=np.int64).reshape(-1, 2)
        return a[:, 0].copy(), a[:, 1].copy(), np.asarray([r[2] for r in rows], dtype=np.float32)

    def csr(reviews):
        lens = np.fromiter((len(r) for r in reviews), dtype=np.int64, count=len(reviews))
        tok = np.fromiter((t for r in reviews for t in r), dtype=np.int64, count=int(lens.sum()))
        if tok.size and (tok.min() < 0 or tok.max() >= 2 ** 31):
            raise ValueError("token ids must fit in int32")
        return tok.astype(np.int32), np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)

    tu, ti, ty = ratings(train)
    out["train_user"], out["train_item"], out["train_y"] = tu, ti, ty
    out["tok"], out["rev_off"] = csr([user_reviews[int(u)][this_index[int(u)][int(i)][0]] for u, i in zip(tu, ti)])
    import os
    if os.path.exists(data_dir + "negs.pkl"):                    # data_scripts/make_negative_sets.py: negs[user] = [[positive], negatives]
        negs = _load_pickle(data_dir + "negs")
        out["negs_users"] = np.asarray(list(negs), dtype=np.int64)
        out["negs_items"] = np.asarray([list(negs[u][0]) + list(negs[u][1]) for u in negs], dtype=np.int64).reshape(len(negs), -1)
    for split in ("test", "val"):
        u, i, y = ratings(_load_pickle(data_dir + split))
        held = [test_reviews.get(int(a), {}).get(int(b), [0]) for a, b in zip(u, i)]
        out[split + "_user"], out[split + "_item"], out[split + "_y"] = u, i, y
        out[split + "_tok"], out[split + "_off"] = csr(held)
    return out


def load_data(hyper_params: dict, device="cuda"):
    """``data.load_data(hyper_params)`` (data.py:449-482) over device-resident reviews: returns
    ``(train_reader, test_reader, val_reader, hyper_params)`` with ``total_users / total_items / total_words``
    filled in like the reference (:468-470).  ``negs.pkl``, when present, gives the test reader ``iter_negs``."""
    a = reference_pickles_to_arrays(hyper_params["data_dir"])
    for k in ("total_users", "total_items", "total_words"):
        hyper_params[k] = a[k]
    store = ReviewStore(a["tok"], a["rev_off"], a["train_user"], a["train_item"], a["total_users"], a["total_items"], device)
    train = CsrReader(hyper_params, store, a["train_y"], train=True)
    negs = (a["negs_users"], a["negs_items"]) if "negs_users" in a else None
    test, val = (CsrReader(hyper_params, store, a[s + "_y"], train=False, users=a[s + "_user"], items=a[s + "_item"],
                           this_tok=a[s + "_tok"], this_off=a[s + "_off"], negs=negs if s == "test" else None)
                 for s in ("test", "val"))
    return train, test, val, hyper_params
