"""Synthetic Amazon-shaped batches in the layout the reference's readers yield
(data_fast.py:99-109 / data.py:293-301): ``([this_reviews, users_who_reviewed, reviewed_items,
user_reviews, item_reviews, user_id, item_id], y)`` with int64 ids and fp32 ratings.

Distributions follow SURVEY.md section 8(d): token ids Zipf(s=1.0) over ranks 1..V-1, document
lengths log-normal (median 300, sigma 1.0) clipped to [20, T] with the tail padded by id 0 (pad
tokens are real rows: data.py:198-199), user / item ids Zipf(s=0.8), ratings categorical
{1..5} with p = (.05, .05, .1, .25, .55).  numpy ``default_rng(seed)``; nothing is read from disk.
"""
import numpy as np
import torch

RATING_P = (0.05, 0.05, 0.10, 0.25, 0.55)


class _Zipf:
    """Inverse-CDF sampler of P(rank r) ~ r^-s over r = 1..n."""

    def __init__(self, n, s):
        w = np.arange(1, n + 1, dtype=np.float64) ** (-s)
        self.cdf = np.cumsum(w)
        self.cdf /= self.cdf[-1]

    def draw(self, rng, shape):
        u = rng.random(shape)
        return (np.searchsorted(self.cdf, u, side="left") + 1).astype(np.int64)


def _docs(rng, zipf, n, T, full_length=False):
    """[n,T] int64 token ids: Zipf tokens up to a log-normal length, then id 0.  ``full_length``: no padding at all
    (every document has T informative rows: the worst case for the conv kernel's padding-run shortcut)."""
    tok = zipf.draw(rng, (n, T))
    if full_length:
        return tok
    length = np.clip(np.exp(rng.normal(np.log(300.0), 1.0, n)), 20, T).astype(np.int64)
    tok[np.arange(T)[None, :] >= length[:, None]] = 0
    return tok


class SyntheticReader:
    """A pool of ``n_batches`` pre-generated batches, cycled by ``iter()``.

    ``device=None`` keeps the batches in (optionally pinned) host memory -- the end-to-end path
    copies them per step like the reference's reader does; a CUDA device keeps them resident.
    Slots the model never reads (``this_reviews`` and the neighbour lists for DeepCoNN) are None,
    exactly as the reference's ``iter_simple`` does for MF (data.py:350-358)."""

    def __init__(self, hyper_params, batch_size, n_batches, V, seed=1234, device=None, pin=False,
                 rank=0, model_type=None):
        mt = model_type or hyper_params["model_type"]
        rng = np.random.default_rng(seed + 7919 * rank)
        U, I = hyper_params["total_users"], hyper_params["total_items"]
        T = hyper_params.get("input_length", 1000)
        R, W = hyper_params.get("narre_num_reviews", 10), hyper_params.get("narre_num_words", 200)
        zt, zu, zi = _Zipf(V - 1, 1.0), _Zipf(U, 0.8), _Zipf(I, 0.8)
        full = bool(hyper_params.get("synthetic_full_length", False))      # bench.py --full-length (not SURVEY 8d's shape)
        _d = _docs
        _docs_ = lambda rng_, z_, n_, T_: _d(rng_, z_, n_, T_, full)
        self.batch_size, self.model_type = batch_size, mt
        self.batches = []

        def put(a, dtype):
            t = torch.from_numpy(np.ascontiguousarray(a)).to(dtype)
            if device is not None:
                return t.to(device)
            return t.pin_memory() if pin else t

        B = batch_size
        for _ in range(n_batches):
            uid, iid = zu.draw(rng, B) - 1, zi.draw(rng, B) - 1          # 0-based ids < U, I
            y = rng.choice(5, size=B, p=RATING_P).astype(np.float32) + 1.0
            this = nb_u = nb_i = None
            if mt in ("deepconn", "deepconn++"):
                ud, idoc = _docs_(rng, zt, B, T), _docs_(rng, zt, B, T)
            elif mt in ("transnet", "transnet++"):
                ud, idoc, this = _docs_(rng, zt, B, T), _docs_(rng, zt, B, T), _docs_(rng, zt, B, T)
            elif mt == "NARRE":
                ud = _docs_(rng, zt, B * R, W).reshape(B, R, W)
                idoc = _docs_(rng, zt, B * R, W).reshape(B, R, W)
                # neighbour ids, padded with the reference's pad id U+1 / I+1 (data.py:275-276)
                nb_u, nb_i = zu.draw(rng, (B, R)) - 1, zi.draw(rng, (B, R)) - 1
                k = rng.integers(1, R + 1, B)
                pad = np.arange(R)[None, :] >= k[:, None]
                nb_u[pad], nb_i[pad] = U + 1, I + 1
                # item side: users who reviewed the item; user side: items the user reviewed
                nb_u, nb_i = nb_u, nb_i
            else:                                                        # MF family
                ud = idoc = None
            conv = lambda a: None if a is None else put(a, torch.int64)
            data = [conv(this), conv(nb_u), conv(nb_i), conv(ud), conv(idoc), conv(uid), conv(iid)]
            self.batches.append((data, put(y, torch.float32)))

    def __len__(self):
        return len(self.batches)

    def iter(self, eval=False, steps=None):
        n = steps if steps is not None else len(self.batches)
        for i in range(n):
            yield self.batches[i % len(self.batches)]

    def bytes_per_batch(self):
        data, y = self.batches[0]
        return sum(t.numel() * t.element_size() for t in data if t is not None) + y.numel() * y.element_size()
