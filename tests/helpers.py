"""Shared helpers for the parity tests (golden loading, tolerances)."""
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
MODEL_TYPES = ["deepconn", "deepconn++", "bias_only", "MF_dot", "MF", "NARRE", "transnet", "transnet++"]
DIM_KEYS = ("E", "T", "L", "V", "U", "I", "B", "NB", "R", "W", "NEIGH")

# north_star tolerance: ratings / train-loop MSE to 1e-4 relative; grads and post-Adam params
# use rel 1e-4 + abs 1e-5..1e-6 (SURVEY.md 8c "parity caveats": accumulation order differs).
RTOL = 1e-4


def load_golden(mt):
    z = np.load(os.path.join(GOLDEN, mt.replace("+", "p") + ".npz"))
    dims = dict(zip(DIM_KEYS, [int(x) for x in z["dims"]]))
    return z, dims


def golden_hp(mt, dims, dropout=0.0):
    return {"model_type": mt, "latent_size": dims["L"], "word_embed_size": dims["E"], "dropout": dropout,
            "total_users": dims["U"], "total_items": dims["I"], "lr": 0.002, "weight_decay": 1e-6,
            "batch_size": dims["B"]}


def golden_state(z, prefix, device="cpu", dtype=torch.float32):
    n = len(prefix) + 1
    return {k[n:]: torch.from_numpy(z[k]).to(device=device, dtype=dtype) for k in z.files if k.startswith(prefix + ".")}


def golden_data(z, prefix, device="cpu"):
    data = []
    for j in range(7):
        k = "%s.d%d" % (prefix, j)
        data.append(torch.from_numpy(z[k]).to(device) if k in z.files else None)
    return data


def golden_batches(z, dims, device="cpu"):
    return [(golden_data(z, "b%d" % b, device), torch.from_numpy(z["b%d.y" % b]).to(device)) for b in range(dims["NB"])]


def assert_close(a, b, rtol=RTOL, atol=1e-6, msg=""):
    a = torch.as_tensor(a).detach().double().cpu()
    b = torch.as_tensor(b).detach().double().cpu()
    assert a.shape == b.shape, "%s shape %s vs %s" % (msg, tuple(a.shape), tuple(b.shape))
    err = (a - b).abs()
    tol = atol + rtol * b.abs()
    bad = err > tol
    assert not bool(bad.any()), "%s: max abs err %.3e (tol %.3e) at %d/%d elems" % (
        msg, float(err.max()), float(tol[bad].min()) if bad.any() else 0.0, int(bad.sum()), err.numel())
