"""Evaluation with the reference's interface (eval.py): ``evaluate`` (:11-62) returns the validation /
test MSE plus the train-count -> per-sample-MSE maps, ``eval_ranking`` (:64-92) the HR@1 of the true item
against the sampled negatives.  Same forward kernels in eval mode; the reference's per-sample Python loops,
which read every squared error back with its own device sync (``float(mse[batch])``, eval.py:52-53), are
replaced by device-side accumulation and ONE read-back at the end of the split."""
import ctypes

import numpy as np
import torch

from ._lib import call
from .ops import _need_cuda, _p, _stream

TRANSNET = ("transnet", "transnet++")


def evaluate(model, criterion, reader, hyper_params, user_count, item_count, review):
    model.eval()
    is_tn = hyper_params["model_type"] in TRANSNET
    se_parts, user_parts, item_parts = [], [], []
    right = conv = None
    total_batches = 0.0
    with torch.no_grad():
        for data, y in reader.iter(eval=True):
            output = model(data)
            if is_tn:
                se = criterion(output[0], y, return_mean=False)
                r, c = criterion(output[1], y).reshape(1), output[2].reshape(1)
                right, conv = (r, c) if right is None else (right + r, conv + c)
            else:
                se = criterion(output, y, return_mean=False)
            se_parts.append(se.reshape(-1))
            # copies: a reader may hand out views of staging buffers it reuses (RaggedReader's two slots)
            user_parts.append(data[5].reshape(-1).clone())
            item_parts.append(data[6].reshape(-1).clone())
            total_batches += 1.0
        if not se_parts:
            return {}, {}, {}
        se_dev = torch.cat(se_parts)
        tail = torch.stack([se_dev.double().sum().float()] + ([right[0], conv[0]] if is_tn else []))
        se = se_dev.cpu().numpy()                                   # the split's only device -> host reads
        tail = tail.cpu().numpy()
        users = torch.cat(user_parts).cpu().numpy()
        items = torch.cat(item_parts).cpu().numpy()
    metrics = {"MSE": round(float(tail[0]) / float(se.shape[0]), 4)}
    if is_tn:
        metrics["MSE_right"] = round(float(tail[1]) / total_batches, 4)
        metrics["MSE_transform"] = round(float(tail[2]) / total_batches, 4)
    user_count_mse_map, item_count_mse_map = {}, {}
    for u, i, v in zip(users.tolist(), items.tolist(), se.tolist()):  # eval.py:42-53, without the per-sample syncs
        cu = user_count.setdefault(u, 0)
        ci = item_count.setdefault(i, 0)
        user_count_mse_map.setdefault(cu, []).append(v)
        item_count_mse_map.setdefault(ci, []).append(v)
    evaluate.last_raw = {"se_sum": float(tail[0]), "n": int(se.shape[0])}
    return metrics, user_count_mse_map, item_count_mse_map


def rows_argmax(scores: torch.Tensor) -> torch.Tensor:
    """Index of the first largest score per row ([N, C] fp32 -> [N] int64)."""
    _need_cuda(scores)
    s = scores.detach().float().contiguous()
    idx = torch.empty(s.shape[0], device=s.device, dtype=torch.int64)
    if s.shape[0]:
        call("r4r_rows_argmax", _p(s), s.shape[0], s.shape[1], _p(idx), _stream())
    return idx


def eval_ranking(model, reader, hyper_params, review=False):
    is_tn = hyper_params["model_type"] in TRANSNET
    tops = []
    with torch.no_grad():
        for data, y in reader.iter_negs(review):
            output = model(data)
            if is_tn:
                output = output[0]
            tops.append(rows_argmax(output.reshape(int(y.shape[0]), -1)))
    if not tops:
        return {}
    top = torch.cat(tops).cpu().numpy()                               # one read-back
    total = float(top.shape[0])
    eval_ranking.last_top = top
    return {"HR@1": round(100.0 * float((top == 0).sum()) / total, 2)}
