"""The reference's run drivers for this path, over the classes of this package: ``main_pytorch`` (main.py:342-399),
``main_NeuMF`` (main.py:289-340) and the ``main`` dispatcher (main.py:401-431).  Same call order, same log text,
same return values; model types outside this path (Surprise, HFT, MPCN) raise."""
import time

import torch

from .eval import eval_ranking, evaluate
from .loss import MSELoss
from .readers import _load_pickle, load_data
from .train import train_complete
from .utils import log_end_epoch, xavier_init

REVIEW_FREE = ("bias_only", "MF", "MF_dot", "NeuMF")
OUT_OF_SCOPE = ("SVD", "kNN", "NMF", "SVD++", "baseline", "HFT", "MPCN")


def load_user_item_counts(hyper_params):
    """utils.py:31-34."""
    return _load_pickle(hyper_params["data_dir"] + "user_count"), _load_pickle(hyper_params["data_dir"] + "item_count")


def model_class(model_type):
    """main.py:349-352."""
    if model_type in ("deepconn", "deepconn++"):
        from .pytorch_models.DeepCoNN import DeepCoNN as Model
    elif model_type in ("transnet", "transnet++"):
        from .pytorch_models.TransNet import TransNet as Model
    elif model_type == "NARRE":
        from .pytorch_models.NARRE import NARRE as Model
    elif model_type in ("bias_only", "MF", "MF_dot"):
        from .pytorch_models.MF import MF as Model
    else:
        raise ValueError("model_type %r is not part of the rating-prediction path" % (model_type,))
    return Model


def _final_metrics(hyper_params, model, test_reader, user_count, item_count, review, start_time):
    criterion = MSELoss(hyper_params)
    metrics, user_map, item_map = evaluate(model, criterion, test_reader, hyper_params, user_count, item_count, review=review)
    if getattr(test_reader, "negs", None) is not None:          # HR@1 needs the sampled negatives (data.py:375-447)
        metrics.update(eval_ranking(model, test_reader, hyper_params, review=review))
    log_end_epoch(hyper_params, metrics, "final", time.time() - start_time, metrics_on="(TEST)")
    return metrics, user_map, item_map


def main_pytorch(hyper_params, gpu_id=None, device="cuda"):
    Model = model_class(hyper_params["model_type"])
    user_count, item_count = load_user_item_counts(hyper_params)
    review_based_model = hyper_params["model_type"] not in REVIEW_FREE
    train_reader, test_reader, val_reader, hyper_params = load_data(hyper_params, device)
    model = Model(hyper_params).to(device)
    xavier_init(model)                                          # main.py:377 (re-initialises the frozen word table too)
    start_time = time.time()
    model = train_complete(hyper_params, Model, train_reader, val_reader, user_count, item_count, model, review=review_based_model)
    return _final_metrics(hyper_params, model, test_reader, user_count, item_count, review_based_model, start_time)


def main_NeuMF(hyper_params, gpu_id=None, device="cuda"):
    from .pytorch_models.NeuMF import GMF, MLP, NeuMF
    user_count, item_count = load_user_item_counts(hyper_params)
    train_reader, test_reader, val_reader, hyper_params = load_data(hyper_params, device)
    start_time = time.time()
    initial_path = hyper_params["model_path"]
    pre = {}
    for suffix, cls in (("_gmf", GMF), ("_mlp", MLP)):          # pre-train both halves (main.py:302-319)
        hyper_params["model_path"] = initial_path + suffix
        m = cls(hyper_params).to(device)
        xavier_init(m)
        pre[suffix] = train_complete(hyper_params, cls, train_reader, val_reader, user_count, item_count, m)
    hyper_params["model_path"] = initial_path
    model = NeuMF(hyper_params).to(device)
    model.init(pre["_gmf"], pre["_mlp"])
    model = train_complete(hyper_params, NeuMF, train_reader, val_reader, user_count, item_count, model)
    return _final_metrics(hyper_params, model, test_reader, user_count, item_count, False, start_time)


def main(hyper_params, gpu_id=None, device="cuda"):
    if gpu_id is not None:
        torch.cuda.set_device(int(gpu_id))
    mt = hyper_params["model_type"]
    if mt in OUT_OF_SCOPE:
        raise ValueError("model_type %r is outside the rating-prediction training path this package implements" % (mt,))
    method = main_NeuMF if mt == "NeuMF" else main_pytorch
    metrics, user_count_mse_map, item_count_mse_map = method(hyper_params, gpu_id=gpu_id, device=device)
    return metrics
