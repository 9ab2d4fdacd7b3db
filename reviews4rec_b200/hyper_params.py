"""The reference's config surface (hyper_params.py:3-95) without its import side effects.

The reference's file is a flat dict plus derived paths, computed -- and ``saved_logs/`` / ``saved_models/``
created in the current directory -- at import time.  ``default_hyper_params()`` returns the same defaults
(:50-80) and ``finalize()`` derives ``common_path`` / ``log_file`` / ``model_path`` / ``data_dir`` with the same
rules (:3-48, :82-95); directories are only created on request.  ``model_type == 'NARRE'`` reads
``hyper_params['only_reviews']``, which the reference's dict lacks (KeyError at import): here it
defaults to False."""
import os


def default_hyper_params() -> dict:
    return {
        "dataset": "InstantVideo", "k_core": 5, "percent_reviews_to_keep": 100,
        "weight_decay": float(1e-6), "lr": 0.002, "epochs": 2, "batch_size": 128, "shuffle_data_every_epoch": False,
        "latent_size": 10, "word_embed_size": 64, "input_length": 1000, "dropout": 0.6,
        "model_type": "bias_only", "lamda": 0.1, "latent_reg": 0.0,
        "narre_num_reviews": 10, "narre_num_words": 100,
    }


def get_common_path(hyper_params: dict) -> str:
    """hyper_params.py:3-48, same strings."""
    method, fm = hyper_params["model_type"], True
    if method == "deepconn++":
        method, fm = "deepconn", False
    mt = hyper_params["model_type"]
    p = str(method) + "_" + str(hyper_params["dataset"]) + "_" + str(hyper_params["k_core"]) + "_core_"
    if mt in ("MF", "MF_dot", "NeuMF"):
        p += "_latent_size_" + str(hyper_params["latent_size"])
    elif mt == "HFT":
        p += "_latent_size_" + str(hyper_params["latent_size"]) + "_percent_reviews_" + str(hyper_params["percent_reviews_to_keep"])
    elif mt in ("deepconn", "deepconn++", "transnet", "transnet++"):
        p += "_word_embed_size_" + str(hyper_params["word_embed_size"]) + "_latent_size_" + str(hyper_params["latent_size"])
        p += "_percent_reviews_" + str(hyper_params["percent_reviews_to_keep"]) + "_fm_" + str(fm)
    elif mt == "NARRE":
        p += "_num_reviews_" + str(hyper_params["narre_num_reviews"]) + "_num_words_" + str(hyper_params["narre_num_words"])
        p += "_word_embed_size_" + str(hyper_params["word_embed_size"]) + "_latent_size_" + str(hyper_params["latent_size"])
        p += "_only_reviews_" + str(hyper_params.get("only_reviews", False))
        p += "_percent_reviews_" + str(hyper_params["percent_reviews_to_keep"])
    elif mt == "MPCN":
        return p + "_latent_size_" + str(hyper_params["latent_size"]) + "_percent_reviews_" + str(hyper_params["percent_reviews_to_keep"])
    p += "_wd_" + str(hyper_params["weight_decay"]) + "_lr_" + str(hyper_params["lr"])
    p += "_dropout_" + str(hyper_params["dropout"]) + "_input_len_" + str(hyper_params["input_length"])
    return p


def finalize(hyper_params: dict, make_dirs: bool = False, data_root: str = "data/") -> dict:
    """hyper_params.py:82-95: derived paths (and, on request, the two output directories)."""
    common_path = get_common_path(hyper_params)
    hyper_params["common_path"] = common_path
    hyper_params["log_file"] = "saved_logs/" + common_path
    hyper_params["model_path"] = "saved_models/" + common_path
    if make_dirs:
        os.makedirs("saved_logs/", exist_ok=True)
        os.makedirs("saved_models/", exist_ok=True)
    d = data_root + hyper_params["dataset"] + "/" + str(hyper_params["k_core"]) + "_core/"
    if hyper_params["percent_reviews_to_keep"] != 100:
        d += str(hyper_params["percent_reviews_to_keep"]) + "_percent/"
    hyper_params["data_dir"] = d
    return hyper_params
