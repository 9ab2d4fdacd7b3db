"""Host-side helpers with the reference's names and meaning (utils.py): pickle loading of the
word vectors, xavier re-initialisation, and the TransNet optimizer groups."""
import pickle

import torch


def load_obj(name):
    """utils.py:23-25 -- `name` has no extension; the file is `<name>.pkl`."""
    with open(name + ".pkl", "rb") as f:
        return pickle.load(f)


def xavier_init(model):
    """utils.py:65-68 -- xavier-uniform on every parameter with more than one dim, INCLUDING the
    frozen word table (SURVEY.md finding 2)."""
    for p in model.parameters():
        if p.dim() > 1:
            torch.nn.init.xavier_uniform_(p)


def init_transnet_optim(hyper_params, model, optim_cls=torch.optim.Adam):
    """utils.py:70-92 -- [source, source_fm(+id embeddings for transnet++), target, all]."""
    kw = dict(lr=hyper_params["lr"], weight_decay=hyper_params["weight_decay"])
    fm_params = list(model.source_fm.parameters())
    if hyper_params["model_type"] == "transnet++":
        fm_params += [model.user_embedding.weight, model.item_embedding.weight]
    return [optim_cls(model.source.parameters(), **kw), optim_cls(fm_params, **kw),
            optim_cls(model.target.parameters(), **kw), optim_cls(model.parameters(), **kw)]


def file_write(log_file, s, dont_print=False):
    """utils.py:36-40 -- append a line to the run's log file (and echo it)."""
    if not dont_print:
        print(s)
    with open(log_file, "a") as f:
        f.write(s + "\n")


def log_end_epoch(hyper_params, metrics, epoch, time_elapsed, metrics_on="(VAL)"):
    """utils.py:53-63 -- the reference's end-of-epoch banner, same text."""
    string2 = ""
    for m in metrics:
        string2 += " | " + m + " = " + str(metrics[m])
    string2 += " " + metrics_on
    ss = "-" * 89
    ss += "\n| end of epoch {} | time: {:5.2f}s".format(epoch, time_elapsed)
    ss += string2
    ss += "\n"
    ss += "-" * 89
    file_write(hyper_params["log_file"], ss)
