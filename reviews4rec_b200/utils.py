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
