"""CPU test of the host control flow of train.train_complete (main.py:73-136): epoch loop, best-on-validation
checkpoint, reload into a fresh Model.  The per-epoch work (train / evaluate, which need the CUDA kernels) is
replaced by stand-ins; only the driver logic is under test here."""
import torch


class Tiny(torch.nn.Module):
    def __init__(self, hyper_params):
        super().__init__()
        self.w = torch.nn.Parameter(torch.zeros(3))


def test_train_complete_control_flow(tmp_path, monkeypatch):
    import reviews4rec_b200.eval as ev
    import reviews4rec_b200.train as tr
    val_mse = iter([3.0, 1.5, 2.0, 1.5])                     # best at epoch 2; the later tie must not overwrite it
    calls = {"train": 0}

    def fake_train(model, criterion, optimizer, reader, hp):
        calls["train"] += 1
        with torch.no_grad():
            model.w.fill_(float(calls["train"]))             # the model after epoch k holds k
        return {"MSE": 9.0}

    def fake_evaluate(model, criterion, reader, hp, user_count, item_count, review):
        return {"MSE": next(val_mse)}, {}, {}

    monkeypatch.setattr(tr, "train", fake_train)
    monkeypatch.setattr(ev, "evaluate", fake_evaluate)
    hp = {"model_type": "deepconn", "epochs": 4, "lr": 0.002, "weight_decay": 1e-6, "dataset": "x",
          "log_file": str(tmp_path / "log.txt"), "model_path": str(tmp_path / "m.pt")}
    best = tr.train_complete(hp, Tiny, [1, 2, 3], [1], {}, {}, Tiny(hp), review=True)
    assert calls["train"] == 4 and isinstance(best, Tiny) and not best.training
    assert best.w.tolist() == [2.0, 2.0, 2.0]                # epoch 2's weights, reloaded from the checkpoint
    log = open(hp["log_file"]).read()
    assert log.count("| end of epoch") == 4 and "Number of train batches:    3" in log and "MSE = 1.5" in log


def test_main_drivers_control_flow(tmp_path, monkeypatch):
    """main_pytorch / main_NeuMF / main (main.py:289-431): call order and hand-offs, with the device work replaced
    by recorders."""
    import pickle
    import reviews4rec_b200.main as M
    for name in ("user_count", "item_count"):
        with open(tmp_path / (name + ".pkl"), "wb") as f:
            pickle.dump({1: 2}, f, 2)
    log = []
    monkeypatch.setattr(M, "load_data", lambda hp, device: ("TRAIN", "TEST", "VAL", hp))
    monkeypatch.setattr(M, "xavier_init", lambda m: log.append(("xavier", type(m).__name__)))
    monkeypatch.setattr(M, "model_class", lambda mt: Tiny)

    def fake_train_complete(hp, Model, tr_, va, uc, ic, model, review=True):
        log.append(("train_complete", Model.__name__, hp["model_path"], tr_, va, review))
        return model

    monkeypatch.setattr(M, "train_complete", fake_train_complete)
    monkeypatch.setattr(M, "evaluate", lambda model, crit, reader, hp, uc, ic, review: ({"MSE": 1.25}, {0: [1.0]}, {0: [1.0]}))
    hp = {"model_type": "deepconn", "data_dir": str(tmp_path) + "/", "model_path": str(tmp_path / "m.pt"),
          "log_file": str(tmp_path / "log.txt"), "lr": 0.1, "weight_decay": 0.0}
    assert M.main(dict(hp), gpu_id=None, device="cpu") == {"MSE": 1.25}
    assert log == [("xavier", "Tiny"), ("train_complete", "Tiny", hp["model_path"], "TRAIN", "VAL", True)]
    assert "| end of epoch final" in open(hp["log_file"]).read()
    import pytest
    with pytest.raises(ValueError):
        M.main(dict(hp, model_type="HFT"))

    # NeuMF: GMF and MLP pre-trained under <path>_gmf / <path>_mlp, fused, then NeuMF under <path>
    import reviews4rec_b200.pytorch_models.NeuMF as N
    log.clear()
    for cls in ("GMF", "MLP", "NeuMF"):
        monkeypatch.setattr(N, cls, type(cls, (Tiny,), {"init": lambda self, a, b: log.append(("init", type(a).__name__, type(b).__name__))}))
    M.main(dict(hp, model_type="NeuMF"), device="cpu")
    steps = [e for e in log if e[0] in ("train_complete", "init")]
    assert [s[1] for s in steps] == ["GMF", "MLP", "GMF", "NeuMF"] and steps[2] == ("init", "GMF", "MLP")
    assert steps[0][2].endswith("_gmf") and steps[1][2].endswith("_mlp") and steps[3][2] == hp["model_path"] and steps[3][5] is True


def test_config_surface_matches_reference():
    """reviews4rec_b200.hyper_params against strings produced by the unmodified reference hyper_params.py
    (tests/golden/common_paths.json, oracle/gen_golden_config.py)."""
    import json
    import os
    from reviews4rec_b200 import hyper_params as H
    from tests.helpers import GOLDEN
    g = json.load(open(os.path.join(GOLDEN, "common_paths.json")))
    assert H.default_hyper_params() == g["defaults"]
    hp = H.finalize(H.default_hyper_params())
    assert {k: hp[k] for k in g["default_derived"]} == g["default_derived"]
    for case in g["cases"]:
        assert H.get_common_path(case["hyper_params"]) == case["common_path"]
    hp = H.finalize(dict(H.default_hyper_params(), percent_reviews_to_keep=50))
    assert hp["data_dir"] == "data/InstantVideo/5_core/50_percent/"
    narre = dict(H.default_hyper_params(), model_type="NARRE")            # the reference raises KeyError here
    assert "_only_reviews_False_" in H.get_common_path(narre)
