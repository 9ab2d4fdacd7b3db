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
