"""Host-side mirrors of the reference's driver functions, exercised end to end on the GPU:
``readers.load_data`` (data.py:449-482), ``train.train_complete`` (main.py:73-136), the GMF / MLP / NeuMF classes
and the ranking-candidate reader."""
import numpy as np
import pytest
import torch

from tests.helpers import golden_batches, load_golden

from tests.test_docs_oracle import _write_reference_pickles, load_docs_golden
from tests.test_gpu_models import ListReader, build


@pytest.mark.gpu
def test_load_data_from_reference_pickles(tmp_path):
    """readers.load_data == data.load_data over device-resident reviews: the train reader yields the golden
    batches of the reference reader."""
    from reviews4rec_b200.readers import load_data
    z, hp, (U, I, V) = load_docs_golden("deepconn")
    _write_reference_pickles(str(tmp_path), z, U, I)
    hp = dict(hp, data_dir=str(tmp_path) + "/")
    train, test, val, hp2 = load_data(hp, "cuda")
    assert hp2["total_users"] == U and hp2["total_items"] == I and len(train) == int(z["train.nb"][0])
    for b, (data, y) in enumerate(train.iter()):
        for j, d in enumerate(data):
            assert np.array_equal(d.cpu().numpy(), z["train.b%d.d%d" % (b, j)]), (b, j)
    n_eval = sum(int(y.shape[0]) for _, y in test.iter()) + sum(int(y.shape[0]) for _, y in val.iter())
    assert n_eval == len(z["eval_y"])


@pytest.mark.gpu
def test_train_complete_keeps_the_best_validation_checkpoint(tmp_path):
    """main.train_complete's contract (main.py:73-136): epochs of train -> validate, best-on-validation state_dict
    saved and reloaded into a fresh Model returned in eval mode; the log file carries the epoch banners."""
    import reviews4rec_b200 as R
    from reviews4rec_b200.eval import evaluate
    from reviews4rec_b200.train import train_complete
    mt = "deepconn"
    z, dims = load_golden(mt)
    model, hp = build(mt, z, dims)
    hp.update(epochs=3, log_file=str(tmp_path / "log.txt"), model_path=str(tmp_path / "model.pt"), dataset="golden")
    batches = golden_batches(z, dims, "cuda")
    train_reader, val_reader = ListReader(batches[:2]), ListReader(batches[2:])
    best = train_complete(hp, R.DeepCoNN, train_reader, val_reader, {}, {}, model, review=True)
    assert isinstance(best, R.DeepCoNN) and not best.training and next(best.parameters()).is_cuda
    log = open(hp["log_file"]).read()
    assert log.count("| end of epoch") == 3 and "(VAL)" in log and "Number of train batches:    2" in log
    # the returned model is the checkpoint on disk, and its validation MSE is the smallest one logged
    saved = torch.load(hp["model_path"], map_location="cuda")
    for k, v in best.state_dict().items():
        assert torch.equal(v, saved[k])
    logged = [float(line.split("MSE = ")[1].split(" ")[0]) for line in log.splitlines() if "| end of epoch" in line]
    m, _, _ = evaluate(best, R.MSELoss(hp), val_reader, hp, {}, {}, True)
    assert abs(m["MSE"] - min(logged)) < 1e-4


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["GMF", "MLP", "NeuMF"])
def test_neumf_family_vs_reference(name):
    """GMF / MLP / NeuMF drop-ins against the reference's forward and 3-batch main.train run."""
    import reviews4rec_b200 as R
    from reviews4rec_b200.optim import FusedAdam
    from reviews4rec_b200.train import train
    from tests.helpers import assert_close
    from tests.test_neumf import batches, load_neumf, state
    z, hp, NB = load_neumf()
    hp = dict(hp, model_type=name)
    model = getattr(R, name)(hp)
    model.load_state_dict(state(z, name, "init"))
    model = model.cuda()
    bs = batches(z, NB, "cuda")
    rank = [None] * 5 + [torch.from_numpy(z["rank.d5"]).cuda(), torch.from_numpy(z["rank.d6"]).cuda()]
    model.eval()
    with torch.no_grad():
        assert_close(model(bs[0][0]), z["%s.eval.b0" % name], rtol=1e-4, atol=1e-6, msg="eval b0")
        assert_close(model(rank), z["%s.eval.rank" % name], rtol=1e-4, atol=1e-6, msg="eval rank")
    opt = FusedAdam(model.parameters(), lr=hp["lr"], weight_decay=hp["weight_decay"])
    metrics = train(model, R.MSELoss(hp), opt, ListReader(bs), hp)
    assert abs(metrics["MSE"] - float(z["%s.metric.MSE" % name][0])) <= 1e-4
    sd = model.state_dict()
    for k, v in state(z, name, "final").items():
        assert_close(sd[k], v, rtol=1e-4, atol=4e-6, msg="%s final.%s" % (name, k))


@pytest.mark.gpu
def test_neumf_init_fuses_the_pretrained_models():
    import reviews4rec_b200 as R
    from tests.test_neumf import load_neumf, state
    z, hp, NB = load_neumf()
    gmf, mlp, neu = R.GMF(dict(hp, model_type="GMF")), R.MLP(dict(hp, model_type="MLP")), R.NeuMF(dict(hp, model_type="NeuMF"))
    gmf.load_state_dict(state(z, "GMF", "final"))
    mlp.load_state_dict(state(z, "MLP", "final"))
    gmf, mlp, neu = gmf.cuda(), mlp.cuda(), neu.cuda()
    neu.init(gmf, mlp)
    want = state(z, "NeuMF", "init")
    sd = neu.state_dict()
    for k in want:
        if k != "global_bias":
            assert torch.equal(sd[k].cpu(), want[k]), k


@pytest.mark.gpu
@pytest.mark.parametrize("mt", ["deepconn", "NARRE"])
def test_device_ranking_candidates_reproduce_reference_iter_negs(mt):
    """CsrReader.iter_negs == the reference reader's iter_negs (data.py:375-447), golden from the unmodified
    reference: [bsz, 1+5, ...] inputs incl. the positive item's held-out review / reviewer list for all candidates."""
    from reviews4rec_b200.readers import CsrReader, ReviewStore
    z, hp, (U, I, V) = load_docs_golden(mt)
    store = ReviewStore(z["tok"], z["rev_off"], z["train_user"], z["train_item"], U, I, "cuda")
    reader = CsrReader(hp, store, z["eval_y"], train=False, users=z["eval_user"], items=z["eval_item"],
                       this_tok=z["eval_tok"], this_off=z["eval_off"], negs=(z["negs.users"], z["negs.items"]))
    got = list(reader.iter_negs(True))
    assert len(got) == int(z["negs.nb"][0])
    for b, (data, y) in enumerate(got):
        for j, d in enumerate(data):
            want = z["negs.b%d.d%d" % (b, j)]
            assert tuple(d.shape) == tuple(want.shape) and np.array_equal(d.cpu().numpy(), want), (b, j)
        assert np.array_equal(y.cpu().numpy(), z["negs.b%d.y" % b])
    simple = list(reader.iter_negs(False))
    assert simple[0][0][0] is None and np.array_equal(simple[0][0][6].cpu().numpy(), z["negs.b0.d6"])
