"""GMF / MLP / NeuMF: the oracle restatement against the unmodified reference (golden,
oracle/gen_golden_neumf.py) on CPU; the drop-in classes' state_dict contract on CPU."""
import os

import numpy as np
import pytest
import torch

from tests.helpers import GOLDEN, assert_close

NAMES = ["GMF", "MLP", "NeuMF"]


def load_neumf():
    z = np.load(os.path.join(GOLDEN, "neumf.npz"))
    L, U, I, B, NB = [int(x) for x in z["dims"]]
    hp = {"latent_size": L, "dropout": 0.0, "total_users": U, "total_items": I, "lr": 0.002, "weight_decay": 1e-6, "batch_size": B}
    return z, hp, NB


def state(z, name, which, device="cpu"):
    pre = "%s.%s." % (name, which)
    return {k[len(pre):]: torch.from_numpy(z[k]).to(device) for k in z.files if k.startswith(pre)}


def batches(z, NB, device="cpu"):
    t = lambda k: torch.from_numpy(z[k]).to(device)
    return [([None] * 5 + [t("b%d.d5" % b), t("b%d.d6" % b)], t("b%d.y" % b)) for b in range(NB)]


@pytest.mark.parametrize("name", NAMES)
def test_oracle_matches_reference_neumf(name):
    from oracle import r4r_oracle as O
    z, hp, NB = load_neumf()
    hp = dict(hp, model_type=name)
    P = state(z, name, "init")
    bs = batches(z, NB)
    rank = [None] * 5 + [torch.from_numpy(z["rank.d5"]), torch.from_numpy(z["rank.d6"])]
    assert_close(O.forward(P, bs[0][0], hp), z["%s.eval.b0" % name], atol=1e-6, msg="eval b0")
    assert_close(O.forward(P, rank, hp), z["%s.eval.rank" % name], atol=1e-6, msg="eval rank")
    metrics, _, _, _ = O.train_batches(P, bs, hp)
    assert abs(metrics["MSE"] - float(z["%s.metric.MSE" % name][0])) <= 1e-4
    for k, v in state(z, name, "final").items():
        assert_close(P[k], v, atol=2e-6, msg="%s final.%s" % (name, k))


def test_oracle_neumf_init_matches_reference():
    from oracle import r4r_oracle as O
    z, hp, NB = load_neumf()
    got = O.neumf_init(state(z, "GMF", "final"), state(z, "MLP", "final"), state(z, "NeuMF", "init"))
    want = state(z, "NeuMF", "init")                           # the reference ran NeuMF.init on the trained GMF / MLP
    for k in want:
        if k != "global_bias":
            assert torch.equal(got[k], want[k]), k


@pytest.mark.parametrize("name", NAMES)
def test_neumf_modules_keep_the_reference_state_dict(name):
    import reviews4rec_b200 as R
    z, hp, NB = load_neumf()
    model = getattr(R, name)(dict(hp, model_type=name))
    ref = state(z, name, "init")
    sd = model.state_dict()
    assert set(sd) == set(ref)
    for k in ref:
        assert tuple(sd[k].shape) == tuple(ref[k].shape), k
    model.load_state_dict(ref)
    with pytest.raises(RuntimeError):                          # no CPU fallback
        model([None] * 5 + [torch.zeros(2, dtype=torch.int64), torch.zeros(2, dtype=torch.int64)])


def test_neumf_init_fuses_the_pretrained_models_cpu():
    """NeuMF.init moves no data through kernels, so the fusion (NeuMF.py:93-112) is checkable on CPU."""
    import reviews4rec_b200 as R
    z, hp, NB = load_neumf()
    gmf, mlp, neu = R.GMF(dict(hp, model_type="GMF")), R.MLP(dict(hp, model_type="MLP")), R.NeuMF(dict(hp, model_type="NeuMF"))
    gmf.load_state_dict(state(z, "GMF", "final"))
    mlp.load_state_dict(state(z, "MLP", "final"))
    neu.init(gmf, mlp)
    want, sd = state(z, "NeuMF", "init"), neu.state_dict()
    for k in want:
        if k != "global_bias":
            assert torch.equal(sd[k], want[k]), k
