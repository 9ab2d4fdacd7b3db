"""Opt-in trainable word table (NOT reference behaviour -- the reference freezes the table, so
the checker is autograd on the oracle with the table marked trainable)."""
import os
import pickle
import tempfile

import pytest
import torch

from tests.helpers import assert_close

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mode,tol", [("exact", 1e-5), ("f16", 2e-3)])
@pytest.mark.parametrize("N,T,E,V,Fn", [(6, 20, 12, 40, 100), (9, 333, 300, 900, 100), (40, 50, 7, 30, 100), (17, 64, 64, 100, 37),
                                        (3, 5, 1000, 9, 8)])
def test_word_table_gradient_vs_oracle_autograd(mode, tol, N, T, E, V, Fn):
    from oracle import r4r_oracle as O
    from reviews4rec_b200 import ops
    from tests.test_gpu_kernels import _conv_case, gen
    table, idx, w, b = _conv_case(60 + N, N, T, E, V, Fn)
    g = torch.randn(N, Fn, generator=gen(4))
    tc = table.cuda().requires_grad_(True)
    wc, bc = w.cuda().requires_grad_(True), b.cuda().requires_grad_(True)
    ops.conv_pool(idx.cuda(), tc, wc, bc, mode=mode).backward(g.cuda())
    dt = {"exact": torch.float32, "f16": torch.float16}[mode]
    t_r = table.to(dt).float().requires_grad_(True)           # the forward the kernel evaluated (rounded operands in f16 mode)
    w_r = w.to(dt).float().requires_grad_(True)
    b_r = b.clone().requires_grad_(True)
    pooled, _ = O.conv_pool(O.word_gather(t_r, idx), w_r, b_r)
    pooled.backward(g)
    assert tc.grad.shape == table.shape
    assert_close(tc.grad, t_r.grad, rtol=max(tol, 1e-4), atol=tol, msg="dTable")
    assert_close(wc.grad, w_r.grad, rtol=max(tol, 1e-4), atol=tol, msg="dW")
    # rows of tokens that never occur stay exactly zero
    unused = torch.ones(V, dtype=torch.bool)
    unused[idx.reshape(-1)] = False
    assert float(tc.grad[unused.cuda()].abs().max() if unused.any() else 0.0) == 0.0


def test_deepconn_with_trainable_word_table_steps_the_table():
    import reviews4rec_b200 as R
    from oracle import r4r_oracle as O
    from reviews4rec_b200 import ops
    from reviews4rec_b200.optim import FusedAdam
    V, E, L, U, I, B, T = 80, 32, 6, 20, 15, 8, 90
    hp = {"model_type": "deepconn", "latent_size": L, "word_embed_size": E, "dropout": 0.0, "total_users": U, "total_items": I,
          "lr": 0.002, "weight_decay": 1e-6, "train_word_table": True}
    P = O.init_params(hp, V, seed=3)
    tmp = tempfile.mkdtemp()
    with open(os.path.join(tmp, "word2vec.pkl"), "wb") as f:
        pickle.dump(torch.zeros(V, E).tolist(), f, 2)
    hp["data_dir"] = tmp
    g = torch.Generator().manual_seed(1)
    ri = lambda hi, *s: torch.randint(0, hi, s, generator=g, dtype=torch.int64)
    data = [None, None, None, ri(V, B, T), ri(V, B, T), ri(U + 1, B), ri(I + 1, B)]
    y = torch.randint(1, 6, (B,), generator=g).float()
    ops.set_conv_mode("exact")
    model = R.DeepCoNN(hp)
    assert model.word2vec.weight.requires_grad
    model.load_state_dict(P)
    model = model.cuda().train()
    opt = FusedAdam(model.parameters(), lr=hp["lr"], weight_decay=hp["weight_decay"])
    before = model.word2vec.weight.detach().clone()
    out = model([d if d is None else d.cuda() for d in data])
    R.MSELoss(hp)(out, y.cuda()).backward()
    # oracle: same forward with the table as a leaf
    Q = {k: v.clone().requires_grad_(True) for k, v in P.items()}
    ref = O.deepconn_forward(Q, data, hp, train=True)
    O.mse(ref, y).backward()
    assert_close(model.word2vec.weight.grad, Q["word2vec.weight"].grad, rtol=1e-4, atol=1e-6, msg="word2vec.weight.grad")
    opt.step()
    torch.cuda.synchronize()
    moved = (model.word2vec.weight.detach() != before).any(dim=1)
    used = torch.zeros(V, dtype=torch.bool)
    used[data[3].reshape(-1)] = True
    used[data[4].reshape(-1)] = True
    assert bool(moved[used.cuda()].any())
