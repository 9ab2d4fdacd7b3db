"""The fused DeepCoNN / DeepCoNN++ head (r4r_deepconn_head_fwd / _bwd: fc + dropout + cat + FM or MLP + biases + squared
error in one kernel each way) against the op-by-op path, the reference's golden vectors, and -- with keep masks handed to
both sides -- the oracle's dropout arithmetic (common_pytorch_models.py:33-57, DeepCoNN.py:61-72, loss.py:7-11)."""
import pytest
import torch

from tests.helpers import assert_close, golden_batches, golden_data, load_golden

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mt", ["deepconn", "deepconn++"])
def test_fused_head_equals_op_by_op_path(mt):
    import reviews4rec_b200 as R
    from tests.test_gpu_models import build
    z, dims = load_golden(mt)
    data, y = golden_batches(z, dims, "cuda")[0]
    res = []
    for fused in (True, False):
        model, hp = build(mt, z, dims, mode="exact")
        model.fused_head = fused
        model.train()
        out = model(data)
        R.MSELoss(hp)(out, y).backward()
        res.append((out.detach(), {n: p.grad for n, p in model.named_parameters() if p.grad is not None}))
    (oa, ga), (ob, gb) = res
    assert_close(oa, ob, rtol=1e-5, atol=1e-6, msg="rating")
    assert_close(oa, z["train.out0"], rtol=1e-4, atol=1e-5, msg="rating vs reference")
    assert set(ga) == set(gb), set(ga) ^ set(gb)
    for k in ga:
        assert_close(ga[k], gb[k], rtol=1e-4, atol=1e-6, msg="grad " + k)
        assert_close(ga[k], z["grad." + k], rtol=1e-4, atol=1e-6, msg="grad vs reference " + k)


@pytest.mark.parametrize("mt", ["deepconn", "deepconn++"])
def test_forward_with_loss_equals_forward_plus_mse(mt):
    import reviews4rec_b200 as R
    from tests.test_gpu_models import build
    z, dims = load_golden(mt)
    data, y = golden_batches(z, dims, "cuda")[0]
    model, hp = build(mt, z, dims, mode="exact")
    model.train()
    se_sum = torch.zeros(1, device="cuda")
    out, se = model.forward_with_loss(data, y, se_sum)
    se.backward(torch.full_like(se, 1.0 / se.numel()))
    g1 = {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None}
    model.zero_grad()
    out2 = model(data)
    se2 = R.MSELoss(hp)(out2, y, return_mean=False)
    se2.mean().backward()
    assert_close(out, out2, rtol=1e-6, atol=1e-6, msg="rating")
    assert_close(se, se2, rtol=1e-5, atol=1e-6, msg="se")
    assert abs(float(se_sum) - float(se2.sum())) <= 1e-4 * float(se2.sum())
    for n, p in model.named_parameters():
        if p.grad is not None:
            assert_close(g1[n], p.grad, rtol=1e-4, atol=1e-6, msg="grad " + n)


@pytest.mark.parametrize("mt", ["deepconn", "deepconn++"])
def test_dropout_with_exported_masks_matches_the_oracle(mt):
    """Same keep masks on both sides (PyTorch's CPU RNG stream cannot be reproduced on the device, SURVEY.md section 7
    "Dropout"): ratings and every gradient agree with the oracle's x * mask / (1 - p) arithmetic."""
    from oracle import r4r_oracle as O
    import reviews4rec_b200 as R
    from tests.helpers import golden_hp, golden_state
    from tests.test_gpu_models import build
    z, dims = load_golden(mt)
    p, L, B = 0.5, dims["L"], dims["B"]
    data, y = golden_batches(z, dims, "cpu")[0]
    g = torch.Generator().manual_seed(9)
    keep = (torch.rand(B, 3 * L, generator=g) >= p)
    masks = {"user_conv": keep[:, :L].float(), "item_conv": keep[:, L:2 * L].float(), "final": keep[:, 2 * L:].float()}
    hp_o = golden_hp(mt, dims, dropout=p)
    out_ref, se_ref, grads = O.grads_of(golden_state(z, "init"), data, y, hp_o, train=True, masks=masks)
    model, hp = build(mt, z, dims, mode="exact", dropout=p)
    model.train()
    model._r4r_keep_masks = keep.to(torch.uint8).cuda().contiguous()
    out = model([None if d is None else d.cuda() for d in data])
    R.MSELoss(hp)(out, y.cuda()).backward()
    assert_close(out, out_ref, rtol=1e-4, atol=1e-5, msg="rating")
    for n, prm in model.named_parameters():
        if grads.get(n) is not None:
            assert_close(prm.grad, grads[n], rtol=1e-4, atol=1e-6, msg="grad " + n)
        else:
            assert prm.grad is None, n


def test_philox_dropout_statistics_and_stream():
    """In-kernel Philox dropout: the keep rate is 1 - p, masks differ between ratings, stay fixed within a step (forward
    and backward see the same bits) and change after a backward (the step counter advances)."""
    from reviews4rec_b200 import ops
    N, F, L, K, p = 4096, 100, 10, 8, 0.6
    dev = "cuda"
    pooled = torch.ones(N, F, device=dev)
    fc_w = torch.full((L, F), 1.0 / F, device=dev, requires_grad=True)            # latent = 1 before dropout
    fc_b = torch.zeros(L, device=dev)
    fm = (torch.zeros(2 * L, K, device=dev), torch.ones(1, 2 * L, device=dev), torch.zeros(1, device=dev))
    gb = torch.zeros(1, device=dev)
    step = torch.zeros(1, device=dev, dtype=torch.int32)

    def run():
        r, _ = ops.deepconn_head(pooled, pooled, (fc_w, fc_b), (fc_w, fc_b), 0, fm=fm, global_bias=gb, p=p, seed=1234, step=step)
        return r                                                                   # = (#kept of 2L) / (1 - p)

    r1 = run()
    kept = r1.detach() * (1 - p)
    assert abs(float(kept.mean()) / (2 * L) - (1 - p)) < 0.01                      # 81,920 Bernoulli draws
    assert float(kept.std()) > 1.0                                                 # ratings differ: masks are per rating
    assert torch.equal(run().detach(), r1.detach())                                # same step, same masks
    r1.sum().backward()
    assert int(step) == 1
    assert not torch.equal(run().detach(), r1.detach())                            # next step, new masks
    # eval mode (p = 0): no dropout at all
    r0, _ = ops.deepconn_head(pooled, pooled, (fc_w, fc_b), (fc_w, fc_b), 0, fm=fm, global_bias=gb, p=0.0)
    assert_close(r0, torch.full((N,), 2.0 * L), rtol=1e-5, atol=1e-5)
