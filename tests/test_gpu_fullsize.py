"""Size-independent properties at BASELINE.json's full sizes (configs[1]: E=300, F=100, T=1000, V=50,001,
4096 ratings per step; the V=2,000,000 table of SURVEY.md 8d for the stand-alone gather): equalities the domain
guarantees, checked on whole 4096-document launches.  The comparison with the oracle at this shape (256 ratings,
which the CPU finishes in seconds) lives in tests/test_gpu_fullshape_parity.py."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _amazon_docs(N, T, V, seed):
    from reviews4rec_b200.synthetic import _Zipf, _docs
    return torch.from_numpy(_docs(np.random.default_rng(seed), _Zipf(V - 1, 1.0), N, T)).cuda()


@pytest.fixture(scope="module")
def problem():
    g = torch.Generator(device="cuda").manual_seed(0)
    V, E, F = 50001, 300, 100
    table = (torch.rand(V, E, device="cuda", generator=g) - 0.5) * 0.07
    w = (torch.rand(F, 1, 3, E, device="cuda", generator=g) - 0.5) * 0.15
    b = (torch.rand(F, device="cuda", generator=g) - 0.5) * 0.1
    return table, w, b, _amazon_docs(4096, 1000, V, 1)


@pytest.mark.parametrize("mode", ["f16", "bf16"])
def test_conv_batch_split_and_permutation_invariance(problem, mode):
    """Documents are independent: the launch over 4096 documents equals launches over any split / order,
    bit for bit (persistent CTA pairs, longest-first work order and the padding-run shortcut included)."""
    from reviews4rec_b200 import ops
    table, w, b, idx = problem
    sh = ops.ShadowTable()
    p, a = ops.conv_pool_forward(idx, table, w, b, mode, sh)
    p1, a1 = ops.conv_pool_forward(idx[:1500], table, w, b, mode, sh)
    p2, a2 = ops.conv_pool_forward(idx[1500:], table, w, b, mode, sh)
    assert torch.equal(p, torch.cat([p1, p2])) and torch.equal(a, torch.cat([a1, a2]))
    perm = torch.randperm(idx.shape[0], device="cuda", generator=torch.Generator(device="cuda").manual_seed(3))
    pp, ap = ops.conv_pool_forward(idx[perm].contiguous(), table, w, b, mode, sh)
    assert torch.equal(pp, p[perm]) and torch.equal(ap, a[perm])
    # every arg-max points inside the document's T+2 conv positions and pooled is post-ReLU
    assert int(a.min()) >= 0 and int(a.max()) < idx.shape[1] + 2 and float(p.min()) >= 0.0


@pytest.mark.parametrize("mode", ["f16", "bf16"])
def test_conv_is_deterministic_run_to_run(problem, mode):
    """The same launch 20 times, plus the split launches, must give identical bits every time.  Round 1 shipped a
    race in the CTA pair's shared-memory exchange (a remote barrier arrival overtook the loads it was meant to
    follow) that only showed with four TMEM accumulator buffers and many single-tile documents: this shape."""
    from reviews4rec_b200 import ops
    table, w, b, idx = problem
    sh = ops.ShadowTable()
    p0, a0 = ops.conv_pool_forward(idx, table, w, b, mode, sh)
    for r in range(20):
        if r % 4 == 3:
            cut = 300 + 173 * r
            p1, a1 = ops.conv_pool_forward(idx[:cut], table, w, b, mode, sh)
            p2, a2 = ops.conv_pool_forward(idx[cut:], table, w, b, mode, sh)
            p, a = torch.cat([p1, p2]), torch.cat([a1, a2])
        else:
            p, a = ops.conv_pool_forward(idx, table, w, b, mode, sh)
        bad = ((p.view(torch.int32) != p0.view(torch.int32)) | (a != a0)).nonzero()
        assert bad.numel() == 0, "run %d: %d entries differ, first (doc, filter) = %s" % (r, bad.shape[0], bad[0].tolist())


def test_conv_doc_plan_exact_at_full_size(problem):
    from reviews4rec_b200 import ops
    table, w, b, idx = problem
    try:
        ops.set_doc_plan(False)
        p0, a0 = ops.conv_pool_forward(idx, table, w, b, "f16")
        ops.set_doc_plan(True)
        p1, a1 = ops.conv_pool_forward(idx, table, w, b, "f16")
    finally:
        ops.set_doc_plan(True)
    assert torch.equal(p0, p1) and torch.equal(a0, a1)


@pytest.mark.parametrize("mode", ["f16", "exact"])
def test_wgrad_is_linear_in_the_upstream_gradient(problem, mode):
    """dW(g1 + 2 g2) == dW(g1) + 2 dW(g2) (the arg-max selection does not depend on g)."""
    from reviews4rec_b200 import ops
    table, w, b, idx = problem
    n = 4096 if mode == "f16" else 256                       # the fp32 conv is the slow strict-parity path
    idx = idx[:n]
    g = torch.Generator(device="cuda").manual_seed(5)
    g1, g2 = torch.randn(n, 100, device="cuda", generator=g), torch.randn(n, 100, device="cuda", generator=g)

    def grads(gout):
        wc, bc = w.clone().requires_grad_(True), b.clone().requires_grad_(True)
        ops.conv_pool(idx, table, wc, bc, mode=mode).backward(gout)
        return wc.grad, bc.grad

    (wa, ba), (wb, bb), (wc_, bc_) = grads(g1), grads(g2), grads(g1 + 2 * g2)
    scale = float(wc_.abs().max())
    assert float((wa + 2 * wb - wc_).abs().max()) <= 2e-5 * scale + 1e-6
    assert float((ba + 2 * bb - bc_).abs().max()) <= 2e-5 * float(bc_.abs().max()) + 1e-6


def test_word_gather_bit_exact_with_a_table_beyond_l2():
    """SURVEY.md 8d stand-alone gather variant: V = 2,000,000 rows x 300 fp32 = 2.4 GB cannot sit in L2."""
    from reviews4rec_b200 import ops
    V, E, n = 2_000_000, 300, 1 << 20
    g = torch.Generator(device="cuda").manual_seed(7)
    table = torch.rand(V, E, device="cuda", generator=g)
    idx = torch.randint(0, V, (n,), device="cuda", generator=g)
    idx[:4] = torch.tensor([0, V - 1, 0, V - 1], device="cuda")
    out = ops.word_gather(table, idx)
    assert torch.equal(out, table[idx])                      # torch's own gather as the checker
    del table, out


def test_fused_adam_dense_over_a_million_row_table():
    """Finding 5 at config-2 size: every row of user_bias [1,000,002] moves each step, like torch.optim.Adam."""
    from reviews4rec_b200.optim import FusedAdam
    g = torch.Generator(device="cuda").manual_seed(9)
    p0 = torch.rand(1_000_002, device="cuda", generator=g) + 0.5
    grad = torch.zeros_like(p0)
    grad[torch.randint(0, p0.numel(), (4096,), device="cuda", generator=g)] = 0.3     # sparse arrival, dense update
    a, b_ = torch.nn.Parameter(p0.clone()), torch.nn.Parameter(p0.clone())
    oa = FusedAdam([a], lr=0.002, weight_decay=1e-6)
    ob = torch.optim.Adam([b_], lr=0.002, weight_decay=1e-6)
    for _ in range(3):
        a.grad, b_.grad = grad.clone(), grad.clone()
        oa.step(); ob.step()
    assert int((a.detach() != p0).sum()) == p0.numel()        # all rows moved (weight decay through Adam)
    torch.testing.assert_close(a.detach(), b_.detach(), rtol=1e-6, atol=1e-7)
