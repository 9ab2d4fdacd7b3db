"""Kernel-level parity on the GPU: every C-ABI entry point against the CPU oracle / plain torch
on the same seeded inputs.  Integer/byte/index work is bit-exact; floating point uses the
tolerances written next to each check (north_star: ratings within 1e-4 relative)."""
import ctypes

import pytest
import torch

from tests.helpers import assert_close

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def R():
    import reviews4rec_b200 as pkg
    return pkg


@pytest.fixture(scope="module")
def O():
    from oracle import r4r_oracle
    return r4r_oracle


def gen(seed):
    return torch.Generator().manual_seed(seed)


# --------------------------------------------------------------------------- a4 gather
@pytest.mark.parametrize("V,E,shape", [(50, 12, (5, 20)), (1000, 300, (7, 333)), (257, 64, (3, 4, 50)), (33, 7, (9,)), (10, 300, (0, 5))])
def test_word_gather_bit_exact(R, V, E, shape):
    from reviews4rec_b200 import ops
    g = gen(1)
    table = torch.randn(V, E, generator=g)
    idx = torch.randint(0, V, shape, generator=g, dtype=torch.int64)
    out = ops.word_gather(table.cuda(), idx.cuda())
    ref = table.index_select(0, idx.reshape(-1)).reshape(*shape, E)
    assert out.shape == ref.shape
    assert torch.equal(out.cpu(), ref)                      # bit-exact


def test_word_gather_rejects_cpu(R):
    from reviews4rec_b200 import ops
    with pytest.raises(RuntimeError):
        ops.word_gather(torch.randn(4, 4), torch.zeros(2, dtype=torch.int64))


@pytest.mark.parametrize("mode,dt", [("f16", torch.float16), ("bf16", torch.bfloat16)])
def test_shadow_table(R, mode, dt):
    from reviews4rec_b200 import ops
    table = torch.randn(123, 300, generator=gen(2)).cuda()
    sh = ops.ShadowTable()
    t = sh.get(table, mode)
    assert t.shape == (124, 320) and t.dtype == dt              # V rows + the all-zero row the conv padding reads
    assert torch.equal(t[:123, :300].cpu(), table.cpu().to(dt))   # round-to-nearest-even, bit-exact
    assert float(t[:, 300:].abs().max()) == 0.0 and float(t[123].abs().max()) == 0.0
    assert sh.get(table, mode) is t                              # cached while the table is unchanged
    table.add_(1.0)
    assert sh.get(table, mode) is not t                          # version bump -> rebuilt


# --------------------------------------------------------------------------- a7-a10 id rows
@pytest.mark.parametrize("Rr,L,n", [(1000, 10, 517), (50, 1, 300), (12, 32, 64), (100002, 5, 4096)])
def test_rows_gather_scatter(R, Rr, L, n):
    from reviews4rec_b200 import ops
    g = gen(3)
    table = torch.randn(Rr, L, generator=g) if L > 1 else torch.randn(Rr, generator=g)
    ids = torch.randint(0, Rr, (n,), generator=g, dtype=torch.int64)
    ids[: n // 3] = Rr - 1                                       # hot row (NARRE pad id)
    tc = table.cuda().requires_grad_(True)
    out = ops.rows_gather(tc, ids.cuda())
    ref_t = table.clone().requires_grad_(True)
    ref = ref_t[ids]
    assert torch.equal(out.detach().cpu(), ref.detach())        # bit-exact gather
    go = torch.randn(ref.shape, generator=g)
    out.backward(go.cuda())
    ref.backward(go)
    assert_close(tc.grad, ref_t.grad, rtol=1e-5, atol=1e-5, msg="dense id grad")


# --------------------------------------------------------------------------- a5 conv + pool
def _conv_case(seed, N, T, E, V, Fn=100, pad_tail=True):
    g = gen(seed)
    table = torch.randn(V, E, generator=g) * 0.5
    idx = torch.randint(0, V, (N, T), generator=g, dtype=torch.int64)
    if pad_tail and N > 1:
        idx[0, T // 3:] = 0                                      # padded tail: repeated windows -> exact ties
        idx[1, :] = 0
    w = torch.randn(Fn, 1, 3, E, generator=g) * (1.0 / (3 * E) ** 0.5)
    b = torch.randn(Fn, generator=g) * 0.1
    return table, idx, w, b


def _check_argmax(O, table, idx, w, b, pooled, arg, tol):
    """argmax must point at a position whose recomputed activation equals the pooled value."""
    x = O.word_gather(table.double(), idx)
    y = torch.nn.functional.conv2d(x.unsqueeze(1), w.double(), b.double(), padding=(2, 0)).squeeze(-1).relu()  # [N,F,T+2]
    at = y.gather(2, arg.long().unsqueeze(-1)).squeeze(-1)
    assert int(arg.min()) >= 0 and int(arg.max()) < idx.shape[1] + 2
    assert_close(at, y.max(dim=2).values, rtol=tol, atol=tol, msg="activation at argmax")
    assert_close(pooled, y.max(dim=2).values, rtol=tol, atol=tol, msg="pooled")


@pytest.mark.parametrize("N,T,E,V,Fn", [(5, 20, 12, 40, 100), (3, 200, 64, 500, 100), (2, 1000, 300, 2000, 100),
                                        (4, 7, 5, 11, 100), (1, 1, 16, 9, 100), (3, 130, 30, 77, 37)])
def test_conv_pool_exact(R, O, N, T, E, V, Fn):
    from reviews4rec_b200 import ops
    table, idx, w, b = _conv_case(10 + N, N, T, E, V, Fn)
    pooled, arg = ops.conv_pool_forward(idx.cuda(), table.cuda(), w.cuda(), b.cuda(), "exact")
    ref_pooled, ref_arg = O.conv_pool(O.word_gather(table, idx), w, b)
    assert_close(pooled, ref_pooled, rtol=1e-5, atol=1e-5, msg="pooled vs oracle")
    _check_argmax(O, table, idx, w, b, pooled.cpu().double(), arg.cpu(), 1e-5)
    # first-max rule on exact ties: the all-pad document repeats one window from position 2 on
    if N > 1 and T >= 4:
        live = ref_pooled[1] > 0
        assert torch.equal(arg.cpu()[1][live].long(), ref_arg[1][live])
    # determinism: same launch twice -> identical bits
    pooled2, arg2 = ops.conv_pool_forward(idx.cuda(), table.cuda(), w.cuda(), b.cuda(), "exact")
    assert torch.equal(pooled, pooled2) and torch.equal(arg, arg2)


@pytest.mark.parametrize("mode,dt", [("f16", torch.float16), ("bf16", torch.bfloat16)])
@pytest.mark.parametrize("N,T,E,V,Fn", [(5, 20, 12, 40, 100), (3, 200, 64, 500, 100), (4, 1000, 300, 2000, 100),
                                        (300, 130, 30, 77, 37), (2, 1, 16, 9, 100), (3, 127, 300, 100, 64)])
def test_conv_pool_tensor_core(R, O, mode, dt, N, T, E, V, Fn):
    """tcgen05 path: operands rounded to fp16/bf16, fp32 accumulation.  Against the oracle run on the
    SAME rounded operands the only difference is summation order (tolerance 2e-4 abs on O(1) values)."""
    from reviews4rec_b200 import ops
    table, idx, w, b = _conv_case(20 + N, N, T, E, V, Fn)
    pooled, arg = ops.conv_pool_forward(idx.cuda(), table.cuda(), w.cuda(), b.cuda(), mode)
    torch.cuda.synchronize()
    t_r, w_r = table.to(dt).float(), w.to(dt).float()
    ref_pooled, _ = O.conv_pool(O.word_gather(t_r, idx), w_r, b)
    assert_close(pooled, ref_pooled, rtol=2e-4, atol=2e-4, msg="pooled vs oracle on rounded operands")
    _check_argmax(O, t_r, idx, w_r, b, pooled.cpu().double(), arg.cpu(), 2e-4)
    # and within half-precision distance of the exact fp32 result
    ex_pooled, _ = O.conv_pool(O.word_gather(table, idx), w, b)
    tol = 2e-2 if mode == "bf16" else 3e-3
    assert_close(pooled, ex_pooled, rtol=tol, atol=tol, msg="pooled vs fp32 oracle")


def _ragged_docs(seed, N, T, V):
    """Documents whose tails repeat one token (id 0 or any other id) for run lengths that straddle
    every boundary the work plan cares about: 0..5 rows, whole document, tile edges."""
    g = gen(seed)
    idx = torch.randint(1, V, (N, T), generator=g, dtype=torch.int64)
    runs = [0, 1, 2, 3, 4, 5, T, T - 1, T - 2, T // 2] + [T - s for s in (253, 254, 255, 256, 257, 258, 509, 510, 511, 512, 513) if s < T]
    for n in range(N):
        r = runs[n] if n < len(runs) else int(torch.randint(0, T + 1, (1,), generator=g))
        tok = 0 if n % 3 else int(torch.randint(1, V, (1,), generator=g))
        if r > 0:
            idx[n, T - r:] = tok
    return idx


@pytest.mark.parametrize("N,T", [(40, 1000), (25, 600), (12, 257), (9, 300), (1000, 1000)])
def test_doc_plan_kernel(R, N, T):
    """r4r_doc_plan: doc_len = min(T, start of the trailing run + 3); doc_order = the STABLE permutation by
    decreasing length class -- 256 classes of (T+2)/255 windows (documents of one class keep their batch order: no atomics,
    replays exactly)."""
    import ctypes
    from reviews4rec_b200 import _lib
    idx = _ragged_docs(5, N, T, 50)
    d = idx.cuda()
    doc_len = torch.empty(N, dtype=torch.int32, device="cuda")
    order = torch.empty(N, dtype=torch.int32, device="cuda")
    ws = torch.empty(_lib.lib.r4r_doc_plan_ws_bytes(N, T), dtype=torch.uint8, device="cuda")
    vp = lambda t: ctypes.c_void_p(t.data_ptr())
    _lib.call("r4r_doc_plan", vp(d), N, T, vp(doc_len), vp(order), vp(ws), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    want = []
    for n in range(N):
        row = idx[n].tolist()
        s = T - 1
        while s > 0 and row[s - 1] == row[T - 1]:
            s -= 1
        want.append(min(T, s + 3))
    assert doc_len.cpu().tolist() == want
    o = order.cpu().tolist()
    assert sorted(o) == list(range(N))
    cls = lambda n: ((want[n] + 2) * 255) // (T + 2)
    classes = [cls(n) for n in o]
    assert classes == sorted(classes, reverse=True)
    assert o == sorted(range(N), key=lambda n: -cls(n))                            # python's sort is stable


@pytest.mark.parametrize("mode", ["f16", "bf16"])
@pytest.mark.parametrize("N,T,E,V", [(160, 1000, 300, 500), (40, 600, 64, 90), (33, 257, 32, 50),
                                     # window streams: many documents per 256-window tile, partial last round of the deal
                                     (700, 37, 24, 60), (75, 300, 40, 80), (149, 6, 16, 20), (1000, 130, 16, 40)])
def test_conv_doc_plan_is_exact(R, O, mode, N, T, E, V):
    """Cutting documents to their informative prefix must not change a single bit of (pooled, argmax),
    and the arg-max must be the FIRST maximum like F.max_pool1d's."""
    from reviews4rec_b200 import ops
    g = gen(77)
    table = (torch.randn(V, E, generator=g) * 0.5).cuda()
    w = (torch.randn(100, 1, 3, E, generator=g) * (1.0 / (3 * E) ** 0.5)).cuda()
    b = (torch.randn(100, generator=g) * 0.1).cuda()
    idx = _ragged_docs(6, N, T, V).cuda()
    try:
        ops.set_doc_plan(False)
        p0, a0 = ops.conv_pool_forward(idx, table, w, b, mode)
        ops.set_doc_plan(True)
        p1, a1 = ops.conv_pool_forward(idx, table, w, b, mode)
    finally:
        ops.set_doc_plan(True)
    torch.cuda.synchronize()
    assert torch.equal(p0, p1), "pooled changed: max diff %.3e" % float((p0 - p1).abs().max())
    assert torch.equal(a0, a1), "argmax changed at %d entries" % int((a0 != a1).sum())
    dt = torch.float16 if mode == "f16" else torch.bfloat16
    t_r, w_r = table.cpu().to(dt).float(), w.cpu().to(dt).float()
    ref_pooled, ref_arg = O.conv_pool(O.word_gather(t_r, idx.cpu()), w_r, b.cpu())
    assert_close(p1, ref_pooled, rtol=2e-4, atol=2e-4, msg="pooled vs oracle on rounded operands")
    _check_argmax(O, t_r, idx.cpu(), w_r, b.cpu(), p1.cpu().double(), a1.cpu(), 2e-4)
    # where the maximum lies inside the repeated tail the first-max position must match the oracle's
    live = ref_pooled > 0
    agree = (a1.cpu().long() == ref_arg)[live].float().mean()
    assert float(agree) > 0.99, "argmax agreement with F.max_pool1d only %.4f" % float(agree)


@pytest.mark.parametrize("N,T,E,V,Fn", [(6, 20, 12, 40, 100), (9, 333, 300, 900, 100), (40, 50, 7, 30, 100), (17, 64, 64, 100, 37)])
def test_conv_wgrad(R, O, N, T, E, V, Fn):
    from reviews4rec_b200 import ops
    table, idx, w, b = _conv_case(30 + N, N, T, E, V, Fn)
    wc, bc = w.cuda().requires_grad_(True), b.cuda().requires_grad_(True)
    pooled = ops.conv_pool(idx.cuda(), table.cuda(), wc, bc, mode="exact")
    gp = torch.randn(N, Fn, generator=gen(5))
    pooled.backward(gp.cuda())
    wr, br = w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    ref, _ = O.conv_pool(O.word_gather(table, idx), wr, br)
    ref.backward(gp)
    assert_close(wc.grad, wr.grad, rtol=1e-4, atol=1e-5, msg="conv dW vs autograd")
    assert_close(bc.grad, br.grad, rtol=1e-4, atol=1e-5, msg="conv db vs autograd")


# --------------------------------------------------------------------------- heads
@pytest.mark.parametrize("mode,dt", [("f16", torch.float16), ("bf16", torch.bfloat16)])
@pytest.mark.parametrize("N,T,E,V,Fn", [(6, 20, 12, 40, 100), (9, 333, 300, 900, 100), (40, 50, 7, 30, 100), (17, 64, 64, 100, 37),
                                        (5, 40, 1000, 20, 8)])
def test_conv_wgrad_half_rows(R, O, mode, dt, N, T, E, V, Fn):
    """r4r_conv_wgrad_argmax_h == autograd of the conv run on the rounded table (the function the
    tensor-core forward evaluates), through the public autograd op."""
    from reviews4rec_b200 import ops
    table, idx, w, b = _conv_case(40 + N, N, T, E, V, Fn)
    g = torch.randn(N, Fn, generator=gen(3))
    wc, bc = w.cuda().requires_grad_(True), b.cuda().requires_grad_(True)
    pooled = ops.conv_pool(idx.cuda(), table.cuda(), wc, bc, mode=mode)
    pooled.backward(g.cuda())
    t_r = table.to(dt).float()
    w_r = w.to(dt).float().requires_grad_(True)          # forward operands are rounded; the gradient is w.r.t. them
    b_r = b.clone().requires_grad_(True)
    ref, _ = O.conv_pool(O.word_gather(t_r, idx), w_r, b_r)
    # use the kernel's own relu mask/argmax choice where the oracle's differs only by rounding: compare on the
    # entries whose pooled values agree
    ref.backward(g)
    assert_close(bc.grad, b_r.grad, rtol=2e-3, atol=2e-3, msg="db")
    assert_close(wc.grad, w_r.grad, rtol=2e-3, atol=2e-3, msg="dW")


@pytest.mark.parametrize("n,i,o", [(37, 100, 10), (300, 20, 10), (5, 10, 1), (1000, 64, 32), (3, 4, 4)])
def test_linear(R, n, i, o):
    from reviews4rec_b200 import ops
    g = gen(6)
    x, W, b = torch.randn(n, i, generator=g), torch.randn(o, i, generator=g) * 0.2, torch.randn(o, generator=g)
    xc, Wc, bc = [t.cuda().requires_grad_(True) for t in (x, W, b)]
    xr, Wr, br = [t.clone().requires_grad_(True) for t in (x, W, b)]
    y, yr = ops.linear(xc, Wc, bc), torch.nn.functional.linear(xr, Wr, br)
    assert_close(y, yr, rtol=1e-5, atol=1e-5)
    gy = torch.randn(n, o, generator=g)
    y.backward(gy.cuda()); yr.backward(gy)
    for a, r, nm in ((xc, xr, "dx"), (Wc, Wr, "dW"), (bc, br, "db")):
        assert_close(a.grad, r.grad, rtol=1e-4, atol=1e-4, msg=nm)


@pytest.mark.parametrize("n,nf,k", [(64, 20, 8), (513, 10, 8), (7, 64, 32), (128, 14, 8)])
def test_fm(R, O, n, nf, k):
    from reviews4rec_b200 import ops
    g = gen(7)
    x, V = torch.randn(n, nf, generator=g), torch.randn(nf, k, generator=g) * 0.3
    lw, lb = torch.randn(1, nf, generator=g) * 0.3, torch.randn(1, generator=g)
    cu = [t.cuda().requires_grad_(True) for t in (x, V, lw, lb)]
    rf = [t.clone().requires_grad_(True) for t in (x, V, lw, lb)]
    out = ops.fm(cu[0], cu[1], cu[2].view(-1), cu[3])
    ref = O.torch_fm(rf[0], rf[1], rf[2], rf[3])
    assert out.shape == ref.shape == (n, 1)
    assert_close(out, ref, rtol=1e-5, atol=1e-5)
    go = torch.randn(n, 1, generator=g)
    out.backward(go.cuda()); ref.backward(go)
    for a, r, nm in zip(cu, rf, ("dx", "dV", "dlin_w", "dlin_b")):
        assert_close(a.grad, r.grad, rtol=1e-4, atol=1e-4, msg=nm)


def test_mse(R, O):
    g = gen(8)
    out, y = torch.randn(777, generator=g) + 4, torch.randint(1, 6, (777,), generator=g).float()
    crit = R.MSELoss({})
    oc = out.cuda().requires_grad_(True)
    se = crit(oc, y.cuda(), return_mean=False)
    assert_close(se, O.mse(out, y, False), rtol=1e-6, atol=1e-7)
    crit(oc, y.cuda()).backward()
    orr = out.clone().requires_grad_(True)
    O.mse(orr, y).backward()
    assert_close(oc.grad, orr.grad, rtol=1e-6, atol=1e-8)


# --------------------------------------------------------------------------- a12 Adam
def test_fused_adam_matches_torch(R):
    from reviews4rec_b200.optim import FusedAdam
    g = gen(9)
    shapes = [(1000002,), (100, 1, 3, 300), (10, 100), (10,), (1,), (20, 8), (7, 3), (0,)]
    ps = [torch.randn(s, generator=g) for s in shapes]
    mine = [p.clone().cuda().requires_grad_(True) for p in ps]
    ref = [p.clone().cuda().requires_grad_(True) for p in ps]
    o1 = FusedAdam(mine, lr=0.002, weight_decay=1e-6)
    o2 = torch.optim.Adam(ref, lr=0.002, weight_decay=1e-6)
    for step in range(5):
        for i, (a, b) in enumerate(zip(mine, ref)):
            if i == 3 and step % 2 == 1:
                a.grad = None; b.grad = None                 # params without a grad are skipped, step not advanced
                continue
            gr = torch.randn(a.shape, generator=g).cuda()
            if i == 0:
                gr[1000:] = 0                                 # dense table grad with few touched rows
            a.grad = gr.clone(); b.grad = gr.clone()
        o1.step(); o2.step()
    for a, b, s in zip(mine, ref, shapes):
        assert_close(a, b, rtol=1e-6, atol=1e-7, msg="param %s" % (s,))
    # untouched rows of the big table still move (weight decay + bias-corrected moments): finding 5
    assert float((mine[0][5000:] - ps[0].cuda()[5000:]).abs().max()) > 0


@pytest.mark.parametrize("mode", ["exact", "f16"])
def test_conv_pool_empty_batch(R, mode):
    """Zero documents (an empty trailing batch): empty outputs, no launch, zero weight gradients."""
    from reviews4rec_b200 import ops
    g = gen(5)
    table = torch.randn(30, 16, generator=g).cuda()
    w = torch.randn(100, 1, 3, 16, generator=g).cuda().requires_grad_(True)
    b = torch.randn(100, generator=g).cuda().requires_grad_(True)
    idx = torch.zeros(0, 300, dtype=torch.int64, device="cuda")
    p, a = ops.conv_pool_forward(idx, table, w.detach(), b.detach(), mode)
    assert tuple(p.shape) == (0, 100) and tuple(a.shape) == (0, 100)
    out = ops.conv_pool(idx, table, w, b, mode=mode)
    out.sum().backward()
    assert float(w.grad.abs().max()) == 0.0 and float(b.grad.abs().max()) == 0.0


@pytest.mark.parametrize("N,T,E,V", [(40, 600, 64, 90), (64, 1000, 300, 500), (9, 20, 12, 40)])
def test_conv_refine_gives_the_fp32_value_of_the_selected_window(R, O, N, T, E, V):
    """'f16r' mode: arg-max from the tensor-core kernel, value re-evaluated in fp32 (r4r_conv_refine).  The pooled features
    must agree with the fp32 oracle far below the half-precision operand error, and the selected window must be (within
    that error) a maximiser of the fp32 conv."""
    from reviews4rec_b200 import ops
    g = gen(21)
    table = (torch.randn(V, E, generator=g) * 0.5)
    w = (torch.randn(100, 1, 3, E, generator=g) * (1.0 / (3 * E) ** 0.5))
    b = (torch.randn(100, generator=g) * 0.1)
    idx = _ragged_docs(8, N, T, V)
    pooled_ref, arg_ref = O.conv_pool(O.word_gather(table, idx), w, b)
    p_r, a_r = ops.conv_pool_forward(idx.cuda(), table.cuda(), w.cuda(), b.cuda(), "f16r")
    p_h, a_h = ops.conv_pool_forward(idx.cuda(), table.cuda(), w.cuda(), b.cuda(), "f16")
    assert torch.equal(a_r, a_h)                                  # same selection
    scale = float(pooled_ref.abs().max())
    err_r = float((p_r.cpu() - pooled_ref).abs().max())
    err_h = float((p_h.cpu() - pooled_ref).abs().max())
    # where the tensor cores selected the oracle's window the value is the fp32 one; a different window (a near tie the
    # half-precision operands resolved the other way) is below the fp32 maximum by less than the f16 operand error
    live = pooled_ref > 0
    same = (a_r.cpu().long() == torch.as_tensor(arg_ref).long()) & live
    assert float(same.sum()) > 0.95 * float(live.sum())
    assert float((p_r.cpu() - pooled_ref).abs()[same].max()) <= 2e-5 * scale + 1e-6
    assert bool((p_r.cpu() <= pooled_ref + 2e-5 * scale).all())   # never above the true maximum
    assert err_r <= err_h + 1e-6, (err_r, err_h)
    mean_r = float((p_r.cpu() - pooled_ref).abs().mean()); mean_h = float((p_h.cpu() - pooled_ref).abs().mean())
    assert mean_r < 0.1 * mean_h, (mean_r, mean_h)                # on average an order of magnitude closer than plain f16
    # gradients: fp32 rows of the selected windows (the exact-mode weight-gradient kernel)
    wc, bc = w.clone().cuda().requires_grad_(True), b.clone().cuda().requires_grad_(True)
    out = ops.conv_pool(idx.cuda(), table.cuda(), wc, bc, mode="f16r")
    gout = torch.randn(N, 100, generator=g).cuda()
    out.backward(gout)
    ref_w, ref_b = O.conv_wgrad_argmax(O.word_gather(table, idx), a_r.cpu().long(), gout.cpu(), p_r.cpu())
    assert_close(wc.grad.cpu().reshape(ref_w.shape), ref_w, rtol=1e-4, atol=1e-5, msg="dW at the selected windows")
    assert_close(bc.grad.cpu(), ref_b, rtol=1e-4, atol=1e-5, msg="db")
