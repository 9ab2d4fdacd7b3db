"""GPU tests of the row-sharded tables: the csrc/shard.cu kernels against the CPU
emulation with several virtual ranks on one device, the full sharded model path at world size 1 against
the reference's golden training run, and -- when the box has >= 2 GPUs -- a 2-rank NCCL run that must
reproduce the single-process reference at the same global batch."""
import os
import subprocess
import sys

import pytest
import torch

from tests.helpers import ROOT, assert_close, golden_batches, golden_hp, golden_state, load_golden
from tests.shard_emul import CpuKernels

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def S():
    from reviews4rec_b200 import sharded
    return sharded


@pytest.mark.parametrize("P,V,E", [(1, 37, 8), (2, 1001, 300), (3, 50, 12), (8, 50001, 64), (2, 1200000, 4)])
def test_word_lookup_kernels_virtual_ranks(S, P, V, E):
    g = torch.Generator().manual_seed(P)
    full = torch.randn(V, E, generator=g)
    cap = S.rows_local(V, P)
    shards = [S.shard_rows(full, r, P).cuda() for r in range(P)]
    idxs = [torch.randint(0, V, (5, 70), generator=g) for _ in range(P)]          # rank r's documents
    K, C = S.K, CpuKernels()
    reqs = []
    for r in range(P):
        flags = torch.zeros(V, dtype=torch.int32, device="cuda")
        req = torch.zeros(P, 1 + cap, dtype=torch.int64, device="cuda")
        K.mark(idxs[r].cuda(), V, flags)
        K.plan(flags, V, P, cap, req)
        f2, r2 = torch.zeros(V, dtype=torch.int32), torch.zeros(P, 1 + cap, dtype=torch.int64)
        C.mark(idxs[r], V, f2)
        C.plan(f2, V, P, cap, r2)
        assert int(flags.sum()) == 0
        for o in range(P):                                       # same SET of owner-local rows per owner; the order is free
            n = int(r2[o, 0])
            assert int(req[o, 0]) == n
            assert torch.equal(torch.sort(req[o, 1:1 + n].cpu()).values, r2[o, 1:1 + n])
        reqs.append(req)
    # "all-to-all" of the requests, serve, "all-to-all" of the rows, place by original id
    payloads = [torch.empty(P, cap, E, device="cuda") for _ in range(P)]
    for o in range(P):
        rreq = torch.stack([reqs[q][o] for q in range(P)])
        payload = torch.full((P, cap, E), float("nan"), device="cuda")
        K.serve(shards[o], rreq, P, cap, payload)
        for q in range(P):
            payloads[q][o] = payload[q]
    for r in range(P):
        cache = torch.full((V, E), float("nan"), device="cuda")
        K.place(payloads[r].view(P * cap, E), reqs[r], P, cap, cache, V)
        got = cache[idxs[r].cuda()]
        assert torch.equal(got.cpu(), full[idxs[r]])                                # bit-exact rows, read through the ORIGINAL ids
        touched = torch.zeros(V, dtype=torch.bool)
        touched[idxs[r].reshape(-1)] = True
        assert bool(torch.isnan(cache.cpu()[~touched]).all())                       # nothing else was written


@pytest.mark.parametrize("P,R,L,n", [(1, 9, 1, 20), (2, 1002, 10, 4096), (8, 100003, 5, 3000), (4, 7, 32, 257)])
def test_id_lookup_kernels_virtual_ranks(S, P, R, L, n):
    g = torch.Generator().manual_seed(R)
    full = torch.randn(R, L, generator=g)
    shards = [S.shard_rows(full, r, P).cuda() for r in range(P)]
    K = S.K
    cap = n
    ids = [torch.randint(0, R, (n,), generator=g) for _ in range(P)]
    for t in ids:
        t[: n // 3] = R - 1                                                          # hot pad-like row
    gouts = [torch.randn(n, L, generator=g) for _ in range(P)]
    reqs, poss = [], []
    for r in range(P):
        req = torch.empty(P, 1 + cap, dtype=torch.int64, device="cuda")
        pos = torch.empty(n, dtype=torch.int64, device="cuda")
        K.bucket(ids[r].cuda(), R, P, cap, req, pos)
        rq, ps = req.cpu(), pos.cpu()
        assert ps.unique().numel() == n                                              # every id got its own slot
        o, k = ps // cap, ps % cap
        assert torch.equal(o, ids[r] % P) and torch.equal(rq[o, 1 + k], ids[r] // P)
        assert torch.equal(rq[:, 0], torch.bincount(ids[r] % P, minlength=P))
        reqs.append(req)
        poss.append(pos)
    rreqs = [torch.stack([reqs[q][o] for q in range(P)]).contiguous() for o in range(P)]
    recv = [torch.empty(P, cap, L, device="cuda") for _ in range(P)]
    for o in range(P):
        payload = torch.empty(P, cap, L, device="cuda")
        K.serve(shards[o], rreqs[o], P, cap, payload)
        for q in range(P):
            recv[q][o] = payload[q]
    sends = []
    for r in range(P):
        out = torch.empty(n, L, device="cuda")
        K.gather(recv[r].view(P * cap, L), poss[r], out)
        assert torch.equal(out.cpu(), full[ids[r]])
        send = torch.zeros(P * cap, L, device="cuda")
        K.scatter_unique(gouts[r].cuda(), poss[r], send)
        sends.append(send.view(P, cap, L))
    ref = torch.zeros(R, L)
    for r in range(P):
        ref.index_add_(0, ids[r], gouts[r])
    ref *= 0.25
    for o in range(P):
        grads = torch.stack([sends[q][o] for q in range(P)]).contiguous().view(P * cap, L)
        gt = torch.zeros_like(shards[o])
        K.scatter_owner(grads, rreqs[o], P, cap, gt, 0.25)
        assert_close(gt, S.shard_rows(ref, o, P), rtol=1e-5, atol=1e-5, msg="owner %d grads" % o)


@pytest.mark.parametrize("mt", ["deepconn", "deepconn++", "NARRE", "transnet++", "MF_dot"])
def test_sharded_model_world1_matches_reference_training(S, mt):
    """The whole sharded path (plan / serve / remap / bucket / owner scatter through the real kernels,
    world size 1) must reproduce the reference's training run."""
    import reviews4rec_b200 as R
    from reviews4rec_b200.optim import FusedAdam
    from reviews4rec_b200.train import train
    from reviews4rec_b200.utils import init_transnet_optim
    from tests.test_gpu_models import ListReader, build
    z, dims = load_golden(mt)
    model, hp = build(mt, z, dims)
    S.shard_model(model, S.Transport())
    if mt.startswith("transnet"):
        opt = init_transnet_optim(hp, model, FusedAdam)
    else:
        opt = FusedAdam(model.parameters(), lr=hp["lr"], weight_decay=hp["weight_decay"])
    metrics = train(model, R.MSELoss(hp), opt, ListReader(golden_batches(z, dims, "cuda")), hp)
    raw = train.last_raw
    if mt.startswith("transnet"):
        assert_close(raw["se_sum"], float(z["metric.MSE_sum"]), rtol=1e-4, msg="MSE sum")
    else:
        assert abs(metrics["MSE"] - float(z["metric.MSE"])) <= 1e-4 * max(1.0, abs(float(z["metric.MSE"])))
    ref = golden_state(z, "final")
    sd = S.gather_state_dict(model)
    assert set(sd) == set(ref)
    for k in ref:
        atol = hp["lr"] * dims["NB"] if (mt == "NARRE" and k.startswith("attention_scorer_") and k.endswith(".3.bias")) else 4e-6
        assert_close(sd[k], ref[k], rtol=1e-4, atol=atol, msg="%s final.%s" % (mt, k))


@pytest.mark.parametrize("mode", ["f16", "bf16"])
@pytest.mark.parametrize("mt", ["deepconn", "NARRE", "transnet"])
def test_sharded_model_world1_tensor_core_modes(S, mt, mode):
    """Half-precision rows served by the owner's shadow shard: same train-loop MSE as the reference (1e-4)
    and bit-identical ratings to the unsharded model in the same mode."""
    import reviews4rec_b200 as R
    from reviews4rec_b200.optim import FusedAdam
    from reviews4rec_b200.train import train
    from reviews4rec_b200.utils import init_transnet_optim
    from tests.helpers import golden_data
    from tests.test_gpu_models import ListReader, build
    z, dims = load_golden(mt)
    plain, hp = build(mt, z, dims, mode=mode)
    model, _ = build(mt, z, dims, mode=mode)
    S.shard_model(model, S.Transport())
    plain.eval(), model.eval()
    with torch.no_grad():
        a, b = plain(golden_data(z, "b0", "cuda")), model(golden_data(z, "b0", "cuda"))
    a, b = (a if isinstance(a, list) else [a]), (b if isinstance(b, list) else [b])
    assert all(torch.equal(x, y) for x, y in zip(a, b)), "sharded lookup changed the ratings"
    opt = init_transnet_optim(hp, model, FusedAdam) if mt.startswith("transnet") else FusedAdam(model.parameters(), lr=hp["lr"], weight_decay=hp["weight_decay"])
    metrics = train(model, R.MSELoss(hp), opt, ListReader(golden_batches(z, dims, "cuda")), hp)
    raw = train.last_raw
    if mt.startswith("transnet"):
        assert_close(raw["se_sum"], float(z["metric.MSE_sum"]), rtol=1e-4, msg="MSE sum")
    else:
        assert abs(metrics["MSE"] - float(z["metric.MSE"])) <= 1e-4 * max(1.0, abs(float(z["metric.MSE"])))


@pytest.mark.parametrize("mt,mode", [("deepconn", "exact"), ("deepconn", "f16"), ("NARRE", "f16")])
def test_captured_steps_with_prefetched_word_lookup_world1(S, mt, mode):
    """train.CapturedStep(next_data=...): the word lookup of step k+1 rides on a forked branch of step k's graph and
    step k+1 reads the rows from persistent slots.  Same training as captured steps that look up in place."""
    import reviews4rec_b200 as R
    from reviews4rec_b200.optim import FusedAdam
    from reviews4rec_b200.train import CapturedStep
    from tests.test_gpu_models import build
    z, dims = load_golden(mt)
    batches = golden_batches(z, dims, "cuda")
    finals = []
    for prefetch in (False, True):
        model, hp = build(mt, z, dims, mode=mode)
        S.shard_model(model, S.Transport())
        model.train()
        opt = FusedAdam(model.parameters(), lr=hp["lr"], weight_decay=hp["weight_decay"], capturable=True)
        se = torch.zeros(1, device="cuda")
        steps = [CapturedStep(model, R.MSELoss(hp), opt, d, y, se, next_data=batches[(i + 1) % len(batches)][0] if prefetch else None)
                 for i, (d, y) in enumerate(batches)]
        steps[0].prime()
        for _ in range(3):
            for st in steps:
                st.replay()
        torch.cuda.synchronize()
        finals.append((S.gather_state_dict(model), float(se)))
    (a, sa), (b, sb) = finals
    assert abs(sa - sb) <= 1e-5 * abs(sa)
    for k in a:
        atol = 0.002 * 9 if (mt == "NARRE" and k.startswith("attention_scorer_") and k.endswith(".3.bias")) else 4e-6
        assert_close(b[k], a[k], rtol=1e-4, atol=atol, msg="%s %s" % (mt, k))


@pytest.mark.parametrize("transport", ["nccl", "p2p"])
def test_two_rank_training_matches_single_process_reference(transport):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run under gpurun --gpus 2)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(29600 + os.getpid() % 300), os.path.join(ROOT, "tests", "dist_parity.py"), "--transport", transport]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=200, cwd=ROOT)
    assert res.returncode == 0 and "DIST_PARITY_OK" in res.stdout, "\n".join(l for l in (res.stdout + res.stderr).splitlines() if "Error" in l or "error" in l or "assert" in l or "dist_parity" in l)[-3000:]


@pytest.mark.parametrize("sharded_tables", [False, True])
@pytest.mark.parametrize("mt", ["deepconn", "deepconn++", "NARRE", "transnet++", "MF_dot"])
def test_empty_local_batch(S, mt, sharded_tables):
    """A rank's slice of a small global batch can be empty (5 ratings over 8 ranks): forward and backward must run
    (the collectives of the sharded tables still have to be entered) and leave zero gradients."""
    import reviews4rec_b200 as R
    from tests.test_gpu_models import build
    z, dims = load_golden(mt)
    model, hp = build(mt, z, dims, mode="f16")
    if sharded_tables:
        S.shard_model(model, S.Transport())
    model.train()
    data, y = golden_batches(z, dims, "cuda")[0]
    data = [None if d is None else d[:0].contiguous() for d in data]
    out = model(data)
    outs = out if isinstance(out, list) else [out]
    assert tuple(outs[0].shape) == (0,)
    loss = R.MSELoss(hp)(outs[0], y[:0], return_mean=False).sum()
    if mt.startswith("transnet"):
        loss = loss + R.MSELoss(hp)(outs[1], y[:0], return_mean=False).sum()
    loss.backward()
    for n, p in model.named_parameters():
        if p.grad is not None:
            assert float(p.grad.abs().max()) == 0.0, n
