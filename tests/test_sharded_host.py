"""CPU tests of the multi-GPU host logic: shard layout, batch split, and -- under a
world_size-2 gloo group -- the complete lookup / gradient protocol of reviews4rec_b200/sharded.py with
the device kernels replaced by tests/shard_emul.py."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.shard_emul import CpuKernels


def test_shard_unshard_roundtrip():
    from reviews4rec_b200.sharded import rows_local, shard_rows, unshard_rows
    for R, P in [(10, 1), (10, 2), (11, 3), (7, 8), (1000003, 8)]:
        full = torch.arange(R * 3, dtype=torch.float32).reshape(R, 3) if R < 1000 else torch.arange(R, dtype=torch.float32)
        shards = [shard_rows(full, r, P) for r in range(P)]
        assert all(s.shape[0] == rows_local(R, P) for s in shards)
        assert torch.equal(unshard_rows(shards, R), full)
        for r in range(P):                                   # row g lives on rank g % P at local row g // P
            g = torch.arange(r, R, P)
            assert torch.equal(shards[r][g // P], full[g])


def test_shard_batch_covers_the_batch():
    from reviews4rec_b200.sharded import shard_batch
    B = 11
    data = [None, torch.arange(B * 2).reshape(B, 2), None, torch.arange(B * 5).reshape(B, 5), torch.arange(B), torch.arange(B), torch.arange(B)]
    y = torch.arange(B, dtype=torch.float32)
    for P in (1, 2, 4, 8):
        parts = [shard_batch(data, y, r, P) for r in range(P)]
        assert torch.equal(torch.cat([p[1] for p in parts]), y)
        assert torch.equal(torch.cat([p[0][3] for p in parts]), data[3])
        assert all(p[0][0] is None for p in parts)


def test_lookup_protocol_world_size_1_no_process_group(monkeypatch):
    """Degenerate single-rank case (what `bench.py --force-shard` and the world-1 GPU tests run)."""
    from reviews4rec_b200 import ops, sharded
    monkeypatch.setattr(sharded, "K", CpuKernels())
    monkeypatch.setattr(ops, "_conv_mode", "exact")
    g = torch.Generator().manual_seed(3)
    V, E, R, L, n = 29, 5, 13, 4, 11
    full_words, full_rows = torch.randn(V, E, generator=g), torch.randn(R, L, generator=g)
    tr = sharded.Transport()
    assert (tr.world, tr.rank) == (1, 0)
    wt = sharded.ShardedWordTable(full_words, tr)
    idx = torch.randint(0, V, (3, 8), generator=g)
    (d,) = wt.many(idx)
    assert torch.equal(d.table[d.idx], full_words[idx])
    p = torch.nn.Parameter(sharded.shard_rows(full_rows, 0, 1))
    p._r4r_shard = sharded.ShardSpec(R, tr)
    ids = torch.randint(0, R, (n,), generator=g)
    w = torch.randn(n, L, generator=g)
    rows = sharded.sharded_rows_gather(p, ids)
    assert torch.equal(rows, full_rows[ids])
    (rows * w).sum().backward()
    ref = torch.zeros(R, L).index_add_(0, ids, w)
    torch.testing.assert_close(p.grad, ref)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from reviews4rec_b200 import ops, sharded
        ops.set_conv_mode("exact")                            # fp32 rows (the half-precision shadow is built by a CUDA kernel)
        sharded.K = CpuKernels()                              # host protocol under test; kernels emulated
        tr = sharded.Transport()
        assert (tr.world, tr.rank) == (world, rank)
        g = torch.Generator().manual_seed(5)                  # same on every rank
        V, E, R, L, n = 41, 6, 23, 3, 17
        full_words = torch.randn(V, E, generator=g)
        full_rows = torch.randn(R, L, generator=g)
        full_bias = torch.randn(R, generator=g)
        gr = torch.Generator().manual_seed(100 + rank)        # different batches per rank
        idx_a = torch.randint(0, V, (4, 9), generator=gr)
        idx_b = torch.randint(0, V, (3, 9), generator=gr)
        ids = torch.randint(0, R, (n,), generator=gr)
        ids[:5] = R - 1                                       # a hot (pad-like) row
        w = torch.randn(n, L, generator=gr)

        # ---- word table: one exchange, rows bit-exact, read through the ORIGINAL token ids (id-indexed row cache)
        wt = sharded.ShardedWordTable(full_words, tr)
        da, db = wt.many(idx_a, idx_b)
        assert da.table is db.table and da.table.shape == (V, E) and da.shadow is None
        assert da.idx is idx_a or torch.equal(da.idx, idx_a)
        assert torch.equal(da.table[da.idx], full_words[idx_a]) and torch.equal(db.table[db.idx], full_words[idx_b])
        assert int(wt._scratch(torch.device("cpu"))[0].sum()) == 0     # flags left clean for the next step

        # ---- id table + bias vector: forward rows, backward grads land (averaged) on the owners
        p_rows = torch.nn.Parameter(sharded.shard_rows(full_rows, rank, world))
        p_rows._r4r_shard = sharded.ShardSpec(R, tr)
        p_bias = torch.nn.Parameter(sharded.shard_rows(full_bias, rank, world))
        p_bias._r4r_shard = sharded.ShardSpec(R, tr)
        rows = sharded.sharded_rows_gather(p_rows, ids)
        bias = sharded.sharded_rows_gather(p_bias, ids.reshape(1, n))
        assert torch.equal(rows, full_rows[ids]) and torch.equal(bias, full_bias[ids].reshape(1, n))
        loss = (rows * w).sum() / n + (bias.reshape(-1) * w[:, 0]).sum() / n      # local-batch mean, like main.py:58
        loss.backward()

        # single-process reference over the global batch
        all_ids = [torch.empty_like(ids) for _ in range(world)]
        all_w = [torch.empty_like(w) for _ in range(world)]
        dist.all_gather(all_ids, ids)
        dist.all_gather(all_w, w)
        gi, gw = torch.cat(all_ids), torch.cat(all_w)
        ref_rows = full_rows.clone().requires_grad_(True)
        ref_bias = full_bias.clone().requires_grad_(True)
        ((ref_rows[gi] * gw).sum() / gi.numel() + (ref_bias[gi] * gw[:, 0]).sum() / gi.numel()).backward()
        torch.testing.assert_close(p_rows.grad, sharded.shard_rows(ref_rows.grad, rank, world), rtol=1e-6, atol=1e-7)
        torch.testing.assert_close(p_bias.grad, sharded.shard_rows(ref_bias.grad, rank, world), rtol=1e-6, atol=1e-7)

        # ---- replicated parameters: gradient mean; state_dict re-assembly in the reference layout
        class Tiny(torch.nn.Module):
            def __init__(self):
                super().__init__()
                self.user_bias = torch.nn.Parameter(full_bias.clone())
                self.lin = torch.nn.Linear(3, 2)
        m = Tiny()
        with torch.no_grad():
            for p_ in m.lin.parameters():
                p_.fill_(0.5)
        sharded.shard_model(m, tr, shard_word_table=False)
        assert m.user_bias.shape[0] == sharded.rows_local(R, world) and hasattr(m.user_bias, "_r4r_shard")
        for p_ in m.lin.parameters():
            p_.grad = torch.full_like(p_, float(rank + 1))
        m.user_bias.grad = torch.full_like(m.user_bias, 7.0)
        sharded.allreduce_dense_grads(m)
        assert all(torch.allclose(p_.grad, torch.full_like(p_, (world + 1) / 2.0)) for p_ in m.lin.parameters())
        assert torch.equal(m.user_bias.grad, torch.full_like(m.user_bias, 7.0))   # sharded grads are not all-reduced
        sd = sharded.gather_state_dict(m)
        assert torch.equal(sd["user_bias"], full_bias) and set(sd) == {"user_bias", "lin.weight", "lin.bias"}
        out.put((rank, "ok"))
    except Exception as exc:                                  # surface the failure in the parent
        import traceback
        out.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_lookup_protocol_world_size_2_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=150) for _ in range(world)]
    for p in procs:
        p.join(timeout=30)
    for rank, msg in res:
        assert msg == "ok", "rank %d: %s" % (rank, msg)
