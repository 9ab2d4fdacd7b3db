"""Ragged reader (reviews4rec_b200/readers.py): host-side packing is exact on CPU; on the GPU the staged
batches must be bit-identical to the padded int64 arrays the reference's data_fast reader would ship."""
import numpy as np
import pytest
import torch


def _padded(rng, shape, V, T):
    arr = rng.integers(1, V, size=shape + (T,), dtype=np.int64)
    arr[rng.random(shape + (T,)) < 0.05] = 0                 # interior pad-id tokens are ordinary tokens
    lens = rng.integers(0, T + 1, size=shape)
    arr[np.arange(T)[(None,) * len(shape)] >= lens[..., None]] = 0
    flat = arr.reshape(-1, T)
    flat[0] = 0                                               # an all-padding document
    flat[1, :] = rng.integers(1, V, size=T)                   # a full one
    return arr


def test_ragged_docs_roundtrip_cpu():
    from reviews4rec_b200.readers import RaggedDocs
    rng = np.random.default_rng(0)
    for shape, T in [((37,), 50), ((9, 4), 7), ((5,), 1)]:
        arr = _padded(rng, shape, 100, T)
        rd = RaggedDocs(arr, pin=False)
        assert rd.tokens.dtype == torch.int32 and rd.offsets[-1] == rd.tokens.numel()
        assert rd.tokens.numel() <= arr.size and (arr.size == 0 or rd.tokens.numel() < arr.size)
        for lo, hi in [(0, shape[0]), (3, 5), (2, 2), (shape[0] - 1, shape[0])]:
            tok, off = rd.batch_host(lo, hi)
            assert np.array_equal(rd.to_padded(tok, off), arr[lo:hi])
        assert rd.max_batch_tokens(3) >= max(rd.batch_host(b, min(shape[0], b + 3))[0].numel() for b in range(0, shape[0], 3))


def test_ragged_docs_rejects_wide_ids():
    from reviews4rec_b200.readers import RaggedDocs
    with pytest.raises(ValueError):
        RaggedDocs(np.array([[1, 2 ** 31]], dtype=np.int64), pin=False)


@pytest.mark.gpu
@pytest.mark.parametrize("mt,N,B", [("deepconn", 53, 8), ("NARRE", 21, 4), ("transnet", 16, 16), ("MF_dot", 10, 3)])
def test_ragged_reader_matches_padded_arrays(mt, N, B):
    from reviews4rec_b200.readers import RaggedReader
    rng = np.random.default_rng(1)
    T, R, W, V = 60, 3, 9, 200
    hp = {"model_type": mt, "batch_size": B}
    arrays = {k: None for k in "abcdefgh"}
    arrays["f"], arrays["g"] = rng.integers(0, 50, N), rng.integers(0, 40, N)
    arrays["h"] = rng.integers(1, 6, N).astype(np.float64)      # HDF5 'h' is f8 (make_quick_data.py:21-44)
    if mt == "NARRE":
        arrays["d"], arrays["e"] = _padded(rng, (N, R), V, W), _padded(rng, (N, R), V, W)
        arrays["b"], arrays["c"] = rng.integers(0, 52, (N, 10)), rng.integers(0, 42, (N, 10))
    elif mt != "MF_dot":
        arrays["d"], arrays["e"] = _padded(rng, (N,), V, T), _padded(rng, (N,), V, T)
        if mt == "transnet":
            arrays["a"] = _padded(rng, (N,), V, T)
    reader = RaggedReader(hp, arrays, "cuda")
    assert len(reader) == (N + B - 1) // B
    seen = 0
    for epoch in range(2):                                      # slots are reused across epochs
        seen = 0
        for data, y in reader.iter():
            n = y.shape[0]
            for j, k in enumerate("abcdefg"):
                if arrays[k] is None:
                    assert data[j] is None
                else:
                    assert data[j].dtype == torch.int64 and data[j].is_cuda
                    assert np.array_equal(data[j].cpu().numpy(), arrays[k][seen:seen + n]), (k, seen)
            assert y.dtype == torch.float32 and np.array_equal(y.cpu().numpy(), arrays["h"][seen:seen + n].astype(np.float32))
            seen += n
        assert seen == N


def _ragged_on_device(arr):
    from reviews4rec_b200.ops import RaggedIdx
    from reviews4rec_b200.readers import RaggedDocs
    rd = RaggedDocs(arr, pin=False)
    return RaggedIdx(rd.tokens.cuda(), rd.offsets.cuda(), arr.shape, rd.pad_id)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["f16", "bf16", "exact"])
@pytest.mark.parametrize("N,T,E,V", [(70, 1000, 300, 400), (33, 300, 64, 90), (24, 9, 12, 30)])
def test_conv_and_wgrad_on_ragged_docs_match_padded(mode, N, T, E, V):
    """The kernels reading ragged documents must give the bits they give on the expanded padded ids."""
    from reviews4rec_b200 import ops
    rng = np.random.default_rng(4)
    arr = _padded(rng, (N,), V, T)
    g = torch.Generator().manual_seed(2)
    table = (torch.randn(V, E, generator=g) * 0.5).cuda()
    w = (torch.randn(100, 1, 3, E, generator=g) * (1.0 / (3 * E) ** 0.5)).cuda()
    b = (torch.randn(100, generator=g) * 0.1).cuda()
    gout = torch.randn(N, 100, generator=g).cuda()
    rg = _ragged_on_device(arr)
    assert torch.equal(rg.padded().cpu(), torch.from_numpy(arr))
    assert int(ops.doc_lengths(rg).max()) <= T
    outs = []
    try:
        for idx, native in ((torch.from_numpy(arr).cuda(), False), (rg, True), (rg, False)):
            ops.set_ragged_native(native)
            wc, bc = w.clone().requires_grad_(True), b.clone().requires_grad_(True)
            pooled, arg = ops._ConvPool.apply(idx, table, wc, bc, mode, None)
            pooled.backward(gout)
            outs.append((pooled.detach(), arg, wc.grad, bc.grad))
    finally:
        ops.set_ragged_native(False)
    torch.cuda.synchronize()
    assert torch.equal(outs[0][0], outs[2][0]) and torch.equal(outs[0][1], outs[2][1])
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    torch.testing.assert_close(outs[0][2], outs[1][2], rtol=1e-5, atol=1e-6)        # atomics: summation order only
    torch.testing.assert_close(outs[0][3], outs[1][3], rtol=1e-5, atol=1e-6)


@pytest.mark.gpu
@pytest.mark.parametrize("mt", ["deepconn", "NARRE", "transnet++"])
def test_models_read_native_ragged_batches(mt):
    """RaggedReader(native=True) hands ops.RaggedIdx documents to the models: same ratings (bit-exact) and
    the same short training run as with the expanded padded batches."""
    import pickle, tempfile, os
    import reviews4rec_b200 as R
    from reviews4rec_b200 import ops
    from reviews4rec_b200.optim import FusedAdam
    from reviews4rec_b200.readers import RaggedReader
    from reviews4rec_b200.train import train
    from reviews4rec_b200.utils import init_transnet_optim, xavier_init
    rng = np.random.default_rng(7)
    N, B, T, Rv, W, V, E, L, U, I = 24, 8, 300, 4, 40, 120, 32, 6, 30, 20
    tmp = tempfile.mkdtemp()
    with open(os.path.join(tmp, "word2vec.pkl"), "wb") as f:
        pickle.dump(np.zeros((V, E), dtype=np.float32), f, 4)
    hp = {"model_type": mt, "latent_size": L, "word_embed_size": E, "dropout": 0.0, "total_users": U, "total_items": I,
          "lr": 0.002, "weight_decay": 1e-6, "batch_size": B, "data_dir": tmp}
    arrays = {k: None for k in "abcdefgh"}
    arrays["f"], arrays["g"] = rng.integers(0, U, N), rng.integers(0, I, N)
    arrays["h"] = rng.integers(1, 6, N).astype(np.float64)
    if mt == "NARRE":
        arrays["d"], arrays["e"] = _padded(rng, (N, Rv), V, W), _padded(rng, (N, Rv), V, W)
        arrays["b"], arrays["c"] = rng.integers(0, U + 2, (N, Rv)), rng.integers(0, I + 2, (N, Rv))   # one neighbour id per review
    else:
        arrays["d"], arrays["e"] = _padded(rng, (N,), V, T), _padded(rng, (N,), V, T)
        arrays["a"] = _padded(rng, (N,), V, T)
    ops.set_conv_mode("f16")
    results = []
    for native in (False, True):
        ops.set_ragged_native(native)
        torch.manual_seed(0)
        cls = {"deepconn": R.DeepCoNN, "NARRE": R.NARRE, "transnet++": R.TransNet}[mt]
        model = cls(hp)
        xavier_init(model)
        model = model.cuda()
        reader = RaggedReader(hp, arrays, "cuda", native=native)
        model.eval()
        with torch.no_grad():
            first = next(iter(reader.iter()))
            out = model(first[0])
            out = [o.clone() for o in out] if isinstance(out, list) else [out.clone()]
        torch.cuda.synchronize()
        opt = init_transnet_optim(hp, model, FusedAdam) if mt.startswith("transnet") else FusedAdam(model.parameters(), lr=hp["lr"], weight_decay=hp["weight_decay"])
        train(model, R.MSELoss(hp), opt, reader, hp)
        results.append((out, train.last_raw["se_sum"]))
    ops.set_ragged_native(False)
    for a, b in zip(results[0][0], results[1][0]):
        assert torch.equal(a, b)
    assert abs(results[0][1] - results[1][1]) <= 1e-5 * abs(results[0][1])


def test_ragged_idx_shape_logic_cpu():
    """RaggedIdx is a shape-carrying handle: the reshapes the models perform (DeepCoNN.py:40-50, NARRE.py:91-100)
    must keep the document length and share the buffers."""
    from reviews4rec_b200.ops import RaggedIdx
    tok = torch.arange(10, dtype=torch.int32)
    off = torch.tensor([0, 2, 2, 5, 6, 8, 10], dtype=torch.int64)
    r = RaggedIdx(tok, off, (2, 3, 7))                        # NARRE: [B, R, W]
    assert r.dim() == 3 and r.numel() == 42 and tuple(r.shape) == (2, 3, 7)
    f = r.reshape(6, 7)
    assert tuple(f.shape) == (6, 7) and f.tokens is tok and f.offsets is off
    assert tuple(r.reshape(6, -1).shape) == (6, 7) and tuple(f.reshape(2, 3, 7).shape) == (2, 3, 7)
    with pytest.raises(RuntimeError):
        r.reshape(3, 14)                                      # would merge documents
    with pytest.raises(ValueError):
        RaggedIdx(tok, off, (5, 7))                           # offsets do not describe 5 rows
    with pytest.raises(TypeError):
        RaggedIdx(tok.long(), off, (6, 7))
    with pytest.raises(RuntimeError):
        f.padded()                                            # no CPU fallback
