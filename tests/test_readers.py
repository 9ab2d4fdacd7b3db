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
