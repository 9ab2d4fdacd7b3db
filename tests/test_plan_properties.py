"""CPU checks (oracle only) of the two exactness claims the conv launch's work plan rests on (docplan.cu, conv_tc.cu):

1. padding-run shortcut: a document whose rows s..T-1 repeat one token gives the same (pooled, first arg-max) when cut to
   T' = min(T, s+3) rows, with arg-max positions >= T' mapped back by + (T - T');
2. window streams: documents laid end to end with two zero rows between them (shared by the window that closes one and
   the window that opens the next) -- document i of L rows owns exactly the L+2 consecutive windows starting at its first
   zero row, and the max / first arg-max over those windows of ONE conv over the stream equals the per-document conv.

The GPU tests check the kernels; these pin the arithmetic identities themselves against the reference's conv
(common_pytorch_models.py:26-31 via oracle.conv_pool)."""
import pytest
import torch

from oracle import r4r_oracle as O


def _case(seed, N, T, E, V, F=16):
    g = torch.Generator().manual_seed(seed)
    table = torch.randn(V, E, generator=g, dtype=torch.float64)
    w = torch.randn(F, 1, 3, E, generator=g, dtype=torch.float64) / (3 * E) ** 0.5
    b = torch.randn(F, generator=g, dtype=torch.float64) * 0.1
    idx = torch.randint(1, V, (N, T), generator=g)
    runs = [0, 1, 2, 3, 4, T, T - 1, T // 2] + [int(torch.randint(0, T + 1, (1,), generator=g)) for _ in range(N)]
    for n in range(N):
        r = runs[n]
        if r > 0:
            idx[n, T - r:] = 0 if n % 2 else int(torch.randint(1, V, (1,), generator=g))
    return table, idx, w, b


def _full_conv(table, idx, w, b):
    """[N, F, T+2] post-ReLU conv output of every position, as the reference computes it (padding (2, 0))."""
    x = O.word_gather(table, idx)
    return torch.nn.functional.conv2d(x.unsqueeze(1), w, b, padding=(2, 0)).squeeze(-1).relu()


def _assert_first_max(y_row, p, a, ref_a):
    """(p, a) is the maximum of y_row [F, T+2] and a FIRST maximiser.  Windows inside a padding run are equal in exact
    arithmetic, and the reference's own CPU conv breaks such ties by rounding noise (it may report any position of the
    run), so: the value at `a` is the maximum, and `a` is not later than the position the reference reports."""
    live = p > 0                                                   # all-zero (post-ReLU) columns tie everywhere
    at = y_row.gather(1, a.unsqueeze(1)).squeeze(1)
    assert torch.allclose(at[live], y_row.max(dim=1).values[live], rtol=0, atol=1e-12)
    assert bool((a[live] <= ref_a[live]).all())


def _doc_len(row):
    T = len(row)
    s = T - 1
    while s > 0 and row[s - 1] == row[T - 1]:
        s -= 1
    return min(T, s + 3)


@pytest.mark.parametrize("N,T,E,V", [(12, 40, 8, 30), (10, 7, 5, 9), (9, 300, 6, 50)])
def test_padding_run_shortcut_is_exact(N, T, E, V):
    table, idx, w, b = _case(1, N, T, E, V)
    full_p, full_a = O.conv_pool(O.word_gather(table, idx), w, b)
    y = _full_conv(table, idx, w, b)
    for n in range(N):
        Td = _doc_len(idx[n].tolist())
        p, a = O.conv_pool(O.word_gather(table, idx[n:n + 1, :Td]), w, b)
        a = torch.where((a >= Td) & (a < Td + 2), a + (T - Td), a)
        assert torch.allclose(p[0], full_p[n], rtol=0, atol=1e-12)
        _assert_first_max(y[n], p[0], a[0], full_a[n])


@pytest.mark.parametrize("N,T,E,V", [(12, 40, 8, 30), (25, 7, 5, 9), (6, 300, 6, 50)])
def test_window_stream_equals_per_document_conv(N, T, E, V):
    table, idx, w, b = _case(2, N, T, E, V)
    lens = [_doc_len(idx[n].tolist()) for n in range(N)]
    ref_p, ref_a = O.conv_pool(O.word_gather(table, idx), w, b)
    y_full = _full_conv(table, idx, w, b)
    # the stream: per document two zero rows, then its (effective) rows; two more zero rows close the last document
    rows, first = [], []
    for n in range(N):
        first.append(len(rows))
        rows += [None, None] + idx[n, :lens[n]].tolist()
    rows += [None, None]
    x = torch.stack([torch.zeros(E, dtype=torch.float64) if r is None else table[r] for r in rows])      # [R, E]
    # window u covers stream rows u, u+1, u+2: one valid (un-padded) conv over the whole stream
    y = torch.nn.functional.conv2d(x[None, None], w, b).squeeze(-1)[0].relu()                              # [F, R-2]
    assert sum(l + 2 for l in lens) == y.shape[1]                                                          # every window has exactly one owner
    for n in range(N):
        seg = y[:, first[n]: first[n] + lens[n] + 2]
        p, a = seg.max(dim=1)
        a = torch.where((a >= lens[n]) & (a < lens[n] + 2), a + (T - lens[n]), a)
        assert torch.allclose(p, ref_p[n], rtol=0, atol=1e-12)
        _assert_first_max(y_full[n], p, a, ref_a[n])
