"""Stand-alone HBM rooflines of the gather / scatter kernels at V = 2,000,000 rows (SURVEY.md 8d "Standalone kernel
figures": a table that cannot sit in the 126 MB L2).  Each kernel: CUDA-event time over `--iters` launches on rotating
inputs larger than L2, algorithmic bytes / time vs MEASURED_PEAKS.json.  Under `ncu --metrics dram__bytes_*` the same
launches give the DRAM traffic next to the algorithmic bytes (scripts/gpu_rooflines.sh).
usage: python scripts/roofline_standalone.py [--iters 10] [--json out.json]"""
import argparse, ctypes, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from reviews4rec_b200 import _lib, ops
from reviews4rec_b200._lib import call
from reviews4rec_b200.ops import _p, _stream

ap = argparse.ArgumentParser()
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--json", default=None)
a = ap.parse_args()
peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
HBM = float(peaks["hbm_gbs"])
g = torch.Generator(device="cuda").manual_seed(0)
out = {}


def timed(fn, iters, warm=3):
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(warm + i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def report(name, ms, alg_bytes, note):
    gbs = alg_bytes / (ms * 1e-3) / 1e9
    out[name] = {"ms_per_launch": ms, "algorithmic_bytes_per_launch": alg_bytes, "achieved_gbs": gbs, "frac_of_hbm_peak": gbs / HBM,
                 "hbm_peak_gbs": HBM, "note": note}
    print("%-34s %8.3f ms  %9.1f MB algorithmic  %7.1f GB/s = %.3f of the measured %.0f GB/s   (%s)" % (name, ms, alg_bytes / 1e6, gbs, gbs / HBM, HBM, note))


# ---- K1 word_gather: N tokens out of a [2M, 300] fp32 table (2.4 GB); bytes = N * (8 + 4E + 4E)
V, E = 2_000_000, 300
table = torch.rand(V, E, device="cuda", generator=g)
N = 1 << 20
for dist in ("uniform", "zipf"):
    if dist == "uniform":
        idx = [torch.randint(0, V, (N,), device="cuda", generator=g) for _ in range(3)]
    else:
        w = 1.0 / torch.arange(1, V + 1, device="cuda", dtype=torch.float64)
        cdf = torch.cumsum(w, 0); cdf /= cdf[-1].clone()
        idx = [torch.searchsorted(cdf, torch.rand(N, device="cuda", generator=g, dtype=torch.float64)).clamp_(max=V - 1) for _ in range(3)]
    dst = torch.empty(N, E, device="cuda")
    ms = timed(lambda i: call("r4r_word_gather_f32", _p(table), V, E, _p(idx[i % 3]), N, _p(dst), _stream()), a.iters)
    report("word_gather_f32 (%s ids)" % dist, ms, N * (8 + 8 * E), "N=%d tokens, V=%d, E=%d: id + row read + row write" % (N, V, E))
del dst

# ---- K6 rows_scatter_add: id-table gradient scatter into a dense [R, L] gradient; bytes = n*(8 + 4L) + n_unique*2*4L
for R, L, n in ((2_000_000, 32, 1 << 22), (2_000_000, 10, 1 << 22)):
    ids = [torch.randint(0, R, (n,), device="cuda", generator=g) for _ in range(3)]
    gout = torch.rand(n, L, device="cuda", generator=g)
    gt = torch.zeros(R, L, device="cuda")
    nu = int(torch.unique(ids[0]).numel())
    ms = timed(lambda i: call("r4r_rows_scatter_add", _p(gout), _p(ids[i % 3]), n, L, _p(gt), R, _stream()), a.iters)
    report("rows_scatter_add L=%d" % L, ms, n * (8 + 4 * L) + nu * 8 * L, "n=%d uniform ids into [%d, %d] fp32 (%d distinct rows: read-modify-write)" % (n, R, L, nu))
    got = torch.zeros(R, L, device="cuda")
    call("r4r_rows_scatter_add", _p(gout), _p(ids[0]), n, L, _p(got), R, _stream())
    ref = torch.zeros(R, L, device="cuda").index_add_(0, ids[0], gout)
    assert float((got - ref).abs().max()) < 1e-3, "rows_scatter_add mismatch"
    del ids, gout, gt, got, ref

# ---- K5 rows_gather
for R, L, n in ((2_000_000, 32, 1 << 22),):
    ids = [torch.randint(0, R, (n,), device="cuda", generator=g) for _ in range(3)]
    tb = torch.rand(R, L, device="cuda", generator=g)
    o = torch.empty(n, L, device="cuda")
    ms = timed(lambda i: call("r4r_rows_gather", _p(tb), R, L, _p(ids[i % 3]), n, _p(o), _stream()), a.iters)
    report("rows_gather L=%d" % L, ms, n * (8 + 8 * L), "n=%d uniform ids out of [%d, %d] fp32" % (n, R, L))
    del ids, tb, o

# ---- opt-in word-table gradient scatter (f3): <= 3F row updates of 4E bytes per document into [2M, 300]
Nd, T, F = 4096, 1000, 100
idx2 = torch.randint(0, V, (Nd, T), device="cuda", generator=g)
argmax = torch.randint(0, T + 2, (Nd, F), device="cuda", generator=g, dtype=torch.int32)
pooled = torch.rand(Nd, F, device="cuda", generator=g)
gp = torch.randn(Nd, F, device="cuda", generator=g)
w = torch.randn(F, 1, 3, E, device="cuda", generator=g)
gt = torch.zeros(V, E, device="cuda")
ms = timed(lambda i: call("r4r_conv_dgrad_scatter", _p(idx2), Nd, T, _p(argmax), _p(pooled), _p(gp), _p(w), F, E, _p(gt), V, _stream()), a.iters)
rows = Nd * 3 * F
report("conv_dgrad_scatter", ms, Nd * F * 16 + rows * (8 + 2 * 4 * E), "%d documents x 3F row updates into [%d, %d]: read-modify-write of 4E bytes per update" % (Nd, V, E))
if a.json:
    json.dump(out, open(a.json, "w"), indent=1)
