"""Micro-benchmark of the fused gather+conv+pool kernel alone (scripts/, not part of the product).
usage: python scripts/conv_bench.py [--docs N] [--dist amazon|uniform|nopad] [--mode f16]"""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from reviews4rec_b200 import ops
from reviews4rec_b200.synthetic import _Zipf, _docs

ap = argparse.ArgumentParser()
ap.add_argument("--docs", type=int, default=4096)
ap.add_argument("--T", type=int, default=1000)
ap.add_argument("--E", type=int, default=300)
ap.add_argument("--V", type=int, default=50001)
ap.add_argument("--dist", default="amazon")
ap.add_argument("--mode", default="f16")
ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--ragged", action="store_true")
a = ap.parse_args()
rng = np.random.default_rng(0)
pool = []
for i in range(3):
    if a.dist == "amazon":
        idx = _docs(rng, _Zipf(a.V - 1, 1.0), a.docs, a.T)
    elif a.dist == "nopad":
        idx = _Zipf(a.V - 1, 1.0).draw(rng, (a.docs, a.T))
    else:
        idx = rng.integers(0, a.V, (a.docs, a.T))
    if a.ragged:
        from reviews4rec_b200.readers import RaggedDocs
        rd = RaggedDocs(idx, pin=False)
        pool.append(ops.RaggedIdx(rd.tokens.cuda(), rd.offsets.cuda(), idx.shape, 0))
    else:
        pool.append(torch.from_numpy(idx).cuda())
g = torch.Generator(device="cuda").manual_seed(0)
table = (torch.rand(a.V, a.E, device="cuda", generator=g) - 0.5) * 0.07
w = (torch.rand(100, 1, 3, a.E, device="cuda", generator=g) - 0.5) * 0.15
b = torch.zeros(100, device="cuda")
sh = ops.ShadowTable()
for i in range(2):
    ops.conv_pool_forward(pool[0], table, w, b, a.mode, sh)
torch.cuda.synchronize()
ts, tot = [], []
for i in range(a.iters):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sink = []
    ops.set_conv_event_sink(sink)
    e0.record()
    ops.conv_pool_forward(pool[i % 3], table, w, b, a.mode, sh)
    e1.record()
    ops.set_conv_event_sink(None)
    torch.cuda.synchronize()
    ts.append(sink[0][0].elapsed_time(sink[0][1]))
    tot.append(e0.elapsed_time(e1))
ms = sum(ts) / len(ts)
print("whole call (pack + plan + conv): %.3f ms" % (sum(tot) / len(tot)))
fl = 2.0 * (a.T + 2) * 100 * 3 * a.E * a.docs
print("conv_bench ragged=%s dist=%s docs=%d mode=%s env=%s: %.3f ms/launch  %.1f TFLOP/s  %.0f docs/s  alg %.0f GB/s" % (
    a.ragged, a.dist, a.docs, a.mode, {k: v for k, v in os.environ.items() if k.startswith("R4R_")}, ms, fl / ms / 1e9,
    a.docs / ms * 1e3, a.docs * a.T * (8 + 4 * a.E) / ms / 1e6))
if os.environ.get("CONV_PROF"):
    from reviews4rec_b200 import _lib
    import ctypes
    buf = torch.zeros(32, dtype=torch.int64, device="cuda")
    _lib.lib.r4r_conv_debug_profile(ctypes.c_void_p(buf.data_ptr()))
    ops.conv_pool_forward(pool[0], table, w, b, a.mode, sh)
    torch.cuda.synchronize()
    _lib.lib.r4r_conv_debug_profile(ctypes.c_void_p(0))
    v = buf.tolist()
    for r in (0, 1):
        print("rank%d epilogue: total %d  wait_tmem_full %d  bar %d  xchg %d" % (r, *v[r * 16: r * 16 + 4]))
        print("rank%d tma team: total %d  wait_empty %d  tma_issue %d" % (r, *v[r * 16 + 4: r * 16 + 7]))
        print("rank%d cp.async team: total %d  wait_empty %d  wait_copies+publish %d" % (r, *v[r * 16 + 12: r * 16 + 15]))
    print("mma: total %d  wait_tmem_empty %d  wait_full %d" % tuple(v[8:11]))
    print("rank0 epilogue warp 0: tile passes %d  reduction %d  merge+exchange (incl. its waits) %d  documents %d" % tuple(v[24:28]))
