"""Determinism probe of the fused gather+conv+pool kernel (scripts/, not part of the product).
Runs the SAME launch `--reps` times plus split launches and reports every (document, filter) whose pooled value or
arg-max differs from the per-entry majority.  usage: python scripts/conv_repro.py [--reps 30] [--mode f16]"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from reviews4rec_b200 import ops
from reviews4rec_b200.synthetic import _Zipf, _docs

ap = argparse.ArgumentParser()
ap.add_argument("--reps", type=int, default=30)
ap.add_argument("--mode", default="f16")
ap.add_argument("--docs", type=int, default=4096)
ap.add_argument("--noplan", action="store_true")
a = ap.parse_args()
g = torch.Generator(device="cuda").manual_seed(0)
V, E, F, T = 50001, 300, 100, 1000
table = (torch.rand(V, E, device="cuda", generator=g) - 0.5) * 0.07
w = (torch.rand(F, 1, 3, E, device="cuda", generator=g) - 0.5) * 0.15
b = (torch.rand(F, device="cuda", generator=g) - 0.5) * 0.1
idx = torch.from_numpy(_docs(np.random.default_rng(1), _Zipf(V - 1, 1.0), a.docs, T)).cuda()
if a.noplan:
    ops.set_doc_plan(False)
sh = ops.ShadowTable()
dl = ops.doc_lengths(idx)
dbg = None
if os.environ.get("CONV_DBG"):
    import ctypes
    from reviews4rec_b200 import _lib
    dbg = torch.zeros(32, dtype=torch.int64, device="cuda")
    _lib.lib.r4r_conv_debug_profile(ctypes.c_void_p(dbg.data_ptr()))
runs = []
for r in range(a.reps):
    if r % 3 == 2:
        cut = 1500
        p1, a1 = ops.conv_pool_forward(idx[:cut], table, w, b, a.mode, sh)
        p2, a2 = ops.conv_pool_forward(idx[cut:], table, w, b, a.mode, sh)
        runs.append((torch.cat([p1, p2]), torch.cat([a1, a2])))
    else:
        runs.append(ops.conv_pool_forward(idx, table, w, b, a.mode, sh))
torch.cuda.synchronize()
P = torch.stack([r[0] for r in runs]).view(torch.int32)    # bit patterns
A = torch.stack([r[1] for r in runs])
Pm, Am = P.mode(0).values, A.mode(0).values
bad_runs = 0
for r in range(a.reps):
    dp, da = P[r] != Pm, A[r] != Am
    if bool(dp.any()) or bool(da.any()):
        bad_runs += 1
        d = (dp | da).nonzero()
        print("run %d (%s): %d pooled, %d argmax entries differ in %d docs" % (
            r, "split" if r % 3 == 2 else "full", int(dp.sum()), int(da.sum()), d[:, 0].unique().numel()))
        for n, f in d[:12].tolist():
            print("   doc %d (len %d) filter %d: pooled %.9g vs %.9g   argmax %d vs %d" % (
                n, int(dl[n]), f, float(runs[r][0][n, f]), float(Pm.view(torch.float32)[n, f]), int(A[r, n, f]), int(Am[n, f])))
if dbg is not None:
    v = dbg.tolist()
    print("debug counters: xchg doc mismatch %d (last got/expected %d/%d)  remote pair inconsistent %d  local pair inconsistent %d  "
          "remote pos>=npos %d  local pos>=npos %d" % (v[24], v[25] >> 32, v[25] & 0xffffffff, v[26], v[27], v[28], v[29]))
print("conv_repro mode=%s env=%s lib=%s: %d of %d runs differ from the majority" % (
    a.mode, {k: v for k, v in os.environ.items() if k.startswith("R4R_")}, os.environ.get("R4R_LIB", "default"), bad_runs, a.reps))
