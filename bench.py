#!/usr/bin/env python
"""bench.py -- train ratings/sec of the DeepCoNN rating-prediction hot path on B200.

    python bench.py --gpus N --steps K --warmup W            (this repo's sm_100a path)
    python bench.py --impl reference --gpus N --steps K --warmup W   (reference algorithm, host cores)

Workload = BASELINE.json configs[1]: DeepCoNN (`deepconn`, FM head), word_emb E=300, 100 conv
filters, doc_len T=1000, latent 10, V=50,001 words, 1M users / 100k items, synthetic Amazon-shaped
batches (reviews4rec_b200/synthetic.py).  A step = one training batch through main.train()'s body:
forward, per-sample squared error, backward, Adam (lr 0.002, weight_decay 1e-6, dropout 0.6).

One JSON line on stdout (rank 0):
  value     ratings/s with the batches already resident in HBM (a pool larger than L2, cycled)
  e2e       ratings/s through the public API from pinned HOST batches: H2D of the batch and D2H of
            the batch's squared-error sum are inside the timed region, every step
  roofline  the dominant kernel (fused gather+conv+pool): algorithmic bytes per launch / its
            CUDA-event duration inside the timed region, against MEASURED_PEAKS.json
  cpu_baseline  the oracle's CPU restatement of the same step on a bounded sample (N=1 only)
"""
import argparse
import json
import os
import pickle
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

V_WORDS = 50001
HP = {"model_type": "deepconn", "latent_size": 10, "word_embed_size": 300, "input_length": 1000, "dropout": 0.6,
      "total_users": 1000000, "total_items": 100000, "lr": 0.002, "weight_decay": 1e-6, "batch_size": 4096,
      "narre_num_reviews": 10, "narre_num_words": 200}
METRIC = "train ratings/sec DeepCoNN synthetic Amazon-shape"
# dram__bytes_read.sum + dram__bytes_write.sum of one conv_pool_tc launch at B=4096 (ncu --set full,
# profiles/r1_v19_conv_ncu_raw.csv) and of one whole step (ncu launch list profiles/r1_v19_launches_bench.csv / 21 steps)
NCU_CONV_DRAM_BYTES_PER_LAUNCH = 45.8e6
NCU_STEP_DRAM_BYTES = 206.5e6
REF_SAMPLE_B = 128            # ratings per reference-arm step (the reference's own default batch, hyper_params.py:60)


def algorithmic_bytes_per_rating(hp):
    """bytes that must cross HBM once, as the reference stores them
    (int64 token id + fp32 row per token, two docs; ids + rating + rating out)."""
    T, E = hp["input_length"], hp["word_embed_size"]
    return 2 * T * (8 + 4 * E) + 2 * 8 + 4 + 4


def conv_flops_per_doc(hp):
    return 2.0 * (hp["input_length"] + 2) * 100 * 3 * hp["word_embed_size"]


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            m = json.load(f)
        return float(m["hbm_gbs"]), float(m.get("bf16_tflops_sustained", m["bf16_tflops"])), "measured"
    except Exception:
        return 6650.0, 1400.0, "fallback"


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.path = tempfile.mktemp(suffix=".clocks.csv")
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.QUERY,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1])); mx.append(float(f[2]))
                except ValueError:
                    continue
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


# ------------------------------------------------------------------------------------------ reference arm
def oracle_params_from(model_state):
    return {k: v.detach().to("cpu").clone() for k, v in model_state.items()}


def cpu_train_rate(P, hp, n_steps, warmup, seed, budget_s=None):
    """Times the oracle's restatement of main.train()'s batch body (oracle/r4r_oracle.py::train_batches,
    same ATen CPU kernels the reference dispatches) on REF_SAMPLE_B-rating batches.  Returns (ratings/s, ms/step, steps)."""
    import torch
    from oracle import r4r_oracle as O          # reference arm / cpu_baseline: the one place bench.py runs the oracle
    from reviews4rec_b200.synthetic import SyntheticReader
    hp = dict(hp)
    reader = SyntheticReader(hp, REF_SAMPLE_B, max(1, min(4, n_steps)), V_WORDS, seed=seed)
    batches = reader.batches
    opt = None
    for i in range(warmup):
        _, _, _, opt = O.train_batches(P, [batches[i % len(batches)]], hp, opt=opt)
    t0 = time.perf_counter()
    done = 0
    for i in range(n_steps):
        _, _, _, opt = O.train_batches(P, [batches[i % len(batches)]], hp, opt=opt)
        done += 1
        if budget_s is not None and done >= 3 and time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    return done * REF_SAMPLE_B / dt, dt / done * 1e3, done


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from oracle import r4r_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    hp = dict(HP)
    P = O.init_params(hp, V_WORDS, seed=1)
    rate, ms, steps = cpu_train_rate(P, hp, args.steps, args.warmup, seed=1234)
    sample = "%d-rating batches (reference default batch_size) of the same synthetic workload, %d timed steps" % (REF_SAMPLE_B, steps)
    line = {"impl": "reference", "metric": METRIC, "value": rate, "unit": "ratings/s", "n_gpus": args.gpus, "steps": steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "DeepCoNN E=300 F=100 T=1000 L=10 V=50001 U=1M I=100k (BASELINE configs[1])",
                       "batch": REF_SAMPLE_B, "threads": cores},
            "cpu_baseline": {"value": rate, "unit": "ratings/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": rate, "unit": "ratings/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ B200 arm
def build_model(hp, device, seed):
    import numpy as np
    import torch
    import reviews4rec_b200 as R
    from reviews4rec_b200.utils import xavier_init
    tmp = tempfile.mkdtemp(prefix="r4r_bench_")
    with open(os.path.join(tmp, "word2vec.pkl"), "wb") as f:           # values are overwritten by xavier_init (finding 2)
        pickle.dump(np.zeros((V_WORDS, hp["word_embed_size"]), dtype=np.float32), f, 4)
    hp["data_dir"] = tmp
    torch.manual_seed(seed)
    model = R.DeepCoNN(hp)
    xavier_init(model)                                                  # main.py:377
    return model.to(device)


def run_b200(args):
    import torch
    import torch.distributed as dist
    from reviews4rec_b200 import MSELoss, _lib, ops
    from reviews4rec_b200.optim import FusedAdam
    from reviews4rec_b200.synthetic import SyntheticReader
    from reviews4rec_b200.train import CapturedStep

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    group = None
    if world > 1:
        import faulthandler
        faulthandler.dump_traceback_later(900, exit=True)   # a hung collective must not run forever
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")      # stdout carries exactly one JSON line (NCCL logs its version there otherwise)
        dist.init_process_group("nccl", device_id=dev)
        group = dist.group.WORLD

    hp = dict(HP)
    hp["batch_size"] = args.batch
    B, K, W = args.batch, args.steps, max(args.warmup, 3)
    ops.set_conv_mode(args.conv_mode)
    model = build_model(hp, dev, seed=1)                   # same seed on every rank: replicas start identical
    parallelism = "single GPU"
    if world > 1 or args.force_shard:
        # north_star: word + id tables row-sharded (row r on rank r % P), all-to-all index lookups;
        # ratings split across ranks (weak scaling), dense-gradient all-reduce inside the captured step
        from reviews4rec_b200 import sharded
        tr = sharded.Transport(group)
        wtr = None
        if args.table == "sharded" and args.transport == "p2p" and world > 1:
            wtr = sharded.P2PTransport.for_word_table(V_WORDS, hp["word_embed_size"], group)
        sharded.shard_model(model, tr, shard_word_table=(args.table == "sharded"), word_transport=wtr)
        parallelism = "dp%d, ratings split across ranks; %s; id/bias tables row-sharded; dense-grad all-reduce" % (
            world, ("word table row-sharded r%%P, per-step de-duplicated lookup over %s" % (
                "NVLink peer stores (fused gather + all-to-all kernel)" if wtr else "NCCL all-to-all")) if args.table == "sharded"
            else "frozen word table replicated")
    model.train()
    criterion = MSELoss(hp)
    opt = FusedAdam(model.parameters(), lr=hp["lr"], weight_decay=hp["weight_decay"], capturable=True)

    # ---- synthetic split, generated once on the host (numpy, seeded), as the eight arrays a..h of the reference's
    # quick-data format (make_quick_data.py:21-44); RaggedReader keeps it in pinned host memory
    import numpy as np
    from reviews4rec_b200.readers import RaggedReader
    T = hp["input_length"]
    ragged = args.docs == "ragged"
    # resident pool larger than L2 (126 MB): padded int64 batches are 2*B*T*8 B; ragged ones ~0.2x that
    per_batch = 2 * B * T * 8 * (0.21 if ragged else 1.0)
    pool_n = max(2, min(12, -(-160 * 2 ** 20 // int(per_batch))))
    host = SyntheticReader(hp, B, pool_n, V_WORDS, seed=1234, device=None, pin=False, rank=rank)
    cat = lambda j: np.concatenate([b_[0][j].numpy() for b_ in host.batches])
    arrays = {k: None for k in "abcdefgh"}
    arrays.update(d=cat(3), e=cat(4), f=cat(5), g=cat(6), h=np.concatenate([b_[1].numpy() for b_ in host.batches]))
    # the reader always hands ops.RaggedIdx documents over; --docs decides whether the kernels read them directly
    # ("ragged") or rebuild the padded int64 ids inside the captured step first ("padded", r4r_docs_expand)
    ops.set_ragged_native(ragged)
    rr = RaggedReader(hp, arrays, dev, native=True)

    # ---- device-resident copies of every batch (the `value` measurement reads these)
    class _Res:
        batches = []
    res = _Res()
    resident_bytes = 0
    for bi in range(pool_n):
        hd, hy = host.batches[bi]
        if ragged:
            docs = []
            for k in ("d", "e"):
                tok, off = rr.docs[k].batch_host(bi * B, (bi + 1) * B)
                docs.append(ops.RaggedIdx(tok.to(dev), off.to(dev), (B, T), rr.docs[k].pad_id))
                resident_bytes += tok.numel() * 4 + off.numel() * 8
        else:
            docs = [hd[3].to(dev), hd[4].to(dev)]
            resident_bytes += 2 * B * T * 8
        res.batches.append(([None, None, None, docs[0], docs[1], hd[5].to(dev), hd[6].to(dev)], hy.to(dev)))

    # conv positions the launches actually process (documents cut to their informative prefix, exact)
    pos_sum = 0
    for d, _ in res.batches:
        for idx in (d[3], d[4]):
            pos_sum += int(ops.doc_lengths(idx).sum().item()) + 2 * idx.shape[0]
    positions_per_launch = pos_sum / (2.0 * len(res.batches)) if ops.get_doc_plan() else float(B * (T + 2))

    se_sum = torch.zeros(1, device=dev, dtype=torch.float32)
    conv_events = []
    ops.set_conv_event_sink(conv_events)
    steps_res = [CapturedStep(model, criterion, opt, d, y, se_sum, group, float(world)) for d, y in res.batches]
    launches_per_step = steps_res[0].launches              # this library's kernels inside one captured step
    res_events = list(conv_events)
    ops.set_conv_event_sink(None)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: device-resident inputs
    for i in range(W):
        steps_res[i % pool_n].replay()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    se_sum.zero_()
    torch.cuda.synchronize()
    e0.record()
    for i in range(K):
        steps_res[i % pool_n].replay()
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    train_mse = float(se_sum.item()) / (K * B)
    # dominant kernel: duration of its launches in the last min(K, pool) steps of the timed region
    used = min(K, pool_n)
    conv_ms = []
    try:
        for j in range(used):
            slot = (K - 1 - j) % pool_n
            per = len(res_events) // pool_n                     # eager warm pass + captured pass per slot; the captured pair is last
            for a, b in res_events[per * slot + per - 2: per * slot + per]:
                conv_ms.append(a.elapsed_time(b))
    except Exception as exc:                                       # external event nodes unsupported -> measured below
        conv_ms = []
        conv_err = repr(exc)
    clocks = sampler.stop() if sampler else None
    t = torch.tensor([ms_total], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    value = world * B * K / (ms_total * 1e-3)

    # ---- e2e through the public API: RaggedReader (this repo's counterpart of data_fast.DataLoader) holds the
    # split in pinned HOST memory -- int32 tokens up to each document's trailing padding run -- and per step copies
    # a batch H2D (copy stream, double-buffered) into the buffers the captured step reads (as ops.RaggedIdx
    # documents, or expanded to padded int64 with --docs padded); the running SE sum is read back D2H every step.
    main = torch.cuda.current_stream()
    se_e2e = torch.zeros(1, device=dev, dtype=torch.float32)
    steps_e2e = []
    for s_ in range(2):
        d_, y_, _ = rr.stage(s_, s_)
        rr.wait_ready(s_, main)
        torch.cuda.synchronize()
        steps_e2e.append(CapturedStep(model, criterion, opt, d_, y_, se_e2e, group, float(world)))
        rr.release(s_, main)
    se_host = torch.zeros(max(K, W), dtype=torch.float32).pin_memory()
    h2d_log = []

    def e2e_loop(n):
        for i in range(n):
            s_ = i & 1
            rr.stage(i % pool_n, s_)                              # H2D of the ragged batch + expansion, on the copy stream
            h2d_log.append(rr.h2d_bytes_last)
            rr.wait_ready(s_, main)
            steps_e2e[s_].replay()
            rr.release(s_, main)
            se_host[i:i + 1].copy_(se_e2e, non_blocking=True)     # running SE sum, read back every step (main.py:57)

    def timed(loop):
        loop(W)
        barrier()
        se_e2e.zero_()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e0.record()
        loop(K)
        e1.record()
        barrier()
        wall = (time.perf_counter() - t0) * 1e3
        tt = torch.tensor([max(e0.elapsed_time(e1), 0.0), wall], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt[0].item()), float(tt[1].item())

    e2e_ms, e2e_wall = timed(e2e_loop)
    h2d = int(sum(h2d_log[-K:]) / K)
    e2e_value = world * B * K / (e2e_ms * 1e-3)

    # ---- the same with the batches shipped as the reference's reader does: padded int64 [B,T] from pinned host memory
    if True:
        pstatic = [([None, None, None] + [torch.empty(B, T, device=dev, dtype=torch.int64) for _ in range(2)]
                    + [torch.empty(B, device=dev, dtype=torch.int64) for _ in range(2)],
                    torch.empty(B, device=dev, dtype=torch.float32)) for _ in range(2)]
        for (pd, py), (hd, hy) in zip(pstatic, host.batches):
            for dst, src in zip(pd, hd):
                if dst is not None:
                    dst.copy_(src)
            py.copy_(hy)
        steps_pad = [CapturedStep(model, criterion, opt, pd, py, se_e2e, group, float(world)) for pd, py in pstatic]
        hpin = [([None if x is None else x.pin_memory() for x in hd], hy.pin_memory()) for hd, hy in host.batches[:3]]
    copy_stream = torch.cuda.Stream()
    ready = [torch.cuda.Event() for _ in range(2)]
    done = [torch.cuda.Event() for _ in range(2)]

    def padded_loop(n):
        for i in range(n):
            s_ = i & 1
            hd, hy = hpin[i % len(hpin)]
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(done[s_])
                for dst, src in zip(pstatic[s_][0], hd):
                    if dst is not None:
                        dst.copy_(src, non_blocking=True)
                pstatic[s_][1].copy_(hy, non_blocking=True)
                ready[s_].record(copy_stream)
            main.wait_event(ready[s_])
            steps_pad[s_].replay()
            done[s_].record(main)
            se_host[i:i + 1].copy_(se_e2e, non_blocking=True)

    pad_ms, _ = timed(padded_loop)
    pad_value = world * B * K / (pad_ms * 1e-3)
    t = torch.tensor([e2e_ms, e2e_wall], dtype=torch.float64)

    # ---- dominant kernel timed stand-alone if the in-graph events were unavailable
    timing = "cuda events bracketing the kernel inside the captured step, last %d steps of the timed region" % used
    if not conv_ms and world > 1:
        conv_ms = [float("nan")]
    if not conv_ms:
        d, y = res.batches[0]
        conv = model.user_conv.convs[0]
        with torch.no_grad():
            for i in range(3):
                ops.conv_pool_forward(d[3], model.word2vec.weight, conv.weight, conv.bias, args.conv_mode, model.word2vec._shadow)
            for i in range(min(K, pool_n)):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                sink = []
                ops.set_conv_event_sink(sink)
                ops.conv_pool_forward(res.batches[i][0][3], model.word2vec.weight, conv.weight, conv.bias, args.conv_mode, model.word2vec._shadow)
                ops.set_conv_event_sink(None)
                torch.cuda.synchronize()
                conv_ms.append(sink[0][0].elapsed_time(sink[0][1]))
        timing = "cuda events around stand-alone launches on the bench batches (in-graph events unavailable)"

    if rank != 0:
        _finish(world)
        return

    hbm_peak, tf_peak, peak_kind = measured_peaks()
    conv_avg_ms = sum(conv_ms) / len(conv_ms)
    step_ms = ms_total / K
    E, T = hp["word_embed_size"], hp["input_length"]
    # FLOPs one launch executes: sum over its documents of (informative rows + 2 conv positions) x F x 3E x 2
    # (the padding-run shortcut is exact); algorithmic = the reference's dense conv over all T+2 positions
    exec_flops = positions_per_launch * 100 * 3 * E * 2.0
    alg_flops = B * conv_flops_per_doc(hp)
    tflops = exec_flops / (conv_avg_ms * 1e-3) / 1e12
    alg_bytes_launch = B * T * (8 + 4 * E)                           # one tower = one doc per rating, reference storage
    line = {
        "metric": METRIC, "value": value, "unit": "ratings/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": {"f16": "f16", "bf16": "bf16", "exact": "f32"}[args.conv_mode],
        "data": "synthetic",
        "config": {"workload": "DeepCoNN E=300 F=100 T=1000 L=10 V=50001 U=1M I=100k (BASELINE configs[1])",
                   "batch_per_gpu": B, "global_batch": B * world, "conv_mode": args.conv_mode, "dropout": hp["dropout"],
                   "arithmetic": {"f16": "conv operands f16 (private shadow of the frozen word table + packed filters), fp32 accumulation in TMEM; everything else fp32",
                                  "bf16": "conv operands bf16, fp32 accumulation in TMEM; everything else fp32",
                                  "exact": "fp32 everywhere (CUDA-core conv)"}[args.conv_mode],
                   "parallelism": parallelism,
                   "documents": ("ragged on the device (ops.RaggedIdx: int32 tokens before each trailing padding run + offsets); "
                                 "same padded documents as the reference's reader, never materialised") if ragged
                                else "padded int64 [B,T] as the reference's reader yields them",
                   "l2_policy": "inputs larger than L2: %d resident batches, %.0f MB in total, cycled" % (pool_n, resident_bytes / 2 ** 20),
                   "step": "CUDA graph of zero_grad+forward+MSE+backward+Adam"},
        "e2e": {"value": e2e_value, "unit": "ratings/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                "ms_per_step": e2e_ms / K, "wall_ms_per_step": float(t[1].item()) / K,
                "api": "readers.RaggedReader (pinned host split: int32 tokens before each document's trailing padding run) "
                       "-> H2D -> ops.RaggedIdx -> train.CapturedStep (%s)" % (
                           "kernels read the ragged tokens" if ragged else "r4r_docs_expand rebuilds the padded int64 ids inside the captured step"),
                "padded_int64_reader": {"value": pad_value, "ms_per_step": pad_ms / K, "h2d_bytes_per_step": 2 * B * T * 8 + 2 * B * 8 + B * 4,
                                        "note": "batches shipped as data_fast.py does: padded int64 [B,T] per document, PCIe-bound"}},
        "gpu_launches": launches_per_step * K,
        "roofline": {"kernel": "conv_pool_tc_kernel (fused word gather + TextCNN conv + ReLU + max-pool), tcgen05 cta_group::2",
                     "bound": "tensor", "achieved": tflops, "peak": tf_peak, "unit": "TFLOP/s", "frac": tflops / tf_peak,
                     "traffic": NCU_CONV_DRAM_BYTES_PER_LAUNCH if (B == 4096 and world == 1) else None,
                     "peak_kind": peak_kind + " (cuBLAS bf16, sustained)", "ms_per_launch": conv_avg_ms, "launches_per_step": 2,
                     "share_of_step": 2 * conv_avg_ms / step_ms, "timing": timing,
                     "executed_flops_per_launch": exec_flops, "mean_informative_rows_per_doc": positions_per_launch / B - 2,
                     "algorithmic_flops_per_launch": alg_flops, "algorithmic_tflops": alg_flops / (conv_avg_ms * 1e-3) / 1e12,
                     "algorithmic_bytes_per_launch": alg_bytes_launch,
                     "algorithmic_gbs": alg_bytes_launch / (conv_avg_ms * 1e-3) / 1e9,
                     "note": "achieved = executed FLOPs (exact padding-run shortcut applied) / event time; "
                             "algorithmic_* = the reference's dense work (all T+2 positions, fp32 rows + int64 ids) / the same time"},
        "step_roofline": {"bytes_per_rating": algorithmic_bytes_per_rating(hp),
                          "achieved_gbs": algorithmic_bytes_per_rating(hp) * value / world / 1e9,
                          "frac_of_hbm_peak": algorithmic_bytes_per_rating(hp) * value / world / 1e9 / hbm_peak,
                          "traffic_per_step": NCU_STEP_DRAM_BYTES if (B == 4096 and world == 1) else None,
                          "note": "algorithmic bytes as the reference stores them (fp32 rows + int64 ids, SURVEY.md 8d); the step "
                                  "actually moves traffic_per_step bytes of DRAM (frozen table in half precision, L2-resident), "
                                  "which is why the fraction can exceed 1"},
        "train_mse": train_mse,
        "clocks": clocks,
    }
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        P = oracle_params_from(model.state_dict())
        rate, ms, steps = cpu_train_rate(P, hp, 40, 1, seed=1234, budget_s=15.0)
        line["cpu_baseline"] = {"value": rate, "unit": "ratings/s", "cores": cores, "kind": "port",
                                "sample": "%d timed steps of %d ratings (same synthetic workload), oracle/r4r_oracle.py::train_batches" % (steps, REF_SAMPLE_B)}
    print(json.dumps(line), flush=True)
    _finish(world)


def _finish(world):
    """Multi-rank runs leave without tearing NCCL / symmetric memory down: the captured graphs still hold
    communicator work and ProcessGroupNCCL's destructor can block on it."""
    if world > 1:
        import torch
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=HP["batch_size"], help="ratings per GPU per step")
    ap.add_argument("--conv-mode", default="f16", choices=["f16", "bf16", "exact"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--docs", default="padded", choices=["ragged", "padded"],
                    help="how the reader hands documents to the model: padded int64 tensors rebuilt on the device (default) or ops.RaggedIdx")
    ap.add_argument("--table", default="sharded", choices=["sharded", "replicated"], help="word table placement for --gpus > 1")
    ap.add_argument("--force-shard", action="store_true", help="run the sharded-table path at world size 1 (measures its device-side cost)")
    ap.add_argument("--transport", default="nccl", choices=["nccl", "p2p"], help="how sharded word rows travel")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 and world == 1:
        # convenience: re-launch under torchrun, one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", str(29500 + os.getpid() % 1000)] + sys.argv
        sys.exit(subprocess.call(cmd))
    run_b200(args)


if __name__ == "__main__":
    main()
