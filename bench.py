#!/usr/bin/env python
"""bench.py -- train ratings/sec of the rating-prediction hot path on B200.

    python bench.py [--model M] --gpus N --steps K --warmup W            (this repo's sm_100a path)
    python bench.py [--model M] --impl reference --gpus N --steps K --warmup W   (reference algorithm, host cores)

Default workload = BASELINE.json configs[1]: DeepCoNN (`deepconn`, FM head), word_emb E=300, 100 conv filters,
doc_len T=1000, latent 10, V=50,001 words, 1M users / 100k items, synthetic Amazon-shaped batches
(reviews4rec_b200/synthetic.py).  --model NARRE = configs[3] (R=10 reviews x W=200 words, 500k users / 50k items),
--model transnet++ = configs[4] (three towers, bf16 conv operands, restated three-loss step), --model deepconn++
(MLP head + user / item bias tables under the dense Adam).  A step = one training batch through main.train()'s body:
forward, per-sample squared error, backward, Adam (lr 0.002, weight_decay 1e-6, dropout 0.6).

One JSON line on stdout (rank 0):
  value        ratings/s with the batches already resident in HBM (a pool larger than L2, cycled), CUDA-graph step
  e2e          ratings/s through the public API from pinned HOST batches: H2D of the batch and D2H of the batch's
               squared-error sum are inside the timed region, every step
  eager        the drop-in path a maintainer of the reference would run: train.train(model, MSELoss, Adam, reader, hp)
               (main.train's signature, Python-driven launches, no CUDA graph)
  value_fp32   the same step with the fp32 CUDA-core conv (`exact` mode: reference precision), a few steps;
               .refined = the fp32-refined tensor-core mode (`f16r`: tcgen05 selects the arg-max window, value + gradient fp32)
  precision    largest relative rating error of each conv mode against the fp32 oracle on a 64-rating sample
  roofline     the dominant kernel (fused gather+conv+pool, incl. the two work-list kernels of its launch): executed
               FLOPs per launch / its CUDA-event duration inside the timed region, against MEASURED_PEAKS.json;
               traffic from the committed ncu capture (profiles/ncu_traffic.json)
  cpu_baseline the oracle's CPU restatement of the same step on a bounded sample (N=1 only)
"""
import argparse
import importlib.util
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
BASE_HP = {"latent_size": 10, "word_embed_size": 300, "input_length": 1000, "dropout": 0.6, "lr": 0.002, "weight_decay": 1e-6,
           "narre_num_reviews": 10, "narre_num_words": 200}
# model -> (hyper_params overrides, ratings per GPU per step, default conv mode, workload text)
MODELS = {
    "deepconn": (dict(model_type="deepconn", total_users=1000000, total_items=100000), 4096, "f16",
                 "DeepCoNN E=300 F=100 T=1000 L=10 V=50001 U=1M I=100k (BASELINE configs[1])"),
    "deepconn++": (dict(model_type="deepconn++", total_users=1000000, total_items=100000), 4096, "f16",
                   "DeepCoNN++ (MLP head + user/item bias tables) E=300 F=100 T=1000 L=10 V=50001 U=1M I=100k"),
    "NARRE": (dict(model_type="NARRE", total_users=500000, total_items=50000, only_reviews=False), 2048, "f16",
              "NARRE R=10 reviews x W=200 words, 10 neighbour ids, E=300 F=100 L=10 V=50001 U=500k I=50k (BASELINE configs[3])"),
    "transnet++": (dict(model_type="transnet++", total_users=1000000, total_items=100000), 2048, "bf16",
                   "TransNet++ three TextCNN towers, bf16 conv operands, E=300 F=100 T=1000 L=10 V=50001 U=1M I=100k (BASELINE configs[4])"),
}
REF_SAMPLE_B = 128            # ratings per reference-arm step (the reference's own default batch, hyper_params.py:60)


def model_hp(name):
    hp = dict(BASE_HP)
    hp.update(MODELS[name][0])
    return hp


def metric_name(name):
    return "train ratings/sec %s synthetic Amazon-shape" % {"deepconn": "DeepCoNN", "deepconn++": "DeepCoNN++", "NARRE": "NARRE",
                                                             "transnet++": "TransNet++"}[name]


def towers_of(hp):
    """(documents per rating, rows per document) of the conv launches."""
    mt = hp["model_type"]
    if mt == "NARRE":
        return 2 * hp["narre_num_reviews"], hp["narre_num_words"]
    return (3 if mt.startswith("transnet") else 2), hp["input_length"]


def algorithmic_bytes_per_rating(hp):
    """bytes that must cross HBM once, as the reference stores them (int64 token id + fp32 row per
    token; id rows; ids + rating + rating out)."""
    mt, E, L = hp["model_type"], hp["word_embed_size"], hp["latent_size"]
    docs, T = towers_of(hp)
    words = docs * T * (8 + 4 * E)
    if mt == "NARRE":
        return words + 22 * (8 + 4 * L) + 22 * 2 * 4 * L + 24
    if mt == "transnet++":
        return words + 2 * (8 + 20) + 2 * 40 + 24
    if mt == "deepconn++":
        return words + 2 * (4 + 8) + 24
    return words + 2 * 8 + 4 + 4


def adam_stream_bytes_per_step(hp):
    """SURVEY.md 8(d) per-step dense-Adam term: 7 streams x 4 B over every row of every id table the model trains."""
    mt, L, U, I = hp["model_type"], hp["latent_size"], hp["total_users"] + 2, hp["total_items"] + 2
    if mt == "deepconn++":
        return (U + I) * 4 * 7
    if mt == "NARRE":
        return (U + I) * (L + 1) * 4 * 7
    if mt == "transnet++":
        return (U + I) * 5 * 4 * 7
    return 0


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            m = json.load(f)
        return float(m["hbm_gbs"]), float(m["bf16_tflops"]), float(m.get("bf16_tflops_sustained", m["bf16_tflops"])), "measured"
    except Exception:
        return 6650.0, 1590.0, 1400.0, "fallback"


def ncu_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of ``kernel`` from the committed ncu capture
    (profiles/ncu_traffic.json, written by scripts/ncu_summary.py from an `ncu --set full` run of this bench)."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            return json.load(f)["kernels"][kernel]["dram_bytes_per_launch"]
    except Exception:
        return None


def load_synthetic():
    """reviews4rec_b200/synthetic.py loaded BY PATH: the reference arm must not import the product package
    (its __init__ maps libr4r_b200.so); the generator itself only needs numpy and torch."""
    spec = importlib.util.spec_from_file_location("r4r_synthetic", os.path.join(ROOT, "reviews4rec_b200", "synthetic.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.path = tempfile.mktemp(suffix=".clocks.csv")
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.QUERY,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
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
        sm, mx, pw, reasons = [], [], [], set()
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
                except ValueError:
                    continue
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), power_w_max=max(pw), reasons=sorted(reasons), samples=len(sm))
        return out


# ------------------------------------------------------------------------------------------ reference arm
def cpu_train_rate(P, hp, n_steps, warmup, seed, budget_s=None):
    """Times the oracle's restatement of main.train()'s batch body (oracle/r4r_oracle.py, the same ATen CPU kernels
    the reference dispatches) on REF_SAMPLE_B-rating batches.  Returns (ratings/s, ms/step, steps)."""
    from oracle import r4r_oracle as O          # reference arm / cpu_baseline: the one place bench.py runs the oracle
    S = load_synthetic()
    hp = dict(hp)
    batches = S.SyntheticReader(hp, REF_SAMPLE_B, max(1, min(4, n_steps)), V_WORDS, seed=seed).batches
    is_tn = hp["model_type"].startswith("transnet")
    state = {"opt": None}

    def one(b):
        if is_tn:
            state["opt"] = O.transnet_train(P, [b], hp, opts=state["opt"])[3]
        else:
            state["opt"] = O.train_batches(P, [b], hp, opt=state["opt"])[3]

    for i in range(warmup):
        one(batches[i % len(batches)])
    t0 = time.perf_counter()
    done = 0
    for i in range(n_steps):
        one(batches[i % len(batches)])
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
    hp = model_hp(args.model)
    P = O.init_params(hp, V_WORDS, seed=1)
    rate, ms, steps = cpu_train_rate(P, hp, args.steps, args.warmup, seed=1234, budget_s=150.0)
    sample = "%d-rating batches (reference default batch_size) of the same synthetic workload, %d timed steps" % (REF_SAMPLE_B, steps)
    line = {"impl": "reference", "metric": metric_name(args.model), "value": rate, "unit": "ratings/s", "n_gpus": args.gpus, "steps": steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": MODELS[args.model][3], "batch": REF_SAMPLE_B, "threads": cores,
                       "implementation": "oracle/r4r_oracle.py (functional restatement of the reference's nn.Modules + main.train, pinned to "
                                         "golden vectors of the unmodified reference; /root/reference cannot travel to the GPU box)"},
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
    mt = hp["model_type"]
    cls = R.NARRE if mt == "NARRE" else R.TransNet if mt.startswith("transnet") else R.DeepCoNN
    model = cls(hp)
    xavier_init(model)                                                  # main.py:377
    return model.to(device)


def make_optimizer(model, hp, capturable):
    from reviews4rec_b200.optim import FusedAdam
    from reviews4rec_b200.utils import init_transnet_optim
    if hp["model_type"].startswith("transnet"):
        cls = (lambda params, **kw: FusedAdam(params, capturable=capturable, **kw))
        return init_transnet_optim(hp, model, cls)
    return FusedAdam(model.parameters(), lr=hp["lr"], weight_decay=hp["weight_decay"], capturable=capturable)


def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from reviews4rec_b200 import MSELoss, ops
    from reviews4rec_b200.readers import RaggedReader
    from reviews4rec_b200.synthetic import SyntheticReader
    from reviews4rec_b200.train import CapturedStep, train

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

    hp = model_hp(args.model)
    mt = hp["model_type"]
    is_tn = mt.startswith("transnet")
    B = args.batch or MODELS[args.model][1]
    conv_mode = args.conv_mode or MODELS[args.model][2]
    hp["batch_size"] = B
    if args.full_length:
        hp["synthetic_full_length"] = True
    K, W = args.steps, max(args.warmup, 3)
    ops.set_conv_mode(conv_mode)
    model = build_model(hp, dev, seed=1)                   # same seed on every rank: replicas start identical
    parallelism = "single GPU"
    if (world > 1 or args.force_shard) and not is_tn:
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
    elif world > 1:
        # TransNet's three optimizers step three disjoint parameter groups from one forward graph (main.py:35-53 restated);
        # that schedule has no cross-rank exchange in it, so N GPUs run N independent replicas
        group = None
        parallelism = "replicas only: %d independent single-GPU training runs (no data-path collective)" % world
    model.train()
    criterion = MSELoss(hp)
    opt = make_optimizer(model, hp, capturable=True)

    # ---- synthetic split, generated once on the host (numpy, seeded), as the eight arrays a..h of the reference's
    # quick-data format (make_quick_data.py:21-44); RaggedReader keeps it in pinned host memory
    docs_per_rating, T = towers_of(hp)
    ragged = args.docs == "ragged"
    per_batch = docs_per_rating * B * T * 8 * (0.21 if ragged else 1.0)
    pool_n = max(2, min(12, -(-160 * 2 ** 20 // int(per_batch))))      # resident pool larger than L2 (126 MB)
    host = SyntheticReader(hp, B, pool_n, V_WORDS, seed=1234, device=None, pin=False, rank=rank)
    cat = lambda j: None if host.batches[0][0][j] is None else np.concatenate([b_[0][j].numpy() for b_ in host.batches])
    arrays = dict(zip("abcdefg", [cat(j) for j in range(7)]))
    arrays["h"] = np.concatenate([b_[1].numpy() for b_ in host.batches])
    # the reader always hands ops.RaggedIdx documents over; --docs decides whether the kernels read them directly
    # ("ragged") or rebuild the padded int64 ids inside the captured step first ("padded", r4r_docs_expand)
    ops.set_ragged_native(ragged)
    rr = RaggedReader(hp, arrays, dev, native=True)
    doc_slots = [j for j in (0, 3, 4) if host.batches[0][0][j] is not None]

    # ---- device-resident copies of every batch (the `value` measurement reads these)
    res_batches, resident_bytes = [], 0
    for bi in range(pool_n):
        hd, hy = host.batches[bi]
        data = [None if x is None else x.to(dev) for x in hd]
        if ragged:
            for j, k in ((0, "a"), (3, "d"), (4, "e")):
                if hd[j] is not None:
                    tok, off = rr.docs[k].batch_host(bi * B, (bi + 1) * B)
                    data[j] = ops.RaggedIdx(tok.to(dev), off.to(dev), tuple(hd[j].shape), rr.docs[k].pad_id)
                    resident_bytes += tok.numel() * 4 + off.numel() * 8
        else:
            resident_bytes += sum(hd[j].numel() * 8 for j in doc_slots)
        res_batches.append((data, hy.to(dev)))

    # conv positions the launches actually process (documents cut to their informative prefix, exact)
    pos_sum, n_docs = 0, 0
    plan_on = ops.get_doc_plan()
    for d, _ in res_batches:
        for j in doc_slots:
            idx = d[j] if isinstance(d[j], ops.RaggedIdx) else d[j].reshape(-1, T)
            rows = int(idx.shape[0]) if not isinstance(idx, ops.RaggedIdx) else int(np.prod(idx.shape[:-1]))
            pos_sum += (int(ops.doc_lengths(idx).sum().item()) if plan_on else rows * T) + 2 * rows
            n_docs += rows
    n_towers = len(doc_slots)
    positions_per_launch = pos_sum / float(n_towers * len(res_batches))
    docs_per_launch = n_docs / float(n_towers * len(res_batches))

    se_sum = torch.zeros(1, device=dev, dtype=torch.float32)
    conv_events = []
    ops.set_conv_event_sink(conv_events)
    # sharded word table: every step's graph also carries the NEXT step's word lookup on a forked branch (the table is
    # frozen, the lookup depends on token ids only), so the all-to-all latency hides behind this step's conv
    prefetch = (world > 1 or args.force_shard) and not is_tn and args.table == "sharded" and not args.no_prefetch
    steps_res = [CapturedStep(model, criterion, opt, d, y, se_sum, group, float(world),
                              next_data=res_batches[(bi + 1) % pool_n][0] if prefetch else None)
                 for bi, (d, y) in enumerate(res_batches)]
    launches_per_step = steps_res[0].launches              # this library's kernels inside one captured step
    res_events = list(conv_events)
    ops.set_conv_event_sink(None)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_max(*vals):
        t = torch.tensor(list(vals), device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(x) for x in t.tolist()]

    # ---- value: device-resident inputs
    steps_res[0].prime()
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
    (ms_total,) = reduce_max(e0.elapsed_time(e1))
    train_mse = float(se_sum.item()) / (K * B)
    value = world * B * K / (ms_total * 1e-3)
    # dominant kernel: duration of its launches in the last min(K, pool) steps of the timed region
    used = min(K, pool_n)
    conv_ms = []
    try:
        per = len(res_events) // pool_n                         # eager warm pass + captured pass per slot; the captured launches are last
        for j in range(used):
            slot = (K - 1 - j) % pool_n
            for a, b in res_events[per * slot + per - n_towers: per * slot + per]:
                conv_ms.append(a.elapsed_time(b))
    except Exception:                                             # external event nodes unsupported -> measured below
        conv_ms = []

    # ---- e2e through the public API: RaggedReader (this repo's counterpart of data_fast.DataLoader) holds the
    # split in pinned HOST memory -- int32 tokens up to each document's trailing padding run -- and per step copies
    # a batch H2D (copy stream, double-buffered) into the buffers the captured step reads; the running SE sum is
    # read back D2H every step.
    main = torch.cuda.current_stream()
    se_e2e = torch.zeros(1, device=dev, dtype=torch.float32)
    n_slots = 3 if prefetch else 2
    if prefetch:
        rr = RaggedReader(hp, arrays, dev, slots=3, native=True)
    staged = []
    for s_ in range(n_slots):
        d_, y_, _ = rr.stage(s_, s_)
        rr.wait_ready(s_, main)
        staged.append((d_, y_))
    torch.cuda.synchronize()
    steps_e2e = [CapturedStep(model, criterion, opt, d_, y_, se_e2e, group, float(world),
                              next_data=staged[(s_ + 1) % n_slots][0] if prefetch else None) for s_, (d_, y_) in enumerate(staged)]
    for s_ in range(n_slots):
        rr.release(s_, main)
    se_host = torch.zeros(max(K, W), dtype=torch.float32).pin_memory()
    h2d_log = []

    def e2e_loop_pair(n):
        for i in range(n):
            s_ = i & 1
            rr.stage(i % pool_n, s_)                              # H2D of the ragged batch, on the copy stream
            h2d_log.append(rr.h2d_bytes_last)
            rr.wait_ready(s_, main)
            steps_e2e[s_].replay()
            rr.release(s_, main)
            se_host[i:i + 1].copy_(se_e2e, non_blocking=True)     # running SE sum, read back every step (main.py:57)

    def e2e_loop_prefetch(n):
        # three slots: step i computes on slot i % 3 (rows fetched by step i-1's forked branch) and looks up the batch
        # in slot (i+1) % 3, whose H2D overlapped step i-1; the H2D of batch i+2 overlaps this step
        rr.stage(0, 0)
        h2d_log.append(rr.h2d_bytes_last)
        rr.wait_ready(0, main)
        steps_e2e[0].prime()                                      # the run's first batch has no predecessor
        rr.stage(1 % pool_n, 1)
        h2d_log.append(rr.h2d_bytes_last)
        for i in range(n):
            s_ = i % 3
            rr.wait_ready((s_ + 1) % 3, main)                     # batch i+1 has landed
            rr.stage((i + 2) % pool_n, (s_ + 2) % 3)              # waits (copy stream) for step i-1's release of that slot
            h2d_log.append(rr.h2d_bytes_last)
            steps_e2e[s_].replay()
            rr.release(s_, main)
            se_host[i:i + 1].copy_(se_e2e, non_blocking=True)

    e2e_loop = e2e_loop_prefetch if prefetch else e2e_loop_pair

    def timed(loop, n):
        loop(min(W, n))
        barrier()
        se_e2e.zero_()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e0.record()
        loop(n)
        e1.record()
        barrier()
        wall = (time.perf_counter() - t0) * 1e3
        return reduce_max(max(e0.elapsed_time(e1), 0.0), wall)

    e2e_ms, e2e_wall = timed(e2e_loop, K)
    h2d = int(sum(h2d_log[-K:]) / K)
    e2e_value = world * B * K / (e2e_ms * 1e-3)

    # ---- the drop-in path: the eager loop with main.train's signature over device-resident batches (what a
    # maintainer gets from the import swap of INTEGRATION.md alone: Python-driven launches, no CUDA graph)
    class _Pool:
        def __init__(self, n):
            self.n = n

        def __len__(self):
            return self.n

        def iter(self, eval=False):
            for i in range(self.n):
                yield res_batches[i % pool_n]

    eager = None
    if not args.no_eager:
        Ke = max(3, min(K, 60))
        opt_e = make_optimizer(model, hp, capturable=False)
        train(model, criterion, opt_e, _Pool(3), hp)
        barrier()
        e0.record()
        m_e = train(model, criterion, opt_e, _Pool(Ke), hp)
        e1.record()
        barrier()
        (eager_ms,) = reduce_max(e0.elapsed_time(e1))
        eager = {"value": world * B * Ke / (eager_ms * 1e-3), "unit": "ratings/s", "ms_per_step": eager_ms / Ke, "steps": Ke,
                 "api": "reviews4rec_b200.train.train(model, MSELoss, FusedAdam, reader, hyper_params) -- main.train's signature (main.py:8-71), "
                        "eager Python loop, device-resident batches", "train_mse": m_e["MSE"]}
    clocks = sampler.stop() if sampler else None

    # ---- reference precision: the same captured step with the fp32 CUDA-core conv (`exact` mode), a few steps;
    # and the largest relative rating error of each conv mode against the fp32 oracle on a 64-rating sample
    value_fp32, precision = None, None
    if world == 1 and not args.no_fp32 and conv_mode != "exact":
        ops.set_conv_mode("exact")
        opt32 = make_optimizer(model, hp, capturable=True)
        se32 = torch.zeros(1, device=dev, dtype=torch.float32)
        steps32 = [CapturedStep(model, criterion, opt32, d, y, se32, None, 1.0) for d, y in res_batches[:2]]
        K32 = 4
        steps32[0].replay()
        torch.cuda.synchronize()
        e0.record()
        for i in range(K32):
            steps32[i % 2].replay()
        e1.record()
        torch.cuda.synchronize()
        ms32 = e0.elapsed_time(e1) / K32
        value_fp32 = {"value": B / (ms32 * 1e-3), "unit": "ratings/s", "ms_per_step": ms32, "steps": K32, "dtype": "f32",
                      "note": "conv on CUDA cores in fp32 (r4r_conv_pool_simt), everything else identical; device-resident inputs"}
        del steps32
        # fp32-grade at tensor-core speed: the tcgen05 kernel selects each filter's arg-max window, r4r_conv_refine re-evaluates
        # it in fp32 from the fp32 table and filters, and the weight gradient is the fp32 one (conv mode "f16r")
        ops.set_conv_mode("f16r")
        optr = make_optimizer(model, hp, capturable=True)
        stepsr = [CapturedStep(model, criterion, optr, d, y, se32, None, 1.0) for d, y in res_batches]
        Kr = max(3, min(K, 40))
        for i in range(3):
            stepsr[i % pool_n].replay()
        torch.cuda.synchronize()
        e0.record()
        for i in range(Kr):
            stepsr[i % pool_n].replay()
        e1.record()
        torch.cuda.synchronize()
        msr = e0.elapsed_time(e1) / Kr
        value_fp32["refined"] = {"value": B / (msr * 1e-3), "unit": "ratings/s", "ms_per_step": msr, "steps": Kr, "conv_mode": "f16r",
                                 "note": "tensor cores select the arg-max window, its value and the weight gradient are fp32 (r4r_conv_refine)"}
        del stepsr
        ops.set_conv_mode(conv_mode)
    if world == 1 and not args.no_cpu_baseline:
        from oracle import r4r_oracle as O          # checker leg (precision + cpu_baseline), never the product path
        P = {k: v.detach().to("cpu").clone() for k, v in model.state_dict().items()}
        hd, hy = host.batches[0]
        sample = [None if x is None else x[:64] for x in hd]
        with torch.no_grad():
            ref = O.forward(P, sample, hp, train=False)
            ref = ref[0] if isinstance(ref, list) else ref
            precision = {}
            model.eval()
            for m in ("exact", "f16r", "f16", "bf16"):
                ops.set_conv_mode(m)
                out = model([None if x is None else x.to(dev) for x in sample])
                out = (out[0] if isinstance(out, list) else out).cpu()
                precision[m] = float(((out - ref).abs() / ref.abs()).max())
            model.train()
            ops.set_conv_mode(conv_mode)
        precision["sample"] = "64 ratings of the bench workload, eval forward, trained parameters of this run, vs oracle fp32 on the host"

    # ---- dominant kernel timed stand-alone if the in-graph events were unavailable
    timing = "cuda events bracketing the kernel inside the captured step, last %d steps of the timed region" % used
    if not conv_ms:
        conv_ms = [float("nan")]
        timing = "unavailable (in-graph event nodes not supported here)"

    if rank != 0:
        _finish(world)
        return

    hbm_peak, tf_burst, tf_sustained, peak_kind = measured_peaks()
    conv_avg_ms = sum(conv_ms) / len(conv_ms)
    step_ms = ms_total / K
    E = hp["word_embed_size"]
    # FLOPs one launch executes: sum over its documents of (informative rows + 2 conv positions) x F x 3E x 2
    # (the padding-run shortcut is exact); algorithmic = the reference's dense conv over all T+2 positions
    exec_flops = positions_per_launch * 100 * 3 * E * 2.0
    alg_flops = docs_per_launch * 2.0 * (T + 2) * 100 * 3 * E
    tflops = exec_flops / (conv_avg_ms * 1e-3) / 1e12
    alg_bytes_launch = docs_per_launch * T * (8 + 4 * E)             # one launch = one tower, reference storage
    conv_kernel = "conv_pool_simt_kernel" if conv_mode == "exact" else "conv_pool_tc_kernel"
    bpr, adam_b = algorithmic_bytes_per_rating(hp), adam_stream_bytes_per_step(hp)
    step_gbs = (bpr * value / world + adam_b * (value / world / B)) / 1e9
    line = {
        "metric": metric_name(args.model), "value": value, "unit": "ratings/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": {"f16": "f16", "bf16": "bf16", "exact": "f32"}[conv_mode],
        "data": "synthetic",
        "config": {"workload": MODELS[args.model][3] + (" -- FULL-LENGTH documents (no padding; not the SURVEY 8d shape)" if args.full_length else ""),
                   "batch_per_gpu": B, "global_batch": B * world, "conv_mode": conv_mode, "dropout": hp["dropout"],
                   "arithmetic": {"f16": "conv operands f16 (private shadow of the frozen word table + packed filters), fp32 accumulation in TMEM; everything else fp32",
                                  "bf16": "conv operands bf16, fp32 accumulation in TMEM; everything else fp32",
                                  "exact": "fp32 everywhere (CUDA-core conv)"}[conv_mode],
                   "parallelism": parallelism,
                   "documents": ("ragged on the device (ops.RaggedIdx: int32 tokens before each trailing padding run + offsets); "
                                 "same padded documents as the reference's reader, never materialised") if ragged
                                else "padded int64 as the reference's reader yields them",
                   "l2_policy": "inputs larger than L2: %d resident batches, %.0f MB in total, cycled" % (pool_n, resident_bytes / 2 ** 20),
                   "step": "CUDA graph of zero_grad+forward+MSE+backward+Adam" + (" (restated three-loss TransNet step, three optimizers)" if is_tn else "")},
        "e2e": {"value": e2e_value, "unit": "ratings/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                "ms_per_step": e2e_ms / K, "wall_ms_per_step": e2e_wall / K,
                "api": "readers.RaggedReader (pinned host split: int32 tokens before each document's trailing padding run) "
                       "-> H2D -> ops.RaggedIdx -> train.CapturedStep (%s)" % (
                           "kernels read the ragged tokens" if ragged else "r4r_docs_expand rebuilds the padded int64 ids inside the captured step")},
        "eager": eager,
        "value_fp32": value_fp32,
        "precision": precision,
        "gpu_launches": launches_per_step * K,
        "roofline": {"kernel": "%s (fused word gather + TextCNN conv + ReLU + max-pool)%s" % (
                         conv_kernel, "" if conv_mode == "exact" else ", tcgen05 cta_group::2"),
                     "bound": "tensor", "achieved": tflops, "peak": tf_burst, "unit": "TFLOP/s", "frac": tflops / tf_burst,
                     "traffic": ncu_traffic(conv_kernel) if (args.model == "deepconn" and B == 4096 and world == 1) else None,
                     "peak_kind": peak_kind + " (cuBLAS bf16 burst: the timed region is far shorter than the power-capped sustained regime)",
                     "frac_of_sustained_peak": tflops / tf_sustained,
                     "ms_per_launch": conv_avg_ms, "launches_per_step": n_towers,
                     "share_of_step": n_towers * conv_avg_ms / step_ms, "timing": timing,
                     "executed_flops_per_launch": exec_flops, "mean_informative_rows_per_doc": positions_per_launch / docs_per_launch - 2,
                     "algorithmic_flops_per_launch": alg_flops, "algorithmic_tflops": alg_flops / (conv_avg_ms * 1e-3) / 1e12,
                     "algorithmic_bytes_per_launch": alg_bytes_launch,
                     "algorithmic_gbs": alg_bytes_launch / (conv_avg_ms * 1e-3) / 1e9,
                     "note": "achieved = executed FLOPs (exact padding-run shortcut applied) / event time; "
                             "algorithmic_* = the reference's dense work (all T+2 positions, fp32 rows + int64 ids) / the same time"},
        "step_roofline": {"bytes_per_rating": bpr, "adam_stream_bytes_per_step": adam_b,
                          "achieved_gbs": step_gbs, "frac_of_hbm_peak": step_gbs / hbm_peak,
                          "traffic_per_step": ncu_traffic("__step__") if (args.model == "deepconn" and B == 4096 and world == 1) else None,
                          "note": "algorithmic bytes as the reference stores them (fp32 rows + int64 ids, SURVEY.md 8d) plus the dense-Adam stream; "
                                  "the step actually moves traffic_per_step bytes of DRAM (frozen table in half precision, L2-resident), "
                                  "which is why the fraction can exceed 1"},
        "train_mse": train_mse,
        "clocks": clocks,
    }
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        P = {k: v.detach().to("cpu").clone() for k, v in model.state_dict().items()}
        rate, ms, steps = cpu_train_rate(P, hp, 40, 1, seed=1234, budget_s=15.0)
        line["cpu_baseline"] = {"value": rate, "unit": "ratings/s", "cores": cores, "kind": "port",
                                "sample": "%d timed steps of %d ratings (same synthetic workload), oracle/r4r_oracle.py" % (steps, REF_SAMPLE_B)}
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
    ap.add_argument("--model", default="deepconn", choices=sorted(MODELS))
    ap.add_argument("--batch", type=int, default=0, help="ratings per GPU per step (default: per model)")
    ap.add_argument("--conv-mode", default=None, choices=["f16", "bf16", "exact"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-eager", action="store_true")
    ap.add_argument("--no-fp32", action="store_true")
    ap.add_argument("--full-length", action="store_true",
                    help="synthetic documents without padding (every document has T informative rows): the step when the exact "
                         "padding-run shortcut has nothing to skip; NOT the SURVEY 8(d) workload")
    ap.add_argument("--docs", default="padded", choices=["ragged", "padded"],
                    help="how the reader hands documents to the model: padded int64 tensors rebuilt on the device (default) or ops.RaggedIdx")
    ap.add_argument("--table", default="sharded", choices=["sharded", "replicated"], help="word table placement for --gpus > 1")
    ap.add_argument("--force-shard", action="store_true", help="run the sharded-table path at world size 1 (measures its device-side cost)")
    ap.add_argument("--transport", default="nccl", choices=["nccl", "p2p"], help="how sharded word rows travel")
    ap.add_argument("--no-prefetch", action="store_true", help="sharded word table: look rows up inside the step that uses them (round-1 behaviour)")
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
