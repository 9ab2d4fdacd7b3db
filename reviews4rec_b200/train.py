"""The training driver of the hot path with the reference's signature and metrics
(main.py:8-71 ``train``), plus the restated TransNet step.

``train`` is interchangeable with the reference's own ``main.train`` -- the model classes of this
package also run unmodified under the reference's loop.  Differences, both documented in
SURVEY.md section 8: (1) the per-batch ``float(torch.sum(loss))`` D2H sync (main.py:57) is replaced
by a device-side accumulator read once per epoch; (2) the TransNet branch, which raises on
torch >= 1.5 in the reference (in-place optimizer step between ``backward(retain_graph=True)``
calls), is the exact restatement: three ``autograd.grad`` calls on one graph, then the three
optimizer steps in the reference's order (target, source, source_fm).
"""
import os

import torch

TRANSNET = ("transnet", "transnet++")


def _transnet_param_groups(model, hyper_params):
    src = list(model.source.parameters())
    sfm = list(model.source_fm.parameters())
    if hyper_params["model_type"] == "transnet++":
        sfm += [model.user_embedding.weight, model.item_embedding.weight]
    tgt = [p for p in model.target.parameters() if p.requires_grad]
    return src, sfm, tgt


def transnet_step(model, criterion, optimizer, data, y, hyper_params):
    """One TransNet batch (main.py:35-53 restated).  Returns (per-sample source SE, loss_target, loss_transform)."""
    optimizer_source, optimizer_source_fm, optimizer_target = optimizer[0], optimizer[1], optimizer[2]
    src, sfm, tgt = _transnet_param_groups(model, hyper_params)
    out = model(data)
    loss_target = criterion(out[1], y)
    loss_transform = out[2]
    se_source = criterion(out[0], y, return_mean=False)
    g_t = torch.autograd.grad(loss_target, tgt, retain_graph=True, allow_unused=True)
    g_s = torch.autograd.grad(loss_transform, src, retain_graph=True, allow_unused=True)
    g_f = torch.autograd.grad(torch.mean(se_source), sfm, allow_unused=True)
    for params, grads, opt in ((tgt, g_t, optimizer_target), (src, g_s, optimizer_source), (sfm, g_f, optimizer_source_fm)):
        for p, g in zip(params, grads):
            p.grad = g
        opt.step()
        for p in params:
            p.grad = None
    return se_source.detach(), loss_target.detach(), loss_transform.detach()


def train(model, criterion, optimizer, reader, hyper_params):
    """One epoch over ``reader.iter()``; returns ``{'MSE': round(sum SE / N, 4), ...}`` like main.py:66-71."""
    model.train()
    is_tn = hyper_params["model_type"] in TRANSNET
    dev = next(model.parameters()).device
    acc = torch.zeros(3, device=dev, dtype=torch.float64)       # sum SE, sum loss_target, sum loss_transform
    total_x, total_batches = 0.0, 0.0
    for data, y in reader.iter():
        model.zero_grad()
        if is_tn:
            for o in optimizer:
                o.zero_grad()
            se, lt, lx = transnet_step(model, criterion, optimizer, data, y, hyper_params)
            acc[0] += se.sum()
            acc[1] += lt
            acc[2] += lx
            total_x += float(int(se.shape[0]))
        else:
            optimizer.zero_grad()
            out = model(data)
            loss = criterion(out, y, return_mean=False)
            acc[0] += loss.detach().sum()
            torch.mean(loss).backward()
            optimizer.step()
            total_x += float(int(out.shape[0]))
        total_batches += 1
    sums = acc.tolist()                                          # the only D2H sync of the epoch
    metrics = {"MSE": round(sums[0] / float(total_x), 4)}
    if is_tn:
        metrics["MSE_target"] = round(sums[1] / float(total_batches), 4)
        metrics["MSE_transform"] = round(sums[2] / float(total_batches), 4)
    train.last_raw = {"se_sum": sums[0], "target_sum": sums[1], "transform_sum": sums[2], "n": total_x, "batches": total_batches}
    return metrics


def train_complete(hyper_params, Model, train_reader, val_reader, user_count, item_count, model, review=True,
                   optim_cls=None):
    """``main.train_complete`` (main.py:73-136): MSELoss, Adam (or the four TransNet optimizers of
    utils.init_transnet_optim), ``hyper_params['epochs']`` epochs of train -> validate -> log, the state_dict with
    the best validation MSE saved to ``hyper_params['model_path']``, and that checkpoint reloaded into a fresh
    ``Model(hyper_params)`` which is returned in eval mode.  ``optim_cls`` defaults to this package's FusedAdam
    (same constructor arguments and semantics as the reference's torch.optim.Adam)."""
    import datetime as dt
    import time

    from .eval import evaluate
    from .loss import MSELoss
    from .optim import FusedAdam
    from .utils import file_write, init_transnet_optim, log_end_epoch

    optim_cls = optim_cls or FusedAdam
    log = hyper_params["log_file"]
    file_write(log, "\n\nSimulation run on: " + str(dt.datetime.now()) + "\n\n")
    file_write(log, "Data reading complete!")
    file_write(log, "Number of train batches: {:4d}".format(len(train_reader)))
    file_write(log, "Number of validation batches: {:4d}".format(len(val_reader)))
    criterion = MSELoss(hyper_params)
    if hyper_params["model_type"] in TRANSNET:
        optimizer = init_transnet_optim(hyper_params, model, optim_cls)
    else:
        optimizer = optim_cls(model.parameters(), lr=hyper_params["lr"], weight_decay=hyper_params["weight_decay"])
    file_write(log, str(model))
    file_write(log, "\nModel Built!\nStarting Training...\n")
    device = next(model.parameters()).device
    try:
        best_mse = float("inf")
        for epoch in range(1, hyper_params["epochs"] + 1):
            t0 = time.time()
            metrics = train(model, criterion, optimizer, train_reader, hyper_params)
            metrics, _, _ = evaluate(model, criterion, val_reader, hyper_params, user_count, item_count, review=review)
            metrics["dataset"] = hyper_params.get("dataset")
            log_end_epoch(hyper_params, metrics, epoch, time.time() - t0, metrics_on="(VAL)")
            if metrics["MSE"] < best_mse:                       # main.py:122-126
                print("Saving model...")
                torch.save(model.state_dict(), hyper_params["model_path"])
                best_mse = metrics["MSE"]
    except KeyboardInterrupt:
        print("Exiting from training early")
    best = Model(hyper_params).to(device)                       # main.py:131-134
    best.load_state_dict(torch.load(hyper_params["model_path"], map_location=device))
    best.eval()
    return best


class CapturedStep:
    """One training batch of ``train()`` -- zero_grad, forward, per-sample SE, backward of the mean,
    optional data-parallel gradient all-reduce, optimizer step (main.py:26-60; for TransNet the restated
    three-loss step of ``transnet_step`` with the optimizer triple of ``utils.init_transnet_optim``) --
    recorded ONCE into a CUDA graph over static input buffers and replayed per batch, so a step
    costs one graph launch instead of ~40 Python-driven kernel launches.

    ``data`` / ``y`` are the static device buffers the graph reads (copy each batch into them, or
    build one CapturedStep per resident batch).  ``se_sum`` is a device scalar that every replay
    ADDS the batch's sum of squared errors to (main.py:57 without its per-batch D2H sync).
    The optimizer must be graph-safe: ``FusedAdam(capturable=True)``.
    """

    def __init__(self, model, criterion, optimizer, data, y, se_sum=None, group=None, grad_div=1.0, next_data=None):
        """``next_data``: the static input buffers of the step that will be replayed AFTER this one.  With a
        row-sharded word table the graph then carries that step's word lookup (mark / plan / all-to-all / serve /
        all-to-all / place) on a forked branch next to this step's compute, and this step reads the rows its
        predecessor fetched: the table is frozen, so a lookup depends on the batch's token ids only."""
        self.model, self.data, self.y = model, data, y
        dev = y.device
        self.se_sum = se_sum if se_sum is not None else torch.zeros(1, device=dev, dtype=torch.float32)
        hp = getattr(model, "hyper_params", {})
        is_tn = hp.get("model_type") in TRANSNET
        for o in (optimizer if isinstance(optimizer, (list, tuple)) else [optimizer]):
            if hasattr(o, "prepare"):
                o.prepare()
        words = None
        if next_data is not None and hasattr(model, "word_inputs"):
            from .sharded import ShardedWordTable
            words = next((m for m in model.modules() if isinstance(m, ShardedWordTable)), None)
        if words is not None:
            cur_idx, nxt_idx = model.word_inputs(data), model.word_inputs(next_data)
            cur_slot, nxt_slot = words.reserve(*cur_idx), words.reserve(*nxt_idx)
            if not cur_slot["executed"]:
                words.fill(cur_slot, *cur_idx)          # rows for the eager pass below (a predecessor's fill is only captured so far)
            self._prime = lambda: words.fill(cur_slot, *cur_idx)
        with torch.no_grad():
            model(data)             # eager pass: builds the shadow word table and lazy kernel attributes outside the graph
        model.zero_grad(set_to_none=True)
        side = torch.cuda.Stream(device=dev) if words is not None else None
        # d(mean se)/d se[n] = 1/N, as a constant made once outside the graph (no fill kernel per replay)
        self._gse = torch.full((int(y.numel()),), 1.0 / max(1, int(y.numel())), device=dev, dtype=torch.float32)
        self.graph = torch.cuda.CUDAGraph()
        torch.cuda.synchronize()
        from . import _lib, ops
        launches0 = _lib.launch_count
        if words is not None:
            # the conv kernel owns every SM it runs on; leave 8 of the 74 SM pairs to the forked lookup branch and its
            # NCCL kernels.  Measured at 2 x B200 with the window-stream conv kernel, ms per step for 60 / 63 / 66 / 68 / 70
            # pairs: 1.098 / 1.085 / 1.066 / 1.088 / 1.096 (the faster the conv, the more of the step the branch needs)
            _lib.lib.r4r_conv_set_clusters(int(os.environ.get("R4R_PREFETCH_CLUSTERS", "66")))
        with torch.cuda.graph(self.graph):
            if words is not None:                   # forked branch: the NEXT step's word lookup
                cur = torch.cuda.current_stream()
                side.wait_stream(cur)
                with torch.cuda.stream(side):
                    words.fill(nxt_slot, *nxt_idx)
            ops.arena_begin(dev)                    # one memset for all the zeroed gradient buffers of the step
            if is_tn:
                if group is not None:
                    raise RuntimeError("CapturedStep: the TransNet three-loss step is single-process (replicas only)")
                se, _, _ = transnet_step(model, criterion, optimizer, data, y, hp)
                self.se_sum += se.sum()
                out = None
            else:
                if hasattr(model, "forward_with_loss"):
                    # fused head: rating, squared error and its batch sum come out of one kernel (loss.py:7-11 folded in)
                    out, se = model.forward_with_loss(data, y, self.se_sum)
                    # mean over the batch: the upstream gradient 1/N of every se[n] (no reduction kernels)
                    se.backward(self._gse.view_as(se))
                else:
                    out = model(data)
                    se = criterion(out, y, return_mean=False)
                    self.se_sum += se.detach().sum()
                    torch.mean(se).backward()
                if group is not None:
                    # replicated parameters: mean over ranks; row-sharded tables already received their
                    # rows' gradients from every rank inside the backward (sharded.py)
                    from .sharded import allreduce_dense_grads
                    allreduce_dense_grads(model, group, int(grad_div))
                optimizer.step()
            ops.arena_end()
            if words is not None:
                torch.cuda.current_stream().wait_stream(side)
        if words is not None:
            _lib.lib.r4r_conv_set_clusters(0)
        self.launches = _lib.launch_count - launches0       # kernels of this library one replay launches
        self.out = out

    def prime(self):
        """Fetch this step's word rows now (eagerly): needed once before the first replay of a run whose steps
        prefetch each other's rows, and whenever the static input buffers were refilled behind the graphs' back."""
        if getattr(self, "_prime", None) is not None:
            self._prime()

    def replay(self):
        self.graph.replay()
