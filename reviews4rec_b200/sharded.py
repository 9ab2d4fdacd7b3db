"""Row-sharded word / id tables across the GPUs of one box (BASELINE.json north_star).

The reference is single-process (no ``torch.distributed`` call anywhere), so the contract
here is "the same numbers as the single-process reference at the same global batch":

* ratings are independent, so the batch is split across ranks (``shard_batch``); the loss mean is over
  the global batch, hence gradients of replicated parameters are averaged over ranks
  (``allreduce_dense_grads``) and gradients arriving at a table shard are scaled by ``1 / world``;
* row ``r`` of a table lives on rank ``r % P`` at local row ``r // P`` (spreads the Zipf head and the pad
  ids); the local shard has ``ceil(R / P)`` rows;
* the frozen word table (DeepCoNN.py:15) is looked up forward-only: per step the rank de-duplicates the
  token ids of its documents, asks the owners for the rows and receives a compact per-step row cache that
  the unchanged conv / wgrad kernels read through their ORIGINAL token ids (the cache is indexed by id);
* the trainable id tables / bias vectors (``nn.Embedding(sparse=False)`` / ``Tensor.gather`` at
  MF.py:45-53, NARRE.py:87-88,110-116, TransNet.py:108-109, DeepCoNN.py:70-71) get rows back per id, send
  row gradients to the owners in the backward, and every local row is then updated by the dense Adam
  exactly as the reference updates every row of the full table (SURVEY.md finding 5).

Transport: NCCL ``all_to_all_single`` with equal splits (no host sync, CUDA-graph capturable), or --
``P2PTransport`` -- the served rows are written straight into the requester's symmetric-memory receive
buffer by ``r4r_shard_serve_p2p`` over NVLink (the gather is the all-to-all).

Device work goes through ``K`` (C-ABI kernels of csrc/shard.cu); the CPU tests of the protocol replace
``K`` with an emulation that lives in ``tests/``.
"""
import ctypes
from typing import List, Optional, Sequence

import torch
import torch.distributed as dist
import torch.nn as nn

from ._lib import call
from .ops import _p, _stream


# ------------------------------------------------------------------------------------------ layout
def rows_local(R: int, P: int) -> int:
    return (R + P - 1) // P


def shard_rows(full: torch.Tensor, rank: int, P: int) -> torch.Tensor:
    """Rows ``rank, rank+P, ...`` of ``full`` ([R] or [R, L]), zero-padded to ``ceil(R/P)`` rows."""
    R = full.shape[0]
    out = full.new_zeros((rows_local(R, P),) + tuple(full.shape[1:]))
    mine = full[rank::P]
    out[: mine.shape[0]] = mine
    return out


def unshard_rows(shards: Sequence[torch.Tensor], R: int) -> torch.Tensor:
    """Inverse of ``shard_rows`` given the shards of ranks 0..P-1."""
    P = len(shards)
    full = shards[0].new_empty((R,) + tuple(shards[0].shape[1:]))
    for r, s in enumerate(shards):
        n = len(range(r, R, P))
        full[r::P] = s[:n]
    return full


def shard_batch(data, y, rank: int, P: int):
    """Contiguous 1/P slice of a reader batch (the last ranks get the short remainder)."""
    B = y.shape[0]
    per = (B + P - 1) // P
    lo, hi = min(B, rank * per), min(B, (rank + 1) * per)
    return [None if d is None else d[lo:hi] for d in data], y[lo:hi]


# ------------------------------------------------------------------------------------------ kernels
class _DeviceKernels:
    """Thin wrappers over the C ABI (include/r4r_b200.h, K8)."""

    @staticmethod
    def mark(idx, V, flags):
        call("r4r_shard_mark", _p(idx), idx.numel(), V, _p(flags), _stream())

    @staticmethod
    def plan(flags, V, P, cap, req):
        call("r4r_shard_plan", _p(flags), V, P, cap, _p(req), _stream())

    @staticmethod
    def bucket(ids, R, P, cap, req, pos):
        call("r4r_shard_bucket", _p(ids), ids.numel(), R, P, cap, _p(req), _p(pos), _stream())

    @staticmethod
    def serve(shard, rreq, P, cap, out):
        row_bytes = (shard.numel() // shard.shape[0]) * shard.element_size()
        call("r4r_shard_serve", _p(shard), shard.shape[0], row_bytes, _p(rreq), P, cap, _p(out), _stream())

    @staticmethod
    def serve_p2p(shard, rreq, P, cap, out_ptrs):
        row_bytes = (shard.numel() // shard.shape[0]) * shard.element_size()
        arr = (ctypes.c_void_p * P)(*out_ptrs)
        call("r4r_shard_serve_p2p", _p(shard), shard.shape[0], row_bytes, _p(rreq), P, cap,
             ctypes.cast(arr, ctypes.c_void_p), _stream())

    @staticmethod
    def place(rows, req, P, cap, cache, V):
        row_bytes = (cache.numel() // cache.shape[0]) * cache.element_size()
        call("r4r_shard_place", _p(rows), _p(req), P, cap, row_bytes, _p(cache), V, _stream())

    @staticmethod
    def gather(table, pos, out):
        L = table.numel() // table.shape[0]
        call("r4r_rows_gather", _p(table), table.shape[0], L, _p(pos), pos.numel(), _p(out), _stream())

    @staticmethod
    def scatter_unique(gout, pos, send):
        L = send.numel() // send.shape[0]
        call("r4r_rows_scatter_add", _p(gout), _p(pos), pos.numel(), L, _p(send), send.shape[0], _stream())

    @staticmethod
    def scatter_owner(grads, rreq, P, cap, gtable, scale):
        L = gtable.numel() // gtable.shape[0]
        call("r4r_shard_scatter_add", _p(grads), _p(rreq), P, cap, L, _p(gtable), gtable.shape[0], float(scale), _stream())


K = _DeviceKernels()


# ------------------------------------------------------------------------------------------ transport
class Transport:
    """Equal-split all-to-all between the ranks of ``group`` (block q of ``inp`` goes to rank q)."""

    def __init__(self, group=None):
        self.group = group
        self.world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        self.rank = dist.get_rank(group) if self.world > 1 else 0

    def all_to_all(self, out: torch.Tensor, inp: torch.Tensor) -> torch.Tensor:
        if self.world == 1:
            out.view(-1).copy_(inp.view(-1))
        else:
            dist.all_to_all_single(out.view(-1), inp.view(-1), group=self.group)
        return out

    def agree_cap(self, n: int, device) -> int:
        """max over ranks of ``n`` (one tiny all-reduce + a host read: not CUDA-graph capturable)."""
        if self.world == 1:
            return n
        t = torch.tensor([n], device=device, dtype=torch.int64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        return int(t.item())

    def exchange_rows(self, shard, rreq, P, cap, row_shape, dtype):
        """Owner side of a lookup + the rows' way back: returns the [P * cap, *row_shape] rows this rank
        asked for, compact: block q holds the rows requested of owner q in request order."""
        payload = torch.empty((P, cap) + tuple(row_shape), device=shard.device, dtype=dtype)
        K.serve(shard, rreq, P, cap, payload)
        out = torch.empty((P * cap,) + tuple(row_shape), device=shard.device, dtype=dtype)
        self.all_to_all(out, payload)
        return out


class P2PTransport(Transport):
    """Rows travel by peer stores: every rank owns a symmetric-memory receive buffer; the owner's serve
    kernel writes requester q's block directly at ``peer_buffer[q] + my_rank * block_bytes`` over
    NVLink, so only requested rows cross the links and no staging copy exists.  Requests (a few KB)
    still use the NCCL all-to-all.  Needs ``torch.distributed._symmetric_memory`` (CUDA, world > 1)."""

    def __init__(self, group=None, max_block_bytes: int = 0):
        super().__init__(group)
        import torch.distributed._symmetric_memory as symm
        self._symm = symm
        self.block_cap = int(max_block_bytes)
        self.buf = symm.empty(self.world * self.block_cap + 4096, dtype=torch.uint8, device=torch.device("cuda", torch.cuda.current_device()))
        self.hdl = symm.rendezvous(self.buf, group or dist.group.WORLD)
        self.peer_ptrs = [int(p) for p in self.hdl.buffer_ptrs]

    @classmethod
    def for_word_table(cls, V: int, E: int, group=None):
        """Sized for either payload of a sharded word lookup: fp32 rows (exact mode) or Epad-wide half rows."""
        world = dist.get_world_size(group)
        epad = ((E + 63) // 64) * 64
        return cls(group, max_block_bytes=rows_local(V, world) * max(E * 4, epad * 2))

    def exchange_rows(self, shard, rreq, P, cap, row_shape, dtype):
        n_row = 1
        for d in row_shape:
            n_row *= d
        row_bytes = n_row * torch.empty((), dtype=dtype).element_size()
        block = cap * row_bytes
        if block > self.block_cap:
            raise RuntimeError("P2PTransport: block of %d bytes exceeds the symmetric buffer (%d per peer)" % (block, self.block_cap))
        self.hdl.barrier(channel=0)                      # every rank has consumed the previous contents of its buffer
        ptrs = [self.peer_ptrs[q] + self.rank * block for q in range(P)]
        K.serve_p2p(shard, rreq, P, cap, ptrs)
        self.hdl.barrier(channel=1)                      # all peers' stores into my buffer have landed
        # A VIEW of the symmetric buffer, valid until this rank's NEXT exchange_rows: peers only write after the
        # channel-0 barrier of that call, i.e. after this rank's stream has finished everything queued before it.
        # Both consumers copy out at once on the same stream (ShardedWordTable._lookup places the rows into its
        # per-step cache, _ShardedRows.forward gathers them into its output), so one buffer may carry word rows
        # and id rows alike.
        rows = P * cap
        return self.buf[: rows * row_bytes].view(dtype).view((rows,) + tuple(row_shape))


# ------------------------------------------------------------------------------------------ word table
class ShardedWordTable(nn.Module):
    """Drop-in for ``WordTable`` holding only rows ``rank, rank+P, ...`` of the frozen word table.
    ``many(idx_a, idx_b, ...)`` does ONE exchange for all documents of the step and returns ``Docs``
    handles over the per-step row cache."""

    def __init__(self, full_weight: torch.Tensor, transport: Transport):
        super().__init__()
        self.transport = transport
        P, rank = transport.world, transport.rank
        self.V, self.E = int(full_weight.shape[0]), int(full_weight.shape[1])
        self.P, self.cap = P, rows_local(self.V, P)
        self.weight = nn.Parameter(shard_rows(full_weight.detach().float(), rank, P).contiguous(), requires_grad=False)
        self.requires_grad = False                      # the inert attribute the reference sets (DeepCoNN.py:16)
        self._scr = None
        self._slots = {}
        from . import ops
        self._own_shadow = ops.ShadowTable()

    def _scratch(self, dev):
        if self._scr is None or self._scr[0].device != dev:
            flags = torch.zeros(self.V, device=dev, dtype=torch.int32)
            req = torch.zeros(self.P, 1 + self.cap, device=dev, dtype=torch.int64)
            rreq = torch.zeros_like(req)
            self._scr = (flags, req, rreq)
        return self._scr

    @staticmethod
    def _key(idx_list):
        from . import ops
        return tuple(((i.tokens.data_ptr(), i.offsets.data_ptr()) if isinstance(i, ops.RaggedIdx) else (i.data_ptr(), 0)) + tuple(i.shape)
                     for i in idx_list)

    def _new_cache(self, mode):
        """Row cache indexed by the ORIGINAL token id: [V, E] fp32 (exact mode) or [V + 1, Epad] half rows with the
        conv kernel's all-zero padding row at V.  Only the rows of the current step's tokens are valid."""
        dev = self.weight.device
        if mode == "exact":
            return torch.empty(self.V, self.E, device=dev, dtype=torch.float32)
        own = self._own_shadow.get(self.weight, mode)
        cache = torch.empty(self.V + 1, own.shape[1], device=dev, dtype=own.dtype)
        cache[self.V:].zero_()
        return cache

    def reserve(self, *idx_list):
        """A persistent row cache for the lookup of these index tensors -- the static input buffers of a captured
        step.  ``fill`` writes it, ``many`` on the same tensors returns it without exchanging: the lookup of step
        k+1 can then run on a side stream / graph branch while step k computes, which the frozen table permits
        (the lookup depends on the batch's token ids only)."""
        from . import ops
        key = self._key(idx_list)
        slot = self._slots.get(key)
        if slot is None:
            mode = ops.get_conv_mode()
            slot = self._slots[key] = {"mode": mode, "rows": self._new_cache(mode), "filled": False, "executed": False, "docs": None}
        return slot

    def fill(self, slot, *idx_list):
        """Run the exchange for ``idx_list`` into a reserved slot (current stream; graph-capturable)."""
        slot["docs"] = self._lookup(idx_list, slot)
        slot["filled"] = True
        if not torch.cuda.is_current_stream_capturing():
            slot["executed"] = True                     # a captured fill only runs when its graph is replayed

    def many(self, *idx_list):
        slot = self._slots.get(self._key(idx_list)) if self._slots else None
        if slot is not None and slot["filled"]:
            from . import ops
            if slot["mode"] != ops.get_conv_mode():
                raise RuntimeError("prefetched word rows were fetched for conv mode %r" % slot["mode"])
            return slot["docs"]
        return self._lookup(idx_list, None)

    def _lookup(self, idx_list, into):
        from . import ops
        from .pytorch_models.common_pytorch_models import Docs
        dev = self.weight.device
        flags, req, rreq = self._scratch(dev)
        # ragged documents are expanded first (the sharded lookup marks padded id tensors)
        idx_list = [(i.padded() if isinstance(i, ops.RaggedIdx) else i).contiguous() for i in idx_list]
        for idx in idx_list:
            K.mark(idx, self.V, flags)
        K.plan(flags, self.V, self.P, self.cap, req)
        self.transport.all_to_all(rreq, req)
        mode = ops.get_conv_mode()
        cache = into["rows"] if into else self._new_cache(mode)
        if mode == "exact":
            # strict-parity mode: fp32 rows, the conv / wgrad kernels read the cache as their word table
            rows = self.transport.exchange_rows(self.weight, rreq, self.P, self.cap, (self.E,), torch.float32)
            K.place(rows, req, self.P, self.cap, cache, self.V)
            table, shadow = cache, None
        else:
            # tensor-core modes: the owner serves rows of its half-precision shadow shard (built once: the
            # table is frozen), already in the conv kernel's layout -> half the NVLink bytes, no conversion
            own = self._own_shadow.get(self.weight, mode)
            rows = self.transport.exchange_rows(own, rreq, self.P, self.cap, (own.shape[1],), own.dtype)
            K.place(rows, req, self.P, self.cap, cache, self.V)
            table, shadow = None, ops.PrebuiltShadow(cache, self.V, self.E, mode)
        return tuple(Docs(idx, table, shadow) for idx in idx_list)

    def forward(self, idx):
        return self.many(idx)[0]

    def materialize(self, idx):
        from . import ops
        mode = ops.get_conv_mode()
        ops.set_conv_mode("exact")                      # fp32 rows
        try:
            d = self.forward(idx)
        finally:
            ops.set_conv_mode(mode)
        return ops.word_gather(d.table, d.idx)


# ------------------------------------------------------------------------------------------ id tables
class ShardSpec:
    """Attached to a parameter as ``param._r4r_shard``: its data is the local shard of an [R, ...] table."""

    def __init__(self, R: int, transport: Transport, grad_scale: Optional[float] = None, agree_cap: bool = False):
        self.R, self.transport = int(R), transport
        self.P, self.rank = transport.world, transport.rank
        self.grad_scale = (1.0 / self.P) if grad_scale is None else float(grad_scale)
        # The equal-split all-to-all needs the same per-peer capacity on every rank.  False: every rank
        # looks up the same number of ids per call (equal batch slices; no host sync, graph-capturable).
        # True: ranks may differ (ragged last batch); the capacity is agreed by a MAX all-reduce.
        self.agree_cap = bool(agree_cap)


class _ShardedRows(torch.autograd.Function):
    @staticmethod
    def forward(ctx, table, ids, spec):
        ids = ids.contiguous()
        n, P = ids.numel(), spec.P
        dev = table.device
        cap = spec.transport.agree_cap(max(n, 1), dev) if spec.agree_cap else max(n, 1)
        L = table.numel() // table.shape[0]
        req = torch.empty(P, 1 + cap, device=dev, dtype=torch.int64)
        pos = torch.empty(n, device=dev, dtype=torch.int64)
        K.bucket(ids.view(-1), spec.R, P, cap, req, pos)
        rreq = spec.transport.all_to_all(torch.empty_like(req), req)
        recv = spec.transport.exchange_rows(table, rreq, P, cap, (L,), torch.float32)
        out = torch.empty(n, L, device=dev, dtype=torch.float32)
        if n:
            K.gather(recv, pos, out)
        ctx.save_for_backward(pos, rreq)
        ctx.spec, ctx.cap, ctx.tshape, ctx.L = spec, cap, tuple(table.shape), L
        return out.reshape(*ids.shape) if table.dim() == 1 else out.reshape(*ids.shape, L)

    @staticmethod
    def backward(ctx, gout):
        pos, rreq = ctx.saved_tensors
        spec, cap, L, P = ctx.spec, ctx.cap, ctx.L, ctx.spec.P
        dev = gout.device
        send = torch.zeros(P * cap, L, device=dev, dtype=torch.float32)
        if pos.numel():
            K.scatter_unique(gout.contiguous().view(-1, L), pos, send)
        recv = spec.transport.all_to_all(torch.empty_like(send), send)
        gtable = torch.zeros(ctx.tshape, device=dev, dtype=torch.float32)
        K.scatter_owner(recv, rreq, P, cap, gtable, spec.grad_scale)
        return gtable, None, None


def sharded_rows_gather(table: torch.Tensor, ids: torch.Tensor) -> torch.Tensor:
    return _ShardedRows.apply(table, ids, table._r4r_shard)


# ------------------------------------------------------------------------------------------ model surgery
SHARDED_PARAMS = ("user_bias", "item_bias", "user_embedding.weight", "item_embedding.weight")


def _word_table_owners(model):
    from .pytorch_models.common_pytorch_models import WordTable
    for mod in model.modules():
        for name, child in list(mod.named_children()):
            if isinstance(child, WordTable):
                yield mod, name, child


def shard_model(model: nn.Module, transport: Transport, shard_word_table: bool = True,
                word_transport: Optional[Transport] = None, agree_cap: bool = False) -> nn.Module:
    """In place: the word table and the id tables / bias vectors of ``model`` (built with the FULL
    reference-shaped parameters, identical on every rank) keep only this rank's rows.
    ``word_transport`` (e.g. a ``P2PTransport``) carries the word rows; ids / id rows use ``transport``.
    ``agree_cap=True`` lets ranks look up different numbers of ids per call (see ``ShardSpec``)."""
    P, rank = transport.world, transport.rank
    if shard_word_table:
        for mod, name, child in list(_word_table_owners(model)):
            if child.weight.requires_grad:
                raise RuntimeError("the sharded word lookup is forward-only (frozen table, as in the reference); "
                                   "train_word_table needs the replicated table")
            setattr(mod, name, ShardedWordTable(child.weight.data, word_transport or transport).to(child.weight.device))
    params = dict(model.named_parameters())
    for key in SHARDED_PARAMS:
        p = params.get(key)
        if p is None:
            continue
        R = p.shape[0]
        p.data = shard_rows(p.data, rank, P).contiguous()
        p._r4r_shard = ShardSpec(R, transport, agree_cap=agree_cap)
    model._r4r_transport = transport
    return model


def dense_parameters(model: nn.Module) -> List[nn.Parameter]:
    """Replicated trainable parameters (their gradients are averaged across ranks)."""
    return [p for p in model.parameters() if p.requires_grad and not hasattr(p, "_r4r_shard")]


def allreduce_dense_grads(model: nn.Module, group=None, world: Optional[int] = None) -> None:
    """grad <- mean over ranks, for the replicated parameters that received a gradient; one flat
    all-reduce (the dense parameters are ~0.2 M floats)."""
    world = world if world is not None else dist.get_world_size(group)
    grads = [p.grad for p in dense_parameters(model) if p.grad is not None]
    if world == 1 or not grads:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, group=group)
    flat /= world
    ofs = 0
    for g in grads:
        g.copy_(flat[ofs:ofs + g.numel()].view_as(g))
        ofs += g.numel()


def gather_state_dict(model: nn.Module, group=None) -> dict:
    """Reference-layout ``state_dict`` (full tables, original keys/shapes) assembled on every rank:
    what ``main.py:123-126`` would ``torch.save``."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    sd = {k: v.detach() for k, v in model.state_dict().items()}
    params = dict(model.named_parameters())

    def gather(local, R):
        if world == 1:
            return unshard_rows([local], R)
        parts = [torch.empty_like(local) for _ in range(world)]
        dist.all_gather(parts, local.contiguous(), group=group)
        return unshard_rows(parts, R)

    out = {}
    for k, v in sd.items():
        p = params.get(k)
        if p is not None and hasattr(p, "_r4r_shard"):
            out[k] = gather(v, p._r4r_shard.R)
        else:
            out[k] = v.clone()
    for mod, name, child in [(m, n, c) for m in model.modules() for n, c in m.named_children() if isinstance(c, ShardedWordTable)]:
        prefix = [k for k, m in model.named_modules() if m is child][0]
        out[prefix + ".weight"] = gather(child.weight.detach(), child.V)
    return out
