"""Autograd-visible operators of the hot path; each one is a thin host wrapper over one or two
C-ABI calls (include/r4r_b200.h).  Tensors only provide device memory and the current stream.

There is deliberately no CPU implementation: a non-CUDA tensor raises RuntimeError.
"""
import ctypes
import os
from typing import Optional

import torch

from . import _lib
from ._lib import call

NUM_FILTERS = 100          # common_pytorch_models.py:11
_MODES = ("exact", "f16", "bf16", "f16r", "bf16r")
_conv_mode = os.environ.get("R4R_CONV_MODE", "f16")
_doc_plan = os.environ.get("R4R_DOC_PLAN", "1") != "0"      # skip the repeated-padding tail of documents (exact)
_PAIR_TILE = 256                                             # conv positions per CTA-pair tile of conv_pool_tc


def set_doc_plan(on: bool) -> None:
    global _doc_plan
    _doc_plan = bool(on)


def get_doc_plan() -> bool:
    return _doc_plan


# How the fused conv / wgrad kernels treat ``RaggedIdx`` documents in the tensor-core modes:
# False (default) = rebuild the padded int64 ids first (r4r_docs_expand, two ~10 us kernels that become part
# of a captured step); True = read the ragged tokens directly (r4r_*_ragged entry points: exact same results,
# no padded tensor at all, but the conv's token prefetch is currently ~8 % slower on that path).
_ragged_native = os.environ.get("R4R_RAGGED_NATIVE", "0") == "1"


def set_ragged_native(on: bool) -> None:
    global _ragged_native
    _ragged_native = bool(on)


def doc_lengths(idx: "torch.Tensor") -> "torch.Tensor":
    """Informative prefix length of every document of ``idx`` [N,T] (r4r_doc_plan): rows after it repeat
    one token and cannot change the max-pooled features."""
    if isinstance(idx, RaggedIdx):
        rg = idx.reshape(-1, idx.shape[-1])
        N, T, dev = int(rg.shape[0]), int(rg.shape[1]), rg.device
    else:
        _need_cuda(idx)
        idx, rg = _i64c(idx), None
        (N, T), dev = idx.shape, idx.device
    doc_len = torch.empty(N, device=dev, dtype=torch.int32)
    order = torch.empty(N, device=dev, dtype=torch.int32)
    ws = torch.empty(_lib.lib.r4r_doc_plan_ws_bytes(N, T), device=dev, dtype=torch.uint8)
    if N and rg is None:
        call("r4r_doc_plan", _p(idx), N, T, _p(doc_len), _p(order), _p(ws), _stream())
    elif N:
        call("r4r_doc_plan_ragged", _p(rg.offsets), N, T, _p(doc_len), _p(order), _p(ws), _stream())
    return doc_len


def set_conv_mode(mode: str) -> None:
    """'exact' = fp32 CUDA-core kernel (strict parity); 'f16' / 'bf16' = tcgen05 tensor-core kernel
    reading a private half-precision shadow of the frozen word table (fp32 accumulation); 'f16r' / 'bf16r' = the
    tensor-core kernel only SELECTS each filter's arg-max window, whose value is then re-evaluated in fp32 from the
    fp32 table and filters (r4r_conv_refine) and whose gradient is the fp32 one: fp32-grade results at tensor-core speed."""
    global _conv_mode
    if mode not in _MODES:
        raise ValueError("conv mode must be one of %s" % (_MODES,))
    _conv_mode = mode


def get_conv_mode() -> str:
    return _conv_mode


# Optional measurement hook (bench.py): when a list is installed here, every launch of the dominant
# kernel (r4r_conv_pool_tc / r4r_conv_pool_simt) is bracketed by a pair of timing CUDA events on the
# launching stream; under graph capture they become external event-record nodes, re-recorded at
# every replay.
_conv_event_sink = None


def set_conv_event_sink(sink):
    global _conv_event_sink
    _conv_event_sink = sink


class _ConvTimer:
    def __enter__(self):
        self.on = _conv_event_sink is not None
        if self.on:
            ext = torch.cuda.is_current_stream_capturing()
            self.e0 = torch.cuda.Event(enable_timing=True, external=ext)
            self.e1 = torch.cuda.Event(enable_timing=True, external=ext)
            self.e0.record()

    def __exit__(self, *exc):
        if self.on:
            self.e1.record()
            _conv_event_sink.append((self.e0, self.e1))


# ------------------------------------------------------------------------------------ zero arena
# The parameter-gradient kernels accumulate with atomics, so every backward needs zeroed dW / db / dV buffers:
# ~14 tiny fill kernels per DeepCoNN step.  Inside a captured step they are carved out of ONE flat buffer that
# a single memset clears at the start of the step (train.CapturedStep calls arena_begin / arena_end).
class _ZeroArena:
    buf: Optional[torch.Tensor] = None
    ofs = 0
    active = False
    FLOATS = 1 << 20


def arena_begin(device, floats: int = _ZeroArena.FLOATS) -> None:
    a = _ZeroArena
    if a.buf is None or a.buf.device != torch.device(device) or a.buf.numel() < floats:
        a.buf = torch.empty(floats, device=device, dtype=torch.float32)
    a.buf.zero_()
    a.ofs, a.active = 0, True


def arena_end() -> None:
    _ZeroArena.active = False


def zeros_f32(shape, device) -> torch.Tensor:
    """Zero-filled fp32 tensor: a 16-byte aligned slice of the step's arena when one is active."""
    a = _ZeroArena
    n = 1
    for d in shape:
        n *= int(d)
    if a.active and a.buf.device == torch.device(device) and a.ofs + n <= a.buf.numel():
        out = a.buf[a.ofs:a.ofs + n].view(tuple(shape))
        a.ofs += (n + 3) & ~3
        return out
    return torch.zeros(tuple(shape), device=device, dtype=torch.float32)


def _p(t: Optional[torch.Tensor]):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("reviews4rec_b200 kernels run on CUDA tensors only (got a %s tensor); "
                               "there is no CPU fallback" % t.device)


def _f32c(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.float32:
        raise TypeError("expected float32, got %s" % t.dtype)
    return t.contiguous()


def _i64c(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.int64:
        raise TypeError("ids must be int64 (LongTensor) as in the reference, got %s" % t.dtype)
    return t.contiguous()


# ------------------------------------------------------------------------------------ gather
def word_gather(table: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """nn.Embedding forward on the frozen word table (DeepCoNN.py:53-54).  Bit-exact row copies."""
    if isinstance(idx, RaggedIdx):
        idx = idx.padded()
    _need_cuda(table, idx)
    table, idx = _f32c(table), _i64c(idx)
    out = torch.empty(*idx.shape, table.shape[1], device=table.device, dtype=torch.float32)
    if idx.numel() == 0:
        return out
    call("r4r_word_gather_f32", _p(table), table.shape[0], table.shape[1], _p(idx), idx.numel(), _p(out), _stream())
    return out


class ShadowTable:
    """fp16/bf16 padded copy of a frozen fp32 word table, rebuilt only when the table changes
    (xavier_init / load_state_dict bump ``_version``)."""

    def __init__(self):
        self._key = None
        self.tensor = None
        self.epad = 0

    def get(self, table: torch.Tensor, mode: str) -> torch.Tensor:
        key = (table.data_ptr(), table._version, tuple(table.shape), mode)
        if key != self._key or table.requires_grad:     # a trainable table (opt-in) changes under the optimizer's raw kernels
            V, E = table.shape
            self.epad = ((E + 63) // 64) * 64                     # 128-byte aligned rows
            dt = torch.float16 if mode == "f16" else torch.bfloat16
            self.tensor = torch.empty(V + 1, self.epad, device=table.device, dtype=dt)   # row V: zeros (conv padding)
            call("r4r_shadow_build", _p(table), V, E, _p(self.tensor), self.epad,
                 _lib.R4R_DT_F16 if mode == "f16" else _lib.R4R_DT_BF16, _stream())
            self._key = key
        return self.tensor


class RaggedIdx:
    """Token ids of ``[N, T]`` (or ``[N, R, W]``) padded documents kept ragged on the device: row r is
    ``tokens[offsets[r]:offsets[r+1]]`` followed by ``pad_id`` up to the last dim.  It stands in for the
    reference's padded ``LongTensor`` (data_fast.py:102-108) in ``model(data)``: the fused conv / wgrad
    kernels read it directly, everything else goes through ``padded()``."""

    def __init__(self, tokens: torch.Tensor, offsets: torch.Tensor, shape, pad_id: int = 0):
        if tokens.dtype != torch.int32 or offsets.dtype != torch.int64:
            raise TypeError("RaggedIdx wants int32 tokens and int64 offsets")
        self.tokens, self.offsets, self.pad_id = tokens, offsets, int(pad_id)
        self.shape = torch.Size(shape)
        rows = 1
        for d in self.shape[:-1]:
            rows *= int(d)
        if offsets.numel() != rows + 1:
            raise ValueError("offsets must have %d entries for shape %s" % (rows + 1, tuple(self.shape)))

    device = property(lambda self: self.tokens.device)
    is_cuda = property(lambda self: self.tokens.is_cuda)
    dtype = torch.int64

    def dim(self):
        return len(self.shape)

    def numel(self):
        return int(self.shape.numel())

    def contiguous(self):
        return self

    def reshape(self, *shape):
        shape = list(shape[0]) if len(shape) == 1 and not isinstance(shape[0], int) else list(shape)
        total = self.numel()
        if -1 in shape:
            known = 1
            for d in shape:
                known *= d if d != -1 else 1
            shape[shape.index(-1)] = total // max(known, 1)
        if int(torch.Size(shape).numel()) != total or shape[-1] != self.shape[-1]:
            raise RuntimeError("RaggedIdx.reshape must keep the document length (last dim): %s -> %s" % (tuple(self.shape), tuple(shape)))
        return RaggedIdx(self.tokens, self.offsets, shape, self.pad_id)

    view = reshape

    def padded(self) -> torch.Tensor:
        """The int64 tensor the reference's reader would have produced (r4r_docs_expand, exact)."""
        _need_cuda(self.tokens, self.offsets)
        rows, T = self.offsets.numel() - 1, int(self.shape[-1])
        out = torch.empty(tuple(self.shape), device=self.tokens.device, dtype=torch.int64)
        call("r4r_docs_expand", _p(self.tokens), _p(self.offsets), rows, T, self.pad_id, _p(out), _stream())
        return out


class PrebuiltShadow:
    """Half-precision word rows that already are in the conv kernel's layout ([V+1, Epad], row V zero):
    the per-step row cache a sharded word table receives from the owners (sharded.py).  There is no
    fp32 table behind it; ``conv_pool`` is then called with ``table=None``."""

    def __init__(self, tensor: torch.Tensor, V: int, E: int, mode: str):
        self.tensor, self.V, self.E, self.mode = tensor, int(V), int(E), mode
        self.epad = int(tensor.shape[1])

    def get(self, table, mode):
        if mode != self.mode:
            raise RuntimeError("word rows were fetched for conv mode %r, not %r" % (self.mode, mode))
        return self.tensor


# ------------------------------------------------------------------------------------ conv + pool
def conv_pool_forward(idx: torch.Tensor, table: torch.Tensor, conv_w: torch.Tensor, conv_b: torch.Tensor,
                      mode: str, shadow: Optional[ShadowTable] = None, want_shadow: bool = False):
    """Fused gather -> conv(3xE, pad 2) -> relu -> global max-pool.  Returns (pooled [N,F], argmax [N,F])
    (+ the half-precision shadow rows it read, (tensor, Epad) or None, when ``want_shadow``)."""
    ragged = None
    if isinstance(idx, RaggedIdx):
        idx = idx.reshape(-1, idx.shape[-1])
        if mode == "exact" or not _ragged_native:
            idx = idx.padded()                                   # the fp32 kernels (and the default policy) read padded ids
        else:
            ragged, idx = idx, idx.tokens
    _need_cuda(idx, table, conv_w, conv_b)
    conv_w, conv_b = _f32c(conv_w), _f32c(conv_b)
    if ragged is None:
        idx = _i64c(idx)
        N, T = idx.shape
    else:
        _need_cuda(ragged.offsets)
        N, T = int(ragged.shape[0]), int(ragged.shape[1])
    if table is None:
        if not isinstance(shadow, PrebuiltShadow) or mode == "exact":
            raise RuntimeError("conv_pool needs the fp32 word table (or prebuilt half-precision rows in f16/bf16 mode)")
        V, E = shadow.V, shadow.E
    else:
        table = _f32c(table)
        V, E = table.shape
    dev = idx.device
    F = conv_w.shape[0]
    if tuple(conv_w.shape) != (F, 1, 3, E):
        raise RuntimeError("conv weight must be [F,1,3,%d] (window size 3), got %s" % (E, tuple(conv_w.shape)))
    pooled = torch.empty(N, F, device=dev, dtype=torch.float32)
    argmax = torch.empty(N, F, device=dev, dtype=torch.int32)
    used = None
    if N == 0:                                                   # empty batch: nothing to launch
        if mode in ("f16", "bf16"):
            shadow = shadow if shadow is not None else ShadowTable()
            used = (shadow.get(table, mode), shadow.epad, V)
        return (pooled, argmax, used) if want_shadow else (pooled, argmax)
    if mode == "exact":
        keys = torch.empty(N, F, device=dev, dtype=torch.int64)
        with _ConvTimer():
            call("r4r_conv_pool_simt", _p(table), V, E, _p(idx), N, T, _p(conv_w), _p(conv_b), F,
                 _p(pooled), _p(argmax), _p(keys), _stream())
    else:
        refine = mode.endswith("r")
        mode = mode[:-1] if refine else mode
        if refine and (table is None or ragged is not None):
            raise RuntimeError("the fp32-refined conv modes need the fp32 word table and padded ids")
        shadow = shadow if shadow is not None else ShadowTable()
        sh = shadow.get(table, mode)
        dt = _lib.R4R_DT_F16 if mode == "f16" else _lib.R4R_DT_BF16
        nbytes = _lib.lib.r4r_conv_wpack_bytes(E, F)
        if nbytes <= 0:
            raise RuntimeError("conv_pool_tc: unsupported shape E=%d F=%d" % (E, F))
        wpack = torch.empty(nbytes, device=dev, dtype=torch.uint8)
        call("r4r_conv_pack_weights", _p(conv_w), E, F, _p(wpack), dt, _stream())
        doc_len = doc_order = None
        if _doc_plan and N > 0:
            # documents padded with a repeated token are cut to their informative prefix (exact, see
            # r4r_doc_plan in include/r4r_b200.h) and issued longest first
            doc_len = torch.empty(N, device=dev, dtype=torch.int32)
            doc_order = torch.empty(N, device=dev, dtype=torch.int32)
            ws = torch.empty(_lib.lib.r4r_doc_plan_ws_bytes(N, T), device=dev, dtype=torch.uint8)
            if ragged is None:
                call("r4r_doc_plan", _p(idx), N, T, _p(doc_len), _p(doc_order), _p(ws), _stream())
            else:
                call("r4r_doc_plan_ragged", _p(ragged.offsets), N, T, _p(doc_len), _p(doc_order), _p(ws), _stream())
        # scratch for the launch's window streams (documents of each CTA pair laid end to end, see include/r4r_b200.h)
        sws = torch.empty(_lib.lib.r4r_conv_stream_ws_bytes(N, T), device=dev, dtype=torch.uint8)
        with _ConvTimer():
            if ragged is None:
                call("r4r_conv_pool_tc", _p(sh), V, shadow.epad, E, dt, _p(idx), N, T, _p(wpack), _p(conv_b), F,
                     _p(pooled), _p(argmax), _p(doc_len), _p(doc_order), _p(sws), _stream())
            else:
                call("r4r_conv_pool_tc_ragged", _p(sh), V, shadow.epad, E, dt, _p(ragged.tokens), _p(ragged.offsets),
                     ragged.pad_id, N, T, _p(wpack), _p(conv_b), F, _p(pooled), _p(argmax), _p(doc_len), _p(doc_order), _p(sws), _stream())
        used = (sh, shadow.epad, V)
        if refine:
            # fp32 value of the selected window; the backward then is the fp32 gradient (used = None)
            call("r4r_conv_refine", _p(table), V, E, _p(idx), N, T, _p(argmax), _p(conv_w), _p(conv_b), F, _p(pooled), _stream())
            used = None
    return (pooled, argmax, used) if want_shadow else (pooled, argmax)


class _ConvPool(torch.autograd.Function):
    @staticmethod
    def forward(ctx, idx, table, conv_w, conv_b, mode, shadow):
        # the reference freezes the word table (DeepCoNN.py:15): table.requires_grad is the opt-in extension of
        # SURVEY.md 8f-3 (hyper_params['train_word_table']), served by r4r_conv_dgrad_scatter in backward
        ctx.set_materialize_grads(False)                # no zero-filled gradient for the (integer) arg-max output
        ctx.table_grad = table is not None and table.requires_grad
        ctx.conv_w = conv_w.detach() if ctx.table_grad else None
        if isinstance(idx, RaggedIdx):
            idx = idx.reshape(-1, idx.shape[-1])
            if mode == "exact" or mode.endswith("r") or not _ragged_native:
                idx = idx.padded()
        pooled, argmax, used = conv_pool_forward(idx, table, conv_w, conv_b, mode, shadow, want_shadow=True)
        ctx.ragged = idx if isinstance(idx, RaggedIdx) else None
        ctx.save_for_backward(None if ctx.ragged is not None else idx, table, argmax, pooled)
        ctx.wshape = tuple(conv_w.shape)
        ctx.mode, ctx.used = mode, used
        ctx.mark_non_differentiable(argmax)
        return pooled, argmax

    @staticmethod
    def backward(ctx, gpooled, _gargmax):
        idx, table, argmax, pooled = ctx.saved_tensors
        if gpooled is None:
            gpooled = torch.zeros_like(pooled)
        F, _, _, E = ctx.wshape
        rg = ctx.ragged
        N, T = (int(rg.shape[0]), int(rg.shape[1])) if rg is not None else idx.shape
        dW = zeros_f32(ctx.wshape, pooled.device)
        db = zeros_f32((F,), pooled.device)
        if N == 0:                                               # empty batch: zero gradients, nothing to launch
            return None, (torch.zeros_like(table) if ctx.table_grad else None), dW, db, None, None
        if ctx.used is None:
            call("r4r_conv_wgrad_argmax", _p(table), table.shape[0], E, _p(idx), N, T, _p(argmax), _p(pooled),
                 _p(_f32c(gpooled)), F, _p(dW), _p(db), _stream())
        else:
            # f16 / bf16 modes: gradient of what the tensor-core forward computed, from the same shadow rows
            sh, epad, V = ctx.used
            dt = _lib.R4R_DT_F16 if ctx.mode == "f16" else _lib.R4R_DT_BF16
            if rg is None:
                call("r4r_conv_wgrad_argmax_h", _p(sh), V, epad, E, dt, _p(idx), N, T, _p(argmax), _p(pooled),
                     _p(_f32c(gpooled)), F, _p(dW), _p(db), _stream())
            else:
                call("r4r_conv_wgrad_argmax_h_ragged", _p(sh), V, epad, E, dt, _p(rg.tokens), _p(rg.offsets), rg.pad_id, N, T,
                     _p(argmax), _p(pooled), _p(_f32c(gpooled)), F, _p(dW), _p(db), _stream())
        gtable = None
        if ctx.table_grad:
            pidx = rg.padded() if rg is not None else idx
            gtable = torch.zeros_like(table)
            call("r4r_conv_dgrad_scatter", _p(pidx), N, T, _p(argmax), _p(pooled), _p(_f32c(gpooled)), _p(_f32c(ctx.conv_w)), F, E,
                 _p(gtable), table.shape[0], _stream())
        return None, gtable, dW, db, None, None


def conv_pool(idx, table, conv_w, conv_b, mode=None, shadow=None):
    pooled, _ = _ConvPool.apply(idx, table, conv_w, conv_b, mode or _conv_mode, shadow)
    return pooled


# ------------------------------------------------------------------------------------ linear
class _Linear(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, W, b):
        _need_cuda(x, W, b)
        x, W = _f32c(x), _f32c(W)
        b = _f32c(b) if b is not None else None
        n, in_f = x.shape
        out_f = W.shape[0]
        y = torch.empty(n, out_f, device=x.device, dtype=torch.float32)
        if n:
            call("r4r_linear_fwd", _p(x), _p(W), _p(b), n, in_f, out_f, _p(y), _stream())
        ctx.save_for_backward(x, W)
        ctx.has_bias = b is not None
        return y

    @staticmethod
    def backward(ctx, gy):
        x, W = ctx.saved_tensors
        gy = _f32c(gy)
        n, in_f = x.shape
        out_f = W.shape[0]
        dx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        dW = zeros_f32(W.shape, x.device) if ctx.needs_input_grad[1] else None
        db = zeros_f32((out_f,), x.device) if (ctx.has_bias and ctx.needs_input_grad[2]) else None
        if n:
            call("r4r_linear_bwd", _p(x), _p(W), _p(gy), n, in_f, out_f, _p(dx), _p(dW), _p(db), _stream())
        return dx, dW, db


def linear(x: torch.Tensor, W: torch.Tensor, b: Optional[torch.Tensor]) -> torch.Tensor:
    """nn.Linear over the last dim; leading dims are flattened for the kernel."""
    lead = x.shape[:-1]
    y = _Linear.apply(x.reshape(-1, x.shape[-1]), W, b)
    return y.reshape(*lead, W.shape[0])


# ------------------------------------------------------------------------------------ FM
class _FM(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, V, lin_w, lin_b):
        _need_cuda(x, V, lin_w, lin_b)
        x, V, lin_w, lin_b = _f32c(x), _f32c(V), _f32c(lin_w), _f32c(lin_b)
        n, nf = x.shape
        k = V.shape[1]
        out = torch.empty(n, device=x.device, dtype=torch.float32)
        if n:
            call("r4r_fm_fwd", _p(x), _p(V), _p(lin_w), _p(lin_b), n, nf, k, _p(out), _stream())
        ctx.save_for_backward(x, V, lin_w)
        return out

    @staticmethod
    def backward(ctx, gout):
        x, V, lin_w = ctx.saved_tensors
        n, nf = x.shape
        k = V.shape[1]
        dx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        dV = zeros_f32(V.shape, x.device)
        dw = zeros_f32(lin_w.shape, x.device)
        db = zeros_f32((1,), x.device)
        if n:
            call("r4r_fm_bwd", _p(x), _p(V), _p(lin_w), _p(_f32c(gout)), n, nf, k, _p(dx), _p(dV), _p(dw), _p(db), _stream())
        return dx, dV, dw, db


def fm(x, V, lin_w, lin_b) -> torch.Tensor:
    """TorchFM.forward (common_pytorch_models.py:49-57); returns [n,1] like the reference."""
    return _FM.apply(x, V, lin_w, lin_b).unsqueeze(1)


# ------------------------------------------------------------------------------------ fused DeepCoNN head (K3)
HEAD_MAX_L, HEAD_MAX_F, HEAD_MAX_K = 32, 128, 16


def _vp_array(tensors):
    arr = (ctypes.c_void_p * len(tensors))(*[(t.data_ptr() if t is not None else 0) for t in tensors])
    return arr, ctypes.cast(arr, ctypes.c_void_p)


class _DeepConnHead(torch.autograd.Function):
    """fc_u / fc_i (+ dropout) -> cat -> FM + global bias (head 0) or final MLP + biases (head 1) -> rating [-> squared
    error], one kernel forward, one backward (r4r_deepconn_head_fwd / _bwd)."""

    @staticmethod
    def forward(ctx, pu, pi, fuw, fub, fiw, fib, fmV, fmw, fmb, w0, b0, w3, b3, ub, ib, gb, y, head, p, seed, step, masks, se_sum):
        _need_cuda(pu, pi, fuw, fiw, gb)
        ctx.set_materialize_grads(False)                # an unused output (rating when the loss is fused, or se) costs no zero fill
        f = lambda t: None if t is None else _f32c(t)
        pu, pi, fuw, fub, fiw, fib, gb = f(pu), f(pi), f(fuw), f(fub), f(fiw), f(fib), f(gb)
        fmV, fmw, fmb, w0, b0, w3, b3, ub, ib, y = f(fmV), f(fmw), f(fmb), f(w0), f(b0), f(w3), f(b3), f(ub), f(ib), f(y)
        N, F = pu.shape
        L = fuw.shape[0]
        K = fmV.shape[1] if head == 0 else 0
        dev = pu.device
        rating = torch.empty(N, device=dev, dtype=torch.float32)
        se = torch.empty(N, device=dev, dtype=torch.float32) if y is not None else None
        cat = torch.empty(N, 2 * L, device=dev, dtype=torch.float32)
        hid = torch.empty(N, L, device=dev, dtype=torch.float32) if head == 1 else None
        keep = torch.empty(N, 3, device=dev, dtype=torch.int32)
        w3f = None if w3 is None else w3.reshape(-1)
        ptrs = [pu, pi, fuw, fiw, fub, fib, fmV, None if fmw is None else fmw.reshape(-1), fmb, w0, b0, w3f, b3, ub, ib, gb, y,
                masks, step, rating, se, se_sum if y is not None else None, cat, hid, keep]
        keepalive, vp = _vp_array(ptrs)
        call("r4r_deepconn_head_fwd", vp, N, F, L, K, head, float(p), int(seed) & 0xFFFFFFFFFFFFFFFF, _stream())
        ctx.ptrs, ctx.dims = ptrs, (N, F, L, K, head, float(p))
        ctx.shapes = [None if t is None else tuple(t.shape) for t in (fuw, fub, fiw, fib, fmV, fmw, fmb, w0, b0, w3, b3)]
        return (rating, se) if y is not None else (rating, rating.new_zeros(0))

    @staticmethod
    def backward(ctx, g_rating, g_se):
        N, F, L, K, head, p = ctx.dims
        ptrs = ctx.ptrs
        dev = ptrs[0].device
        has_se = ptrs[16] is not None and g_se is not None and g_se.numel() == N
        if g_rating is None and not has_se:
            g_rating = torch.zeros(N, device=dev, dtype=torch.float32)
        g_rating = None if g_rating is None else _f32c(g_rating)
        g_se = _f32c(g_se) if has_se else None
        sh = ctx.shapes
        z = lambda shape: None if shape is None else zeros_f32(shape, dev)
        dpu, dpi = torch.empty(N, F, device=dev), torch.empty(N, F, device=dev)
        dfuw, dfub, dfiw, dfib = z(sh[0]), z(sh[1]), z(sh[2]), z(sh[3])
        dV, dfmw, dfmb, dw0, db0, dw3, db3 = z(sh[4]), z(sh[5]), z(sh[6]), z(sh[7]), z(sh[8]), z(sh[9]), z(sh[10])
        dub = torch.empty(N, device=dev) if head == 1 else None
        dib = torch.empty(N, device=dev) if head == 1 else None
        dg = zeros_f32((1,), dev)
        if N:
            gp = [g_rating, g_se, dpu, dpi, dfuw, dfiw, dfub, dfib, dV, dfmw, dfmb, dw0, db0, dw3, db3, dub, dib, dg]
            ka1, vp1 = _vp_array(ptrs)
            ka2, vp2 = _vp_array(gp)
            call("r4r_deepconn_head_bwd", vp1, vp2, N, F, L, K, head, p, _stream())
        return (dpu, dpi, dfuw, dfub, dfiw, dfib, dV, dfmw, dfmb, dw0, db0, dw3, db3, dub, dib, dg, None,
                None, None, None, None, None, None)


def deepconn_head(pooled_u, pooled_i, fc_u, fc_i, head, fm=None, final=None, ub=None, ib=None, global_bias=None, y=None,
                  p=0.0, seed=0, step=None, masks=None, se_sum=None):
    """Fused DeepCoNN head.  ``fc_u`` / ``fc_i`` = (weight [L,F], bias [L]); ``fm`` = (V [2L,K], lin.weight [1,2L], lin.bias [1])
    for head 0; ``final`` = (final.0.weight [L,2L], final.0.bias, final.3.weight [1,L], final.3.bias) and the gathered bias
    values ``ub`` / ``ib`` [N] for head 1.  Returns (rating [N], se [N] or None).  ``p`` = dropout probability of this call
    (0 in eval mode), ``masks`` = uint8 keep masks [N, 3L] (tests), ``step`` = int32 device counter of the Philox stream
    (advanced by the backward kernel), ``se_sum`` = device scalar the batch's squared-error sum is added to."""
    fmV, fmw, fmb = fm if fm is not None else (None, None, None)
    w0, b0, w3, b3 = final if final is not None else (None, None, None, None)
    rating, se = _DeepConnHead.apply(pooled_u, pooled_i, fc_u[0], fc_u[1], fc_i[0], fc_i[1], fmV, fmw, fmb, w0, b0, w3, b3, ub, ib,
                                     global_bias, y, int(head), float(p), int(seed), step, masks, se_sum)
    return rating, (se if y is not None else None)


def deepconn_head_supported(L, F, K):
    return 0 < L <= HEAD_MAX_L and 0 < F <= HEAD_MAX_F and 0 <= K <= HEAD_MAX_K


# ------------------------------------------------------------------------------------ MSE
class _MSE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, out, y):
        _need_cuda(out, y)
        shape = out.shape
        o, t = _f32c(out).reshape(-1), _f32c(y.expand_as(out)).reshape(-1)
        se = torch.empty_like(o)
        if o.numel():
            call("r4r_mse_fwd", _p(o), _p(t), o.numel(), _p(se), _p(None), _stream())
        ctx.save_for_backward(o, t)
        ctx.shape = shape
        return se.reshape(shape)

    @staticmethod
    def backward(ctx, gse):
        o, t = ctx.saved_tensors
        g = torch.empty_like(o)
        if o.numel():
            call("r4r_mse_bwd", _p(o), _p(t), _p(_f32c(gse).reshape(-1)), o.numel(), _p(g), _stream())
        return g.reshape(ctx.shape), None


def squared_error(out, y):
    return _MSE.apply(out, y)


# ------------------------------------------------------------------------------------ id rows
class _RowsGather(torch.autograd.Function):
    @staticmethod
    def forward(ctx, table, ids):
        _need_cuda(table, ids)
        table, ids = _f32c(table), _i64c(ids)
        L = 1 if table.dim() == 1 else table.shape[1]
        n = ids.numel()
        out = torch.empty((n,) if table.dim() == 1 else (n, L), device=table.device, dtype=torch.float32)
        if n:
            call("r4r_rows_gather", _p(table), table.shape[0], L, _p(ids), n, _p(out), _stream())
        ctx.save_for_backward(ids)
        ctx.tshape = tuple(table.shape)
        return out.reshape(*ids.shape) if table.dim() == 1 else out.reshape(*ids.shape, L)

    @staticmethod
    def backward(ctx, gout):
        (ids,) = ctx.saved_tensors
        L = 1 if len(ctx.tshape) == 1 else ctx.tshape[1]
        # dense gradient, as nn.Embedding(sparse=False) / Tensor.gather produce (SURVEY.md finding 5)
        gtable = torch.zeros(ctx.tshape, device=gout.device, dtype=torch.float32)
        if ids.numel():
            call("r4r_rows_scatter_add", _p(_f32c(gout)), _p(ids), ids.numel(), L, _p(gtable), ctx.tshape[0], _stream())
        return gtable, None


def rows_gather(table, ids):
    """table[ids] for an id-embedding matrix [R,L] or a bias vector [R].  A parameter that
    ``sharded.shard_model`` reduced to this rank's rows is looked up across the ranks."""
    if hasattr(table, "_r4r_shard"):
        from .sharded import sharded_rows_gather
        _need_cuda(table, ids)
        return sharded_rows_gather(table, _i64c(ids))
    return _RowsGather.apply(table, ids)
