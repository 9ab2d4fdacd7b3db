"""FusedAdam: torch.optim.Adam semantics as the reference configures it (main.py:94-96) -- dense
update of every parameter that has a gradient, L2 weight decay folded into the gradient, bias
correction -- executed by ONE multi-tensor kernel launch (r4r_adam_step) per 48 tensors instead
of torch's per-op foreach sweep.  Parameters whose ``.grad`` is None are skipped and their step
count does not advance, exactly like torch (SURVEY.md 7 "dense-Adam semantics").

``capturable=True`` keeps the step count in device memory (one int32 per parameter group, bumped by
r4r_counter_inc) so that ``step()`` can be recorded into a CUDA graph and replayed; the set of
parameters that receive gradients must then be the same on every replay (it is: in this path it
is a function of ``model_type`` only)."""
import ctypes

import torch

from ._lib import call
from .ops import _p, _stream


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, capturable=False):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, capturable=capturable))

    def _init_state(self, p):
        st = self.state[p]
        if not st:
            st["step"] = 0
            st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
            st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
        return st

    def _launch(self, group, ps, step, step_dev):
        n = len(ps)
        arr = ctypes.c_void_p * n
        P = arr(*[p.data_ptr() for p in ps])
        M = arr(*[self.state[p]["exp_avg"].data_ptr() for p in ps])
        V = arr(*[self.state[p]["exp_avg_sq"].data_ptr() for p in ps])
        NUM = (ctypes.c_int64 * n)(*[p.numel() for p in ps])
        grads = [p.grad.contiguous() for p in ps]          # keep alive across the launch
        G = arr(*[g.data_ptr() for g in grads])
        vp = lambda a: ctypes.cast(a, ctypes.c_void_p)
        b1, b2 = group["betas"]
        call("r4r_adam_step", n, vp(P), vp(G), vp(M), vp(V), vp(NUM), step, step_dev, group["lr"], b1, b2,
             group["eps"], group["weight_decay"], _stream())

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        for group in self.param_groups:
            live = []
            for p in group["params"]:
                if p.grad is None or p.numel() == 0:
                    continue
                if not p.is_cuda:
                    raise RuntimeError("FusedAdam runs on CUDA parameters only (no CPU fallback)")
                if p.dtype != torch.float32 or not p.is_contiguous():
                    raise RuntimeError("FusedAdam expects contiguous float32 parameters")
                self._init_state(p)
                live.append(p)
            if not live:
                continue
            if group["capturable"]:
                if "step_dev" not in group:
                    if torch.cuda.is_current_stream_capturing():
                        raise RuntimeError("FusedAdam(capturable=True): call prepare() before graph capture")
                    group["step_dev"] = torch.zeros(1, device=live[0].device, dtype=torch.int32)
                call("r4r_counter_inc", _p(group["step_dev"]), _stream())
                self._launch(group, live, 0, _p(group["step_dev"]))
                continue
            by_step = {}
            for p in live:
                st = self.state[p]
                st["step"] += 1
                by_step.setdefault(st["step"], []).append(p)
            for step, ps in by_step.items():
                self._launch(group, ps, step, ctypes.c_void_p(0))
        return loss

    def prepare(self):
        """Allocate the moment buffers and device step counters outside a graph capture."""
        for group in self.param_groups:
            for p in group["params"]:
                if p.requires_grad and p.numel() and p.is_cuda:
                    self._init_state(p)
                    if group["capturable"] and "step_dev" not in group:
                        group["step_dev"] = torch.zeros(1, device=p.device, dtype=torch.int32)
