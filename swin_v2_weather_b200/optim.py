"""Adam for the B200 path: a drop-in for the optimizer the reference constructs,

    torch.optim.Adam(model.parameters(), lr=params.lr, betas=(0.9, 0.95), fused=True)          (train.py:175-176)

with the same constructor arguments, `state_dict()` layout ('step', 'exp_avg', 'exp_avg_sq' per parameter) and
GradScaler protocol (`_step_supports_amp_scaling`: scale / found_inf are consumed on the device, no host sync).
`step()` is one multi-tensor C-ABI call (`swinb200_adam_step`) that updates fp32 masters and moments and, in the same
pass, rewrites the bf16 shadows the tensor-core GEMMs read -- so no per-step re-cast of 136 M weights follows it.
There is no CPU fallback: CPU parameters raise.

One deliberate difference: when a GradScaler reports an overflow the kernel skips the update on the device, but the
host-side 'step' counter has already advanced (torch's fused Adam rolls it back on the device).  In bf16 -- the mode this
package computes in -- gradients do not overflow and a GradScaler never fires.
"""
from __future__ import annotations

import ctypes
from typing import Iterable, Optional, Tuple

import torch

from . import _lib
from ._lib import SwinB200Error


class Adam(torch.optim.Optimizer):
    def __init__(self, params: Iterable, lr: float = 1e-3, betas: Tuple[float, float] = (0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 0.0, amsgrad: bool = False, *, maximize: bool = False, fused: Optional[bool] = None,
                 foreach: Optional[bool] = None, capturable: bool = False, differentiable: bool = False,
                 decoupled_weight_decay: bool = False):
        if amsgrad or maximize or capturable or differentiable or decoupled_weight_decay:
            raise NotImplementedError("swin_v2_weather_b200.optim.Adam implements the configuration the reference uses "
                                      "(amsgrad / maximize / capturable / differentiable / decoupled_weight_decay = False)")
        if lr < 0 or eps < 0 or not 0 <= betas[0] < 1 or not 0 <= betas[1] < 1 or weight_decay < 0:
            raise ValueError(f"invalid Adam hyper-parameters: lr={lr} betas={betas} eps={eps} weight_decay={weight_decay}")
        # `fused` / `foreach` are accepted for signature compatibility; the step is always the fused kernel.  The stored
        # defaults say fused=False so that torch keeps 'step' as a host scalar tensor when a state_dict is loaded.
        defaults = dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay, amsgrad=False, maximize=False,
                        foreach=None, capturable=False, differentiable=False, fused=False, decoupled_weight_decay=False)
        super().__init__(params, defaults)
        self._step_supports_amp_scaling = True      # torch.amp.GradScaler hands over grad_scale / found_inf instead of unscaling

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        grad_scale = getattr(self, "grad_scale", None)
        found_inf = getattr(self, "found_inf", None)
        from .functional import SHADOWS
        for group in self.param_groups:
            ps, gs, ms, vs, shs, ns, refreshed = [], [], [], [], [], [], []
            step_no = None
            for p in group["params"]:
                if p.grad is None:
                    continue
                if not p.is_cuda or p.dtype != torch.float32:
                    raise SwinB200Error("optim.Adam: parameters must be fp32 CUDA tensors (no CPU fallback)")
                if p.grad.is_sparse:
                    raise SwinB200Error("optim.Adam does not support sparse gradients")
                if not p.is_contiguous():
                    raise SwinB200Error("optim.Adam: parameters must be contiguous")
                g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
                if g.dtype != torch.float32:
                    raise SwinB200Error("optim.Adam: gradients must be fp32")
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = torch.tensor(0.0, dtype=torch.float32)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                if st["step"].is_cuda:      # state loaded from torch's fused Adam keeps 'step' on the device: one sync, once
                    st["step"] = st["step"].detach().to("cpu", torch.float32)
                st["step"] += 1
                s = int(st["step"])
                if step_no is None:
                    step_no = s
                elif s != step_no:       # parameters that joined later: flush what we have and start a new table
                    self._launch(group, step_no, ps, gs, ms, vs, shs, ns, grad_scale, found_inf)
                    self._after(refreshed, found_inf)
                    ps, gs, ms, vs, shs, ns, refreshed = [], [], [], [], [], [], []
                    step_no = s
                sh = SHADOWS.peek(p)
                ps.append(p.data_ptr()); gs.append(g.data_ptr()); ms.append(st["exp_avg"].data_ptr())
                vs.append(st["exp_avg_sq"].data_ptr()); shs.append(0 if sh is None else sh.data_ptr()); ns.append(p.numel())
                refreshed.append((p, g, sh))
            if ps:
                self._launch(group, step_no, ps, gs, ms, vs, shs, ns, grad_scale, found_inf)
                self._after(refreshed, found_inf)
        return loss

    @staticmethod
    def _launch(group, step_no, ps, gs, ms, vs, shs, ns, grad_scale, found_inf):
        n = len(ps)
        arr_p = (ctypes.c_void_p * n)(*ps)
        arr_g = (ctypes.c_void_p * n)(*gs)
        arr_m = (ctypes.c_void_p * n)(*ms)
        arr_v = (ctypes.c_void_p * n)(*vs)
        arr_s = (ctypes.c_void_p * n)(*[s or None for s in shs])
        arr_n = (ctypes.c_longlong * n)(*ns)
        b1, b2 = group["betas"]

        def dev_scalar(t, name):
            if t is None:
                return 0
            if not t.is_cuda or t.dtype != torch.float32 or t.numel() != 1:
                raise SwinB200Error(f"optim.Adam: {name} must be a one-element fp32 CUDA tensor")
            return t.data_ptr()

        _lib.call("swinb200_adam_step", n, arr_p, arr_g, arr_m, arr_v, arr_s, arr_n, float(group["lr"]), float(b1), float(b2),
                  float(group["eps"]), float(group["weight_decay"]), int(step_no), dev_scalar(grad_scale, "grad_scale"),
                  dev_scalar(found_inf, "found_inf"), torch.cuda.current_stream().cuda_stream)

    @staticmethod
    def _after(refreshed, found_inf):
        """The kernel wrote through raw pointers: tell autograd the parameters changed, and tell the shadow cache which
        shadows are already current.  (With a GradScaler the step may have been skipped on the device; the shadows are
        then simply still equal to the unchanged masters.)"""
        from .functional import SHADOWS
        for p, _g, sh in refreshed:
            torch.autograd.graph.increment_version(p)
            if sh is not None:
                SHADOWS.mark_fresh(p)
