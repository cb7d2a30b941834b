"""``ClipAdam``: the optimiser step of ``Raycaster.fit`` (ray_caster.py:327-329) as two kernel launches.

The reference runs ``clip_grad_value_(params, 0.1)``, ``clip_grad_norm_(params, 0.1)`` and ``torch.optim.Adam.step()``
back to back: ~15 small launches over 24 tensors.  ``ffn_clip_adam`` does the same arithmetic (gradients clamped and
rescaled in place, then Adam with L2 weight decay and bias correction) over all tensors at once.  CUDA only: there is no
CPU implementation -- on a CPU model ``Raycaster.fit`` uses the PyTorch calls of the reference.
"""
from __future__ import annotations

import ctypes
import math
from ctypes import c_float, c_int32, c_int64, c_void_p

import torch

from . import _lib


class AdamTensor(ctypes.Structure):
    _fields_ = [("param", c_void_p), ("grad", c_void_p), ("exp_avg", c_void_p), ("exp_avg_sq", c_void_p),
                ("numel", c_int64)]


def _bind(L):
    if getattr(L, "_optim_bound", False):
        return
    L.ffn_clip_adam.argtypes = [ctypes.POINTER(AdamTensor), c_int32] + [c_float] * 9 + [c_void_p, c_int32, c_void_p]
    L._optim_bound = True


class ClipAdam(torch.optim.Optimizer):
    """Adam (``torch.optim.Adam`` semantics: ``lr, betas, eps, weight_decay``) preceded by element-wise gradient clipping
    to ``[-clip_value, clip_value]`` and global-norm clipping to ``max_norm`` (``<= 0`` disables either)."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, clip_value=0.1, max_norm=0.1):
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, clip_value=clip_value,
                        max_norm=max_norm)
        super().__init__(params, defaults)
        if len(self.param_groups) != 1:
            raise ValueError("ClipAdam clips the norm over ONE parameter group (as Raycaster.fit uses it)")
        self._norm_sq = None
        self._key = None
        self._arr = None

    def total_norm(self) -> float:
        """Norm of the value-clipped gradients of the last step (one device->host read)."""
        return math.sqrt(float(self._norm_sq[0].item())) if self._norm_sq is not None else float("nan")

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        L = _lib.lib()
        _bind(L)
        group = self.param_groups[0]
        params = [p for p in group["params"] if p.grad is not None]
        if not params:
            return loss
        device = params[0].device
        key = tuple(id(p) for p in params)
        if self._key != key:            # (re)build the descriptor table; afterwards only the grad pointers move
            for p in params:
                if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
                    raise _lib.FFNError("ClipAdam needs contiguous float32 CUDA parameters")
                state = self.state[p]
                if not state:
                    state["step"] = 0
                    state["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    state["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
            self._arr = (AdamTensor * len(params))(*[
                AdamTensor(p.data_ptr(), 0, self.state[p]["exp_avg"].data_ptr(), self.state[p]["exp_avg_sq"].data_ptr(),
                           p.numel()) for p in params])
            self._key = key
            self._need = 1 + sum((p.numel() + 2047) // 2048 for p in params)
        arr = self._arr
        step = self.state[params[0]]["step"] + 1
        for i, p in enumerate(params):
            g = p.grad
            if g.dtype != torch.float32 or not g.is_contiguous():
                p.grad = g = g.float().contiguous()
            st = self.state[p]
            arr[i].param, arr[i].grad = p.data_ptr(), g.data_ptr()
            arr[i].exp_avg, arr[i].exp_avg_sq = st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr()
            st["step"] = step
        if self._norm_sq is None or self._norm_sq.device != device or self._norm_sq.numel() < self._need:
            self._norm_sq = torch.zeros((max(self._need, 1024),), dtype=torch.float32, device=device)
        beta1, beta2 = group["betas"]
        with _lib.on_device(device):
            _lib._check(L.ffn_clip_adam(arr, len(params), group["clip_value"], group["max_norm"], group["lr"], beta1,
                                        beta2, group["eps"], group["weight_decay"], 1.0 - beta1 ** step,
                                        1.0 - beta2 ** step, self._norm_sq.data_ptr(), self._norm_sq.numel(), _lib._stream()),
                        "ffn_clip_adam")
        return loss
