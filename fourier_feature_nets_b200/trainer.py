"""``FusedTrainer``: the optimisation step of ``Raycaster.fit`` (ray_caster.py:319-329) as two C calls.

``ffn_trainer_backward`` chains forward-with-saves -> loss -> compositing backward -> dgrad -> ffn_wgrad -> head
gradients, ``ffn_trainer_update`` chains ClipAdam -> weight re-pack: ~12 launches issued without interpreter time in
between (through autograd the same step is host bound).  In between the flat gradient buffer can be all-reduced
(data-parallel training).  The arithmetic is identical to the autograd path (same kernels), which stays the general
mechanism (``Raycaster.render`` under autograd, FourierFeatureMLP models, arbitrary losses).

State it owns (all views of three flat fp32 buffers): ``p.grad`` of every parameter, Adam's ``exp_avg`` / ``exp_avg_sq``.
"""
from __future__ import annotations

import ctypes
import math
from ctypes import POINTER, byref, c_float, c_int32, c_int64, c_uint64, c_void_p
from typing import Optional

import torch

from . import _lib
from . import engine as _engine
from .autograd import _flat_grads


class TrainerDesc(ctypes.Structure):
    _fields_ = [("num_linear", c_int32), ("weights", POINTER(c_void_p)), ("biases", POINTER(c_void_p)),
                ("weight_grad_offset", POINTER(c_int64)), ("bias_grad_offset", POINTER(c_int64)),
                ("flat_grad", c_void_p), ("exp_avg", c_void_p), ("exp_avg_sq", c_void_p), ("flat_floats", c_int64)]


def _bind(L):
    if getattr(L, "_trainer_bound", False):
        return
    L.ffn_trainer_create.argtypes = [c_void_p, POINTER(TrainerDesc), POINTER(c_void_p)]
    L.ffn_trainer_destroy.argtypes = [c_void_p]
    L.ffn_trainer_destroy.restype = None
    L.ffn_trainer_workspace_bytes.argtypes = [c_void_p, c_int64, c_int32]
    L.ffn_trainer_workspace_bytes.restype = c_int64
    L.ffn_trainer_backward.argtypes = [c_void_p] + [c_void_p] * 9 + [c_int32, c_uint64, c_int64, c_int32, c_void_p,
                                                                     c_void_p, c_void_p, c_float, c_void_p, c_int64,
                                                                     c_void_p, c_void_p, c_void_p]
    L.ffn_trainer_update.argtypes = [c_void_p] + [c_float] * 9 + [c_void_p, c_int32, c_void_p]
    L.ffn_trainer_set_grad_scale.argtypes = [c_void_p, c_float]
    L._trainer_bound = True


def supported(model) -> bool:
    if getattr(model, "_ffn_kind", None) not in ("nerf", "fourier") or not _engine.supported(model):
        return False
    params = [q for lin in _engine._linear_list(model) for q in (lin.weight, lin.bias)]
    return all(q.is_cuda and q.dtype == torch.float32 and q.is_contiguous() and q.requires_grad for q in params)


class FusedTrainer:
    """Clip + Adam training of a NeRF or FourierFeatureMLP (3 -> [256] * n -> 4) on ray batches against ground-truth tables, two C calls per step.

    ``param_groups`` mimics ``torch.optim.Optimizer`` far enough for ``exponential_lr_decay``."""

    def __init__(self, model, lr: float, weight_decay: float = 0.0, betas=(0.9, 0.999), eps: float = 1e-8,
                 clip_value: float = 0.1, max_norm: float = 0.1):
        if not supported(model):
            raise _lib.FFNError("FusedTrainer needs a float32 NeRF / FourierFeatureMLP of a shape the fused kernels "
                                "cover, on a CUDA device")
        self.model = model
        self.device = next(model.parameters()).device
        self.param_groups = [dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, clip_value=clip_value,
                                  max_norm=max_norm)]
        self.step_count = 0
        L = _lib.lib()
        _bind(L)
        self._L = L
        self.net = _engine.get_engine(model, self.device).net
        lins = _engine._linear_list(model)
        params = []
        for lin in lins:
            params += [lin.weight, lin.bias]
        self.flat_grad, views = _flat_grads(params, self.device)
        self.exp_avg = torch.zeros_like(self.flat_grad)
        self.exp_avg_sq = torch.zeros_like(self.flat_grad)
        base = self.flat_grad.data_ptr()
        offs = [(v.data_ptr() - base) // 4 for v in views]
        for prm, v in zip(params, views):
            prm.grad = v                                   # persistent: every step rewrites the same memory
        model.__dict__["_ffn_flat_grad"] = self.flat_grad  # parallel.allreduce_gradients reduces it in place
        model.__dict__["_ffn_trainer"] = self              # ... and folds the 1 / world size into the update
        self.grad_scale = 1.0
        n = len(lins)
        self._w = (c_void_p * n)(*[lin.weight.data_ptr() for lin in lins])
        self._b = (c_void_p * n)(*[lin.bias.data_ptr() for lin in lins])
        self._gw = (c_int64 * n)(*offs[0::2])
        self._gb = (c_int64 * n)(*offs[1::2])
        self._ptrs = tuple(p.data_ptr() for p in params)
        self._params = params
        desc = TrainerDesc(n, self._w, self._b, self._gw, self._gb, self.flat_grad.data_ptr(),
                           self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(), self.flat_grad.numel())
        h = c_void_p()
        _lib._check(L.ffn_trainer_create(self.net.handle, byref(desc), byref(h)), "ffn_trainer_create")
        self.handle = h
        self._ws: Optional[torch.Tensor] = None
        self._loss = torch.zeros((1,), dtype=torch.float32, device=self.device)
        blocks = 1 + sum((p.numel() + 2047) // 2048 for p in params)
        self._norm = torch.zeros((max(blocks, 1024),), dtype=torch.float32, device=self.device)

    def __del__(self):
        h, self.handle = getattr(self, "handle", None), None
        if h:
            try:
                self._L.ffn_trainer_destroy(h)
            except Exception:
                pass

    def zero_grad(self, set_to_none: bool = False):
        """The step overwrites the gradients; nothing to do (kept for optimiser-like call sites)."""

    def _workspace(self, R: int, S: int) -> torch.Tensor:
        need = int(self._L.ffn_trainer_workspace_bytes(self.net.handle, R, S))
        if need < 0:
            raise _lib.FFNError("ffn_trainer_workspace_bytes failed")
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty((need,), dtype=torch.uint8, device=self.device)
        return self._ws

    def backward(self, rays, gt_colors: torch.Tensor, gt_alphas: Optional[torch.Tensor], alpha_weight: float,
                 lin: torch.Tensor) -> torch.Tensor:
        """Forward + loss + all gradients for one ray batch (a ``RayBundle`` or materialised ``RaySamples`` on the
        device).  Returns the loss as a 0-dim device tensor (valid until the next call)."""
        from .ray_sampler import FocusBundle, RayBundle
        if tuple(p.data_ptr() for p in self._params) != self._ptrs:
            raise _lib.FFNError("model parameters were re-allocated after the FusedTrainer was built")
        if isinstance(rays, FocusBundle):            # t values from the (frozen) coarse model
            with torch.no_grad():
                rays = rays.materialize()
        p = _lib._ptr
        if isinstance(rays, RayBundle):
            R, S = rays.num_rays, rays.num_samples
            keep = [_lib._f32c(rays.starts, "starts"), _lib._f32c(rays.directions, "directions"),
                    _lib._f32c(rays.near, "near"), _lib._f32c(rays.far, "far"), lin,
                    None if rays.jitter is None else _lib._f32c(rays.jitter, "jitter")]
            args = [None, None, None] + keep
            strat, seed = int(rays.stratified), rays.seed
        else:
            R, S = rays.positions.shape[:2]
            keep = [_lib._f32c(rays.positions, "positions"), _lib._f32c(rays.view_directions, "view_directions"),
                    _lib._f32c(rays.t_values, "t_values")]
            args = keep + [None] * 6
            strat, seed = 0, 0
        if S > 256:
            raise _lib.FFNError("training render supports at most 256 samples per ray")
        idx = rays.rays
        if idx.device != self.device or idx.dtype != torch.int64:
            idx = idx.to(self.device, torch.int64)
        idx = idx.contiguous()
        ws = self._workspace(R, S)
        with _lib.on_device(self.device):
            _lib._check(self._L.ffn_trainer_backward(
                self.handle, *[p(t) for t in args], strat, c_uint64(seed & (2 ** 64 - 1)), R, S, p(gt_colors),
                p(gt_alphas), p(idx), float(alpha_weight), p(ws), ws.numel(), p(self._loss), p(self.net._nan_flag),
                _lib._stream()), "ffn_trainer_backward")
        return self._loss[0]

    def update(self):
        """Clip (value, norm) + Adam + re-pack with the current ``param_groups[0]`` hyper-parameters."""
        g = self.param_groups[0]
        self.step_count += 1
        beta1, beta2 = g["betas"]
        with _lib.on_device(self.device):
            _lib._check(self._L.ffn_trainer_update(
                self.handle, g["clip_value"], g["max_norm"], g["lr"], beta1, beta2, g["eps"], g["weight_decay"],
                1.0 - beta1 ** self.step_count, 1.0 - beta2 ** self.step_count, _lib._ptr(self._norm),
                self._norm.numel(), _lib._stream()), "ffn_trainer_update")

    def set_grad_scale(self, scale: float):
        """Multiply every gradient by ``scale`` inside the update kernels (``1 / world_size`` after a summing
        all-reduce): the mean of the data-parallel gradients without another pass over the buffer."""
        if scale != self.grad_scale:
            _lib._check(self._L.ffn_trainer_set_grad_scale(self.handle, float(scale)), "ffn_trainer_set_grad_scale")
            self.grad_scale = float(scale)

    def total_norm(self) -> float:
        return math.sqrt(float(self._norm[0].item()))
