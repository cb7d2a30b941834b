"""Differentiable ``Raycaster.render`` on the CUDA kernels (training step).

forward   ``ffn_train_forward``      fused sampling/encoding/MLP/compositing that also spills what the
                                     backward needs: 16-bit layer outputs (operand dtype), ReLU sign words, encoding rows,
                                     raw network outputs and t values
backward  ``ffn_composite_backward`` d(loss)/d(color, alpha) -> d(loss)/d(raw rgb, sigma) per sample
          ``ffn_train_backward``     the dgrad chain on tcgen05 (transposed bf16 weights) -> dz per layer
          ``ffn_wgrad``              dW = dz^T x and db = column sums of dz for every MMA layer in ONE launch
                                     (split-K tcgen05 GEMM straight from the saved tensors, red.add into one
                                     flat fp32 gradient buffer in the reference's parameter layouts)
          ``ffn_head_wgrad``         the 1- / 3- / 4-row heads on CUDA cores

Gradient operands are bf16 (fp32 accumulate); parity with the fp32 autograd of the reference definition is
checked in tests/test_gpu_training.py (cosine >= 0.999, relative L2 error <= 3e-2 per parameter).
"""
from __future__ import annotations

import ctypes
import math
from ctypes import c_int32, c_int64, c_uint64, c_void_p
from typing import List, Optional

import torch

from . import _lib
from . import engine as _engine


class WgradTensor(ctypes.Structure):
    _fields_ = [("ptr", c_void_p), ("rows", c_int64), ("cols", c_int32), ("slots", c_int32), ("fp16", c_int32)]


class WgradJob(ctypes.Structure):
    _fields_ = [("a_tensor", c_int32), ("a_slot", c_int32), ("a_col0", c_int32), ("n_mtiles", c_int32),
                ("b_tensor", c_int32), ("b_slot", c_int32), ("b_col0", c_int32), ("n_cols", c_int32),
                ("dst", c_void_p), ("dst_stride", c_int32), ("dst_col0", c_int32), ("dst_cols", c_int32),
                ("colmap", c_void_p), ("bias_dst", c_void_p)]


def _bind(L):
    if getattr(L, "_train_bound", False):
        return
    L.ffn_wgrad.argtypes = [ctypes.POINTER(WgradTensor), c_int32, ctypes.POINTER(WgradJob), c_int32, c_void_p]
    L.ffn_mse_loss.argtypes = [c_void_p] * 5 + [c_int64, ctypes.c_float] + [c_void_p] * 4
    L.ffn_train_slots.argtypes = [c_void_p, ctypes.POINTER(c_int32), ctypes.POINTER(c_int32),
                                  ctypes.POINTER(c_int32)]
    L.ffn_net_pack_backward.argtypes = [c_void_p, ctypes.POINTER(c_void_p), c_void_p]
    L.ffn_train_forward.argtypes = [c_void_p] + [c_void_p] * 9 + [c_int32, c_uint64, c_int64, c_int64, c_int32] + \
        [c_void_p] * 9 + [c_void_p]
    L.ffn_composite_backward.argtypes = [c_void_p, c_void_p, c_int64, c_int32, c_void_p, c_void_p, c_void_p,
                                         c_void_p]
    L.ffn_train_backward.argtypes = [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p]
    L.ffn_colsum_bf16.argtypes = [c_void_p, c_int32, c_int64, c_void_p, c_void_p]
    L.ffn_head_wgrad.argtypes = [c_void_p, c_int32, c_int32, c_void_p, c_int64, c_void_p, c_void_p, c_int32, c_int32,
                                 c_void_p]
    L._train_bound = True


def _p(t: Optional[torch.Tensor]) -> c_void_p:
    return c_void_p(0 if t is None else t.data_ptr())


def _bias_grads(L, dz: torch.Tensor) -> torch.Tensor:
    """(stand-alone bias gradients; the training step gets them from ffn_wgrad) Column sums of every saved dz slot in one launch: (n, M, 256) bf16 -> (n, 256) fp32."""
    n, M = dz.shape[0], dz.shape[1]
    out = torch.empty((n, 256), dtype=torch.float32, device=dz.device)
    _lib._check(L.ffn_colsum_bf16(_p(dz), n, M, _p(out), _lib._stream()), "ffn_colsum_bf16")
    return out


def _head_grads(L, d_raw: torch.Tensor, first: int, count: int, h: torch.Tensor, gw: torch.Tensor, gb: torch.Tensor):
    """Gradients of a CUDA-core head into gw (count, cols <= 256), gb (count,):
    gw = d_raw[:, first:first+count]^T @ h[:, :cols], fp32 accumulation."""
    M = d_raw.shape[0]
    _lib._check(L.ffn_head_wgrad(_p(d_raw), first, count, _p(h), M, _p(gw), _p(gb), gw.shape[1],
                                 int(h.dtype == torch.float16), _lib._stream()), "ffn_head_wgrad")


def _wg_tensor(t: torch.Tensor) -> WgradTensor:
    """(slots, rows, cols) bf16 / fp16 tensor -> descriptor."""
    assert t.dim() == 3 and t.is_contiguous() and t.dtype in (torch.bfloat16, torch.float16)
    return WgradTensor(t.data_ptr(), t.shape[1], t.shape[2], t.shape[0], int(t.dtype == torch.float16))


def _wg_job(a_slot, n_mtiles, b_tensor, b_slot, b_col0, n_cols, dst, dst_col0, dst_cols, colmap=None, bias=None):
    return WgradJob(0, a_slot, 0, n_mtiles, b_tensor, b_slot, b_col0, n_cols, dst.data_ptr(), dst.shape[1], dst_col0,
                    dst_cols, 0 if colmap is None else colmap.data_ptr(), 0 if bias is None else bias.data_ptr())


_FLAT_LAYOUT = {}


def _flat_grads(params, device):
    """One zeroed fp32 buffer holding the gradient of every parameter (16-byte aligned views, parameter shapes)."""
    key = tuple(tuple(prm.shape) for prm in params)
    lay = _FLAT_LAYOUT.get(key)
    if lay is None:
        sizes = [(math.prod(shape) + 3) & ~3 for shape in key]
        lay = _FLAT_LAYOUT[key] = (sizes, sum(sizes))
    sizes, total = lay
    flat = torch.zeros((total,), dtype=torch.float32, device=device)
    return flat, [chunk[:math.prod(shape)].view(shape) for chunk, shape in zip(flat.split(sizes), key)]


def _as_param_grads(grads, params):
    out = []
    for g, prm in zip(grads, params):
        if not prm.requires_grad:
            out.append(None)
        elif g.shape == prm.shape and g.dtype == prm.dtype:
            out.append(g)
        else:
            out.append(g.reshape(prm.shape).to(prm.dtype))
    return out


def _run_wgrad(L, tensors, jobs):
    ta = (WgradTensor * len(tensors))(*tensors)
    ja = (WgradJob * len(jobs))(*jobs)
    _lib._check(L.ffn_wgrad(ta, len(tensors), ja, len(jobs), _lib._stream()), "ffn_wgrad")


_PERM_CACHE = {}
_COLMAP_CACHE = {}


def ffmlp_colmaps(E: int, device):
    """Two int32[256] maps for the encoding slots a FourierFeatureMLP's training forward saves (see
    ``RenderFFMLP.backward``): reference weight column of each saved column, -1 = padding.  E = 0: un-encoded MLP."""
    key = ("ffmlp", E, str(device))
    if key not in _COLMAP_CACHE:
        def dst(mc):
            if E:
                e = mc >> 1
                return (mc & 1) * E + e if e < E else -1
            return mc if mc < 3 else -1
        a = [dst(c) for c in range(256)]
        if E:
            b = [dst(320 + c) if c < 192 else dst(256 + c - 192) for c in range(256)]
        else:
            b = [dst(c - 192) if c >= 192 else -1 for c in range(256)]
        _COLMAP_CACHE[key] = (torch.tensor(a, dtype=torch.int32, device=device),
                              torch.tensor(b, dtype=torch.int32, device=device))
    return _COLMAP_CACHE[key]


def enc_colmap(num_freq: int, include_inputs: bool, first: int, device) -> torch.Tensor:
    """int32[64]: reference weight column (offset by ``first``) of each of OUR encoding-chunk columns, -1 = unused."""
    key = (num_freq, bool(include_inputs), first, str(device))
    if key not in _COLMAP_CACHE:
        perm = _enc_permutation(num_freq, include_inputs, "cpu")
        cm = torch.full((64,), -1, dtype=torch.int32)
        cm[perm] = torch.arange(len(perm), dtype=torch.int32) + first
        _COLMAP_CACHE[key] = cm.to(device)
    return _COLMAP_CACHE[key]


def enc_permutation(num_freq: int, include_inputs: bool, device) -> torch.Tensor:
    """index of OUR encoding-chunk column for every reference column of [cos(3F) | sin(3F) | x(3)]."""
    key = (num_freq, bool(include_inputs), str(device))
    if key not in _PERM_CACHE:
        _PERM_CACHE[key] = _enc_permutation(num_freq, include_inputs, device)
    return _PERM_CACHE[key]


def _enc_permutation(num_freq: int, include_inputs: bool, device) -> torch.Tensor:
    idx = []
    for sn in (0, 1):
        for k in range(num_freq):
            for j in range(3):
                idx.append(6 * k + 2 * j + sn)
    if include_inputs:
        idx += [60, 61, 62]
    return torch.tensor(idx, dtype=torch.long, device=device)


class RenderNeRF(torch.autograd.Function):
    """(color, alpha, depth) = render(model, rays); gradients flow to the model parameters only."""

    @staticmethod
    def forward(ctx, model, spec, include_depth, *params):
        L = _lib.lib()
        _bind(L)
        device = params[0].device
        eng = _engine.get_engine(model, device)
        net = eng.net
        ns, nm, nd = c_int32(), c_int32(), c_int32()
        _lib._check(L.ffn_train_slots(net.handle, ctypes.byref(ns), ctypes.byref(nm), ctypes.byref(nd)),
                    "ffn_train_slots")
        R, S = spec["R"], spec["S"]
        M = R * S
        f32 = dict(dtype=torch.float32, device=device)
        color = torch.empty((R, 3), **f32)
        alpha = torch.empty((R,), **f32)
        depth = torch.empty((R,), **f32) if include_depth else None
        raw = torch.empty((M, 4), **f32)
        act_dtype = torch.bfloat16 if eng.operand == "bf16" else torch.float16     # saves are in the operand dtype
        save_h = torch.empty((ns.value, M, 256), dtype=act_dtype, device=device)
        save_mask = torch.empty((nm.value, M, 8), dtype=torch.int32, device=device)
        save_enc = torch.empty((2, M, 64), dtype=act_dtype, device=device)
        if spec["mode"] == "rays":
            t_vals = torch.empty((R, S), **f32)
            a = spec
            args = [None, None, None, a["starts"], a["directions"], a["near"], a["far"], a["lin"], a["jitter"]]
            strat, seed = int(a["stratified"]), a["seed"]
        else:
            t_vals = spec["t_values"]
            args = [spec["positions"], spec["view_directions"], spec["t_values"], None, None, None, None, None, None]
            strat, seed = 0, 0
        keep = [t for t in args if t is not None]
        with _lib.on_device(device):
            _lib._check(L.ffn_train_forward(
                net.handle, *[_p(t) for t in args], strat, c_uint64(seed & (2 ** 64 - 1)), 0, R, S,
                _p(color), _p(alpha), _p(depth), _p(raw), _p(t_vals if spec["mode"] == "rays" else None),
                _p(save_h), _p(save_mask), _p(save_enc), _p(net._nan_flag), _lib._stream()), "ffn_train_forward")
        ctx.model, ctx.net, ctx.R, ctx.S, ctx.n_dz = model, net, R, S, nd.value
        ctx.save_for_backward(raw, t_vals, save_h, save_mask, save_enc, *params)
        ctx.keep = keep
        ctx.mark_non_differentiable(*([depth] if depth is not None else []))
        if depth is None:
            return color, alpha
        return color, alpha, depth

    @staticmethod
    def backward(ctx, g_color, g_alpha, *_):
        L = _lib.lib()
        raw, t_vals, save_h, save_mask, save_enc, *params = ctx.saved_tensors
        model, net, R, S = ctx.model, ctx.net, ctx.R, ctx.S
        M = R * S
        device = raw.device
        gc = (g_color if g_color is not None else torch.zeros((R, 3), device=device)).contiguous().float()
        ga = g_alpha.contiguous().float() if g_alpha is not None else None
        d_raw = torch.empty((M, 4), dtype=torch.float32, device=device)
        dz = torch.empty((ctx.n_dz, M, 256), dtype=torch.bfloat16, device=device)
        lins = _engine._linear_list(model)
        wptr = (c_void_p * len(lins))(*[l.weight.data_ptr() for l in lins])
        with _lib.on_device(device):
            _lib._check(L.ffn_composite_backward(_p(raw), _p(t_vals), R, S, _p(gc), _p(ga), _p(d_raw),
                                                 _lib._stream()), "ffn_composite_backward")
            _lib._check(L.ffn_net_pack_backward(net.handle, wptr, _lib._stream()), "ffn_net_pack_backward")
            _lib._check(L.ffn_train_backward(net.handle, _p(d_raw), _p(save_mask), M, _p(dz), _lib._stream()),
                        "ffn_train_backward")

        p = model.params
        nL = p["num_layers"]
        skips = set(p["skips"])
        nf_p, nf_v, inc = p["num_freq_pos"], p["num_freq_view"], p["include_inputs"]
        # params = [W0, b0, ..., W_{L-1}, b_{L-1}, W_op, b_op, W_bott, b_bott, W_hv, b_hv, W_rgb, b_rgb]
        flat, grads = _flat_grads(params, device)
        W = lambda i: grads[2 * i]          # noqa: E731
        B = lambda i: grads[2 * i + 1]      # noqa: E731
        SH, ENC = 1, 2                       # tensor indices: 0 = dz, 1 = save_h, 2 = save_enc
        jobs = []
        # trunk layers (nerf_model.py:111-116): layer i reads save_h[i-1]; layer 0 and the skip layers read enc_p
        for i in range(nL):
            if i == 0:
                jobs.append(_wg_job(0, 2, ENC, 0, 0, 64, W(0), 0, 64, enc_colmap(nf_p, inc, 0, device), B(0)))
            else:
                jobs.append(_wg_job(i, 2, SH, i - 1, 0, 256, W(i), 0, 256, None, B(i)))
                if i in skips:
                    jobs.append(_wg_job(i, 2, ENC, 0, 0, 64, W(i), 0, 64, enc_colmap(nf_p, inc, 256, device)))
        # bottleneck (nerf_model.py:119) reads the last trunk activation
        jobs.append(_wg_job(nL, 2, SH, nL - 1, 0, 256, W(nL + 1), 0, 256, None, B(nL + 1)))
        # hidden_view (nerf_model.py:121-122): 128 outputs, input [bottleneck | enc_view]
        jobs.append(_wg_job(nL + 1, 1, SH, nL, 0, 256, W(nL + 2), 0, 256, None, B(nL + 2)))
        jobs.append(_wg_job(nL + 1, 1, ENC, 1, 0, 64, W(nL + 2), 0, 64, enc_colmap(nf_v, inc, 256, device)))
        with _lib.on_device(device):
            _run_wgrad(L, [_wg_tensor(dz), _wg_tensor(save_h), _wg_tensor(save_enc)], jobs)
            # opacity_out (nerf_model.py:118): sigma_raw = w_op . h_L + b;  color_out (:123): the saved slot holds
            # relu(hidden_view) in its first 128 columns
            _head_grads(L, d_raw, 3, 1, save_h[nL - 1], grads[2 * nL], grads[2 * nL + 1])
            _head_grads(L, d_raw, 0, 3, save_h[nL + 1], grads[2 * nL + 6], grads[2 * nL + 7])
        model.__dict__["_ffn_flat_grad"] = flat      # every gradient is a view of it (parallel.allreduce_gradients)

        return (None, None, None, *_as_param_grads(grads, params))


class RenderFFMLP(torch.autograd.Function):
    """Same as :class:`RenderNeRF` for ``FourierFeatureMLP`` models (3 -> [256]*H -> 4, no view branch)."""

    @staticmethod
    def forward(ctx, model, spec, include_depth, *params):
        L = _lib.lib()
        _bind(L)
        device = params[0].device
        eng = _engine.get_engine(model, device)
        net = eng.net
        ns, nm, nd = c_int32(), c_int32(), c_int32()
        _lib._check(L.ffn_train_slots(net.handle, ctypes.byref(ns), ctypes.byref(nm), ctypes.byref(nd)), "ffn_train_slots")
        R, S = spec["R"], spec["S"]
        M = R * S
        f32 = dict(dtype=torch.float32, device=device)
        color, alpha = torch.empty((R, 3), **f32), torch.empty((R,), **f32)
        depth = torch.empty((R,), **f32) if include_depth else None
        raw = torch.empty((M, 4), **f32)
        act_dtype = torch.bfloat16 if eng.operand == "bf16" else torch.float16
        save_h = torch.empty((ns.value, M, 256), dtype=act_dtype, device=device)
        save_mask = torch.empty((nm.value, M, 8), dtype=torch.int32, device=device)
        if spec["mode"] == "rays":
            t_vals = torch.empty((R, S), **f32)
            a = spec
            args = [None, None, None, a["starts"], a["directions"], a["near"], a["far"], a["lin"], a["jitter"]]
            strat, seed = int(a["stratified"]), a["seed"]
        else:
            t_vals = spec["t_values"]
            args = [spec["positions"], None, spec["t_values"], None, None, None, None, None, None]
            strat, seed = 0, 0
        with _lib.on_device(device):
            _lib._check(L.ffn_train_forward(
                net.handle, *[_p(t) for t in args], strat, c_uint64(seed & (2 ** 64 - 1)), 0, R, S,
                _p(color), _p(alpha), _p(depth), _p(raw), _p(t_vals if spec["mode"] == "rays" else None),
                _p(save_h), _p(save_mask), _p(None), _p(net._nan_flag), _lib._stream()), "ffn_train_forward")
        ctx.model, ctx.net, ctx.R, ctx.S, ctx.n_dz, ctx.spec = model, net, R, S, nd.value, spec
        ctx.save_for_backward(raw, t_vals, save_h, save_mask, *params)
        ctx.mark_non_differentiable(*([depth] if depth is not None else []))
        return (color, alpha) if depth is None else (color, alpha, depth)

    @staticmethod
    def backward(ctx, g_color, g_alpha, *_):
        L = _lib.lib()
        raw, t_vals, save_h, save_mask, *params = ctx.saved_tensors
        model, net, R, S = ctx.model, ctx.net, ctx.R, ctx.S
        M = R * S
        device = raw.device
        gc = (g_color if g_color is not None else torch.zeros((R, 3), device=device)).contiguous().float()
        ga = g_alpha.contiguous().float() if g_alpha is not None else None
        d_raw = torch.empty((M, 4), dtype=torch.float32, device=device)
        dz = torch.empty((ctx.n_dz, M, 256), dtype=torch.bfloat16, device=device)
        lins = _engine._linear_list(model)
        wptr = (c_void_p * len(lins))(*[l.weight.data_ptr() for l in lins])
        with _lib.on_device(device):
            _lib._check(L.ffn_composite_backward(_p(raw), _p(t_vals), R, S, _p(gc), _p(ga), _p(d_raw), _lib._stream()),
                        "ffn_composite_backward")
            _lib._check(L.ffn_net_pack_backward(net.handle, wptr, _lib._stream()), "ffn_net_pack_backward")
            _lib._check(L.ffn_train_backward(net.handle, _p(d_raw), _p(save_mask), M, _p(dz), _lib._stream()),
                        "ffn_train_backward")
        H = len(lins) - 1
        # layer 0 input: the encoding rows the forward saved (OUR column order: column 2e + s of feature e, s = 0 cos /
        # 1 sin) in the two extra slots of save_h -- slot H: encoding chunks 0..3, slot H + 1: columns [0,192) chunks
        # 5..7, columns [192,256) chunk 4 (or the raw inputs of the un-encoded MLP); ffn_wgrad scatters the weight
        # gradient to the reference's column order s*E + e (fourier_feature_models.py:66-68) through column maps
        E = 0 if model.b_values is None else model.b_values.shape[1]
        cm_a, cm_b = ffmlp_colmaps(E, device)
        nch = (max(2 * E, 3) + 63) // 64
        n1 = min(nch, 4) if E else 0
        flat, grads = _flat_grads(params, device)
        jobs = []
        bias0 = grads[1]
        if n1:
            jobs.append(_wg_job(0, 2, 1, H, 0, 64 * n1, grads[0], 0, 64 * n1, cm_a, bias0))
            bias0 = None
        if nch > 4 or not E:
            jobs.append(_wg_job(0, 2, 1, H + 1, 0, 256, grads[0], 0, 256, cm_b, bias0))
        for i in range(1, H):
            jobs.append(_wg_job(i, 2, 1, i - 1, 0, 256, grads[2 * i], 0, 256, None, grads[2 * i + 1]))
        with _lib.on_device(device):
            _run_wgrad(L, [_wg_tensor(dz), _wg_tensor(save_h)], jobs)
            _head_grads(L, d_raw, 0, 4, save_h[H - 1], grads[2 * H], grads[2 * H + 1])    # final Linear 256 -> 4
        model.__dict__["_ffn_flat_grad"] = flat
        return (None, None, None, *_as_param_grads(grads, params))


class MSELoss(torch.autograd.Function):
    """loss = mean((colors[rays] - color)^2) + alpha_weight * mean((alphas[rays] - alpha)^2) with the ground-truth
    colour zeroed where the ground-truth alpha is 0 (ImageDataset.render/.loss, image_dataset.py:224-262): value and
    gradient in ONE launch (``ffn_mse_loss``) instead of ~10 forward and ~12 backward element-wise launches."""

    @staticmethod
    def forward(ctx, color, alpha, gt_colors, gt_alphas, rays, alpha_weight):
        L = _lib.lib()
        _bind(L)
        R = color.shape[0]
        color = _lib._f32c(color, "color")
        alpha = None if gt_alphas is None else _lib._f32c(alpha, "alpha")
        out = torch.empty((1 + 4 * R,), dtype=torch.float32, device=color.device)
        loss, g_color, g_alpha = out[0], out[1:1 + 3 * R].view(R, 3), out[1 + 3 * R:]
        with _lib.on_device(color.device):
            _lib._check(L.ffn_mse_loss(_p(color), _p(alpha), _p(gt_colors), _p(gt_alphas), _p(rays), R,
                                       float(alpha_weight), _p(out), c_void_p(g_color.data_ptr()),
                                       c_void_p(g_alpha.data_ptr()), _lib._stream()), "ffn_mse_loss")
        ctx.save_for_backward(g_color, g_alpha)
        ctx.has_alpha = gt_alphas is not None
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        g_color, g_alpha = ctx.saved_tensors
        return g_color * grad_out, (g_alpha * grad_out if ctx.has_alpha else None), None, None, None, None


def render_nerf_train(model, ray_samples, include_depth: bool, lin_fn):
    """Entry used by ``Raycaster.render`` under autograd for NeRF models on CUDA."""
    from .ray_sampler import RayBundle
    lins = _engine._linear_list(model)
    params = []
    for l in lins:
        params += [l.weight, l.bias]
    if isinstance(ray_samples, RayBundle):
        b = ray_samples
        R, S = b.num_rays, b.num_samples
        spec = {"mode": "rays", "R": R, "S": S, "starts": _lib._f32c(b.starts, "starts"),
                "directions": _lib._f32c(b.directions, "directions"), "near": _lib._f32c(b.near, "near"),
                "far": _lib._f32c(b.far, "far"), "lin": lin_fn(S),
                "jitter": None if b.jitter is None else _lib._f32c(b.jitter, "jitter"),
                "stratified": b.stratified, "seed": b.seed}
    else:
        R, S = ray_samples.positions.shape[:2]
        spec = {"mode": "samples", "R": R, "S": S,
                "positions": _lib._f32c(ray_samples.positions, "positions"),
                "view_directions": (_lib._f32c(ray_samples.view_directions, "view_directions")
                                    if model.use_view else None),
                "t_values": _lib._f32c(ray_samples.t_values, "t_values")}
    if S > 256:
        raise _lib.FFNError("training render supports at most 256 samples per ray")
    fn = RenderNeRF if model._ffn_kind == "nerf" else RenderFFMLP
    res = fn.apply(model, spec, bool(include_depth), *params)
    if include_depth:
        return res
    return res[0], res[1], None
