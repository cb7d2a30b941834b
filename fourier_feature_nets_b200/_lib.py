"""ctypes binding of ``libffn_b200.so`` (C ABI in ``include/ffn_b200.h``).

There is no fallback: if the shared library is missing or a call fails, an
exception is raised.  PyTorch is used only for device memory and the current
stream; every signature crossing the boundary is plain pointers and sizes.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from ctypes import (POINTER, Structure, byref, c_char_p, c_float, c_int32, c_int64,
                    c_uint64, c_void_p)
from typing import List, Optional, Sequence

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# FFN_LIB: load another build of the same ABI (A/B timing of two kernel versions inside one GPU session)
LIB_PATH = os.environ.get("FFN_LIB") or os.path.join(_HERE, "libffn_b200.so")
CSRC = os.path.join(_HERE, "csrc")

FFN_MAX_LAYERS = 16
FFN_MAX_FREQS = 10
OPERAND_FP16 = 0
OPERAND_BF16 = 1
OPERAND_FP16X3 = 2      # NeRF inference: hi + residual split of every operand, three UMMAs per product

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC"]


class FFNError(RuntimeError):
    pass


class NerfDesc(Structure):
    _fields_ = [("num_layers", c_int32), ("num_channels", c_int32),
                ("num_freq_pos", c_int32), ("num_freq_view", c_int32),
                ("include_inputs", c_int32), ("num_skips", c_int32),
                ("skips", c_int32 * FFN_MAX_LAYERS),
                ("freq_pos", c_float * FFN_MAX_FREQS), ("freq_view", c_float * FFN_MAX_FREQS),
                ("operand_dtype", c_int32)]


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/ffn_b200.cu for sm_100a in-tree (nvcc cross-compiles without a GPU)."""
    src = os.path.join(CSRC, "ffn_b200.cu")
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [
        os.path.join(_HERE, "..", "include", "ffn_b200.h")]
    if not force and os.path.exists(LIB_PATH):
        if os.path.getmtime(LIB_PATH) >= max(os.path.getmtime(d) for d in deps):
            return LIB_PATH
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH, src]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise FFNError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return LIB_PATH


_lib = None


def lib() -> ctypes.CDLL:
    """Load the library (once).  Raises FFNError if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FFNError(
            "libffn_b200.so is not built (%s). Run `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `python -m fourier_feature_nets_b200.build`. There is no CPU fallback for the "
            "CUDA render path." % LIB_PATH)
    L = ctypes.CDLL(LIB_PATH)
    L.ffn_version.restype = c_int32
    L.ffn_last_error.restype = c_char_p
    L.ffn_launch_count.restype = c_int64
    L.ffn_nerf_create.argtypes = [POINTER(NerfDesc), POINTER(c_void_p)]
    L.ffn_ffmlp_create.argtypes = [c_int32, c_int32, c_int32, c_void_p, c_void_p, c_int32,
                                   POINTER(c_void_p)]
    L.ffn_net_destroy.argtypes = [c_void_p]
    L.ffn_net_destroy.restype = None
    L.ffn_net_num_linear.argtypes = [c_void_p]
    L.ffn_net_pack.argtypes = [c_void_p, POINTER(c_void_p), POINTER(c_void_p), c_void_p]
    L.ffn_mlp_forward.argtypes = [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p]
    L.ffn_debug_layer.argtypes = [c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_void_p, c_void_p]
    L.ffn_render_samples.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int32,
                                     c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]
    L.ffn_render_rays.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                  c_void_p, c_int32, c_uint64, c_int64, c_int64, c_int32,
                                  c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]
    L.ffn_composite.argtypes = [c_void_p, c_void_p, c_int64, c_int32, c_void_p, c_void_p, c_void_p,
                                c_void_p, c_void_p, c_void_p]
    L.ffn_blend_weights.argtypes = [c_void_p, c_void_p, c_int64, c_int32, c_void_p, c_void_p]
    L.ffn_generate_rays.argtypes = [c_void_p, c_void_p, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_float),
                                    c_int32, c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]
    L.ffn_voxels_forward.argtypes = [c_void_p, ctypes.POINTER(ctypes.c_float), c_int32, c_float, c_void_p, c_int64,
                                     c_void_p, c_void_p]
    L.ffn_focus_t.argtypes = [c_void_p, c_int32] + [c_void_p] * 9 + [c_int32, c_uint64, c_int64, c_int64, c_int32,
                                                                     c_void_p, c_void_p]
    L.ffn_focus_sample.argtypes = [c_void_p] + [c_void_p] * 11 + [c_int32, c_uint64, c_int64, c_int64, c_int32,
                                                                  c_void_p, c_void_p]
    L.ffn_render_rays_t.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_void_p, c_void_p,
                                    c_void_p, c_void_p, c_void_p]
    _lib = L
    return L


EXPORTED_SYMBOLS = [
    "ffn_version", "ffn_last_error", "ffn_nerf_create", "ffn_ffmlp_create", "ffn_net_destroy",
    "ffn_net_num_linear", "ffn_net_pack", "ffn_mlp_forward", "ffn_render_samples",
    "ffn_render_rays", "ffn_composite", "ffn_blend_weights", "ffn_debug_layer", "ffn_debug_stats", "ffn_launch_count",
    "ffn_focus_t", "ffn_focus_sample", "ffn_render_rays_t", "ffn_generate_rays", "ffn_voxels_forward",
    "ffn_train_slots", "ffn_net_pack_backward", "ffn_train_forward", "ffn_composite_backward",
    "ffn_train_backward", "ffn_colsum_bf16", "ffn_head_wgrad", "ffn_wgrad", "ffn_clip_adam", "ffn_mse_loss",
    "ffn_trainer_create", "ffn_trainer_destroy", "ffn_trainer_workspace_bytes", "ffn_trainer_backward",
    "ffn_trainer_update", "ffn_trainer_set_grad_scale",
]


def _check(rc: int, what: str):
    if rc != 0:
        raise FFNError("%s failed: %s" % (what, lib().ffn_last_error().decode()))


def launch_count() -> int:
    return int(lib().ffn_launch_count())


def _stream() -> c_void_p:
    """Raw handle of PyTorch's current stream on the current device (one C call: ``torch.cuda.current_stream()``
    builds a Python ``Stream`` object and costs ~20 us, nine times per training step)."""
    return c_void_p(torch._C._cuda_getCurrentRawStream(torch._C._cuda_getDevice()))


class _NullCtx:
    def __enter__(self):
        return None

    def __exit__(self, *exc):
        return False


_NULL_CTX = _NullCtx()


def on_device(device):
    """``torch.cuda.device(device)`` only when it is not already the current device (the context manager costs
    ~10 us, the check one C call; a training step enters it a dozen times)."""
    idx = device.index if isinstance(device, torch.device) else int(device)
    if idx is None or idx == torch._C._cuda_getDevice():
        return _NULL_CTX
    return torch.cuda.device(idx)


def _ptr(t: Optional[torch.Tensor]) -> c_void_p:
    if t is None:
        return c_void_p(0)
    return c_void_p(t.data_ptr())


def _f32c(t: torch.Tensor, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise FFNError("%s must be a CUDA tensor (no CPU path in libffn_b200)" % name)
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


class Net:
    """Owns an ``ffn_net_t`` handle."""

    def __init__(self, handle: c_void_p, device: torch.device):
        self.handle = handle
        self.device = device
        self.num_linear = lib().ffn_net_num_linear(handle)
        self._keepalive: List[torch.Tensor] = []
        self._nan_flag = torch.zeros(1, dtype=torch.int32, device=device)

    def __del__(self):
        try:
            if self.handle:
                lib().ffn_net_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    # -- construction ------------------------------------------------------------------
    @staticmethod
    def nerf(num_layers: int, num_channels: int, freq_pos: Sequence[float],
             freq_view: Sequence[float], skips: Sequence[int], include_inputs: bool,
             device: torch.device, operand: int = OPERAND_FP16) -> "Net":
        d = NerfDesc()
        d.num_layers, d.num_channels = num_layers, num_channels
        d.num_freq_pos, d.num_freq_view = len(freq_pos), len(freq_view)
        d.include_inputs = int(bool(include_inputs))
        skips = sorted(set(int(s) for s in skips))
        if len(skips) > FFN_MAX_LAYERS or len(freq_pos) > FFN_MAX_FREQS or len(freq_view) > FFN_MAX_FREQS:
            raise FFNError("NeRF shape not supported by libffn_b200")
        d.num_skips = len(skips)
        for i, s in enumerate(skips):
            d.skips[i] = s
        for i, f in enumerate(freq_pos):
            d.freq_pos[i] = float(f)
        for i, f in enumerate(freq_view):
            d.freq_view[i] = float(f)
        d.operand_dtype = operand
        h = c_void_p()
        with on_device(device):
            _check(lib().ffn_nerf_create(byref(d), byref(h)), "ffn_nerf_create")
        return Net(h, device)

    @staticmethod
    def ffmlp(num_hidden: int, num_channels: int, a_values: Optional[torch.Tensor],
              b_values: Optional[torch.Tensor], device: torch.device,
              operand: int = OPERAND_FP16) -> "Net":
        h = c_void_p()
        with on_device(device):
            if b_values is None:
                _check(lib().ffn_ffmlp_create(num_hidden, num_channels, 0, None, None, operand, byref(h)),
                       "ffn_ffmlp_create")
            else:
                a = a_values.detach().float().cpu().contiguous()
                b = b_values.detach().float().cpu().contiguous()
                if b.shape[0] != 3:
                    raise FFNError("libffn_b200 renders 3-D inputs only")
                _check(lib().ffn_ffmlp_create(num_hidden, num_channels, b.shape[1],
                                              c_void_p(a.data_ptr()), c_void_p(b.data_ptr()), operand,
                                              byref(h)), "ffn_ffmlp_create")
        return Net(h, device)

    # -- weights -----------------------------------------------------------------------
    def pack(self, weights: Sequence[torch.Tensor], biases: Sequence[torch.Tensor]):
        if len(weights) != self.num_linear or len(biases) != self.num_linear:
            raise FFNError("expected %d (weight, bias) pairs" % self.num_linear)
        ws = [_f32c(w.detach(), "weight") for w in weights]
        bs = [_f32c(b.detach(), "bias") for b in biases]
        self._keepalive = ws + bs
        wa = (c_void_p * len(ws))(*[w.data_ptr() for w in ws])
        ba = (c_void_p * len(bs))(*[b.data_ptr() for b in bs])
        with on_device(self.device):
            _check(lib().ffn_net_pack(self.handle, wa, ba, _stream()), "ffn_net_pack")

    # -- launches ----------------------------------------------------------------------
    def mlp_forward(self, positions: torch.Tensor, views: Optional[torch.Tensor]) -> torch.Tensor:
        pos = _f32c(positions, "positions")
        v = _f32c(views, "views") if views is not None else None
        n = pos.shape[0]
        out = torch.empty((n, 4), dtype=torch.float32, device=pos.device)
        with on_device(self.device):
            _check(lib().ffn_mlp_forward(self.handle, _ptr(pos), _ptr(v), n, _ptr(out), _stream()),
                   "ffn_mlp_forward")
        return out

    def debug_layer(self, positions: torch.Tensor, views: Optional[torch.Tensor], layer: int) -> torch.Tensor:
        pos = _f32c(positions, "positions")
        v = _f32c(views, "views") if views is not None else None
        n = pos.shape[0]
        out = torch.zeros((n, 256), dtype=torch.float32, device=pos.device)
        with on_device(self.device):
            _check(lib().ffn_debug_layer(self.handle, _ptr(pos), _ptr(v), n, layer, _ptr(out), _stream()),
                   "ffn_debug_layer")
        return out

    def render_samples(self, positions, view_directions, t_values, include_depth: bool):
        pos = _f32c(positions, "positions")
        R, S = pos.shape[0], pos.shape[1]
        v = _f32c(view_directions, "view_directions") if view_directions is not None else None
        t = _f32c(t_values, "t_values")
        color = torch.empty((R, 3), dtype=torch.float32, device=pos.device)
        alpha = torch.empty((R,), dtype=torch.float32, device=pos.device)
        depth = torch.empty((R,), dtype=torch.float32, device=pos.device) if include_depth else None
        with on_device(self.device):
            _check(lib().ffn_render_samples(self.handle, _ptr(pos), _ptr(v), _ptr(t), R, S, _ptr(color),
                                            _ptr(alpha), _ptr(depth), _ptr(self._nan_flag), _stream()),
                   "ffn_render_samples")
        return color, alpha, depth

    def render_rays(self, starts, directions, near, far, lin, jitter, stratified: bool, seed: int,
                    ray_offset: int, num_samples: int, include_depth: bool, want_t: bool = False):
        o = _f32c(starts, "starts")
        d = _f32c(directions, "directions")
        nr = _f32c(near, "near")
        fr = _f32c(far, "far")
        ln = _f32c(lin, "lin")
        j = _f32c(jitter, "jitter") if jitter is not None else None
        R = o.shape[0]
        color = torch.empty((R, 3), dtype=torch.float32, device=o.device)
        alpha = torch.empty((R,), dtype=torch.float32, device=o.device)
        depth = torch.empty((R,), dtype=torch.float32, device=o.device) if include_depth else None
        t_out = torch.empty((R, num_samples), dtype=torch.float32, device=o.device) if want_t else None
        with on_device(self.device):
            _check(lib().ffn_render_rays(self.handle, _ptr(o), _ptr(d), _ptr(nr), _ptr(fr), _ptr(ln),
                                         _ptr(j), int(bool(stratified)), c_uint64(seed & (2**64 - 1)),
                                         ray_offset, R, num_samples, _ptr(color), _ptr(alpha),
                                         _ptr(depth), _ptr(t_out), _ptr(self._nan_flag), _stream()),
                   "ffn_render_rays")
        return color, alpha, depth, t_out

    def focus_sample(self, starts, directions, near, far, near_u, far_u, lin_c, lin_u, jitter_u, u_focus,
                     stratified: bool, seed: int, num_samples: int, ray_offset: int = 0) -> torch.Tensor:
        """coarse sigma pass of THIS net + inverse-transform sampling -> sorted t (R, num_samples)."""
        o, d = _f32c(starts, "starts"), _f32c(directions, "directions")
        nr, fr = _f32c(near, "near"), _f32c(far, "far")
        nu, fu = _f32c(near_u, "near_u"), _f32c(far_u, "far_u")
        lc, lu = _f32c(lin_c, "lin_c"), _f32c(lin_u, "lin_u")
        ju = _f32c(jitter_u, "jitter_u") if jitter_u is not None else None
        uf = _f32c(u_focus, "u_focus") if u_focus is not None else None
        R = o.shape[0]
        t = torch.empty((R, num_samples), dtype=torch.float32, device=o.device)
        with on_device(self.device):
            _check(lib().ffn_focus_sample(self.handle, _ptr(o), _ptr(d), _ptr(nr), _ptr(fr), _ptr(nu), _ptr(fu),
                                          _ptr(lc), _ptr(lu), _ptr(lc), _ptr(ju), _ptr(uf), int(bool(stratified)),
                                          c_uint64(seed & (2**64 - 1)), ray_offset, R, num_samples, _ptr(t), _stream()),
                   "ffn_focus_sample")
        return t

    def render_rays_t(self, starts, directions, t_values, include_depth: bool):
        o, d, t = _f32c(starts, "starts"), _f32c(directions, "directions"), _f32c(t_values, "t_values")
        R, S = t.shape
        color = torch.empty((R, 3), dtype=torch.float32, device=o.device)
        alpha = torch.empty((R,), dtype=torch.float32, device=o.device)
        depth = torch.empty((R,), dtype=torch.float32, device=o.device) if include_depth else None
        with on_device(self.device):
            _check(lib().ffn_render_rays_t(self.handle, _ptr(o), _ptr(d), _ptr(t), R, S, _ptr(color), _ptr(alpha),
                                           _ptr(depth), _ptr(self._nan_flag), _stream()), "ffn_render_rays_t")
        return color, alpha, depth

    def debug_stats(self):
        out = (ctypes.c_uint64 * 32)()
        lib().ffn_debug_stats.argtypes = [c_void_p, c_void_p]
        _check(lib().ffn_debug_stats(self.handle, out), "ffn_debug_stats")
        return list(out)

    def nan_flag(self) -> int:
        """Read (and clear) the device NaN flag -- one D2H sync."""
        v = int(self._nan_flag.item())
        if v:
            self._nan_flag.zero_()
        return v


def composite(raw: torch.Tensor, t_values: torch.Tensor, include_depth: bool = True,
              want_weights: bool = False, nan_flag: Optional[torch.Tensor] = None):
    """``ffn_composite``: raw (R,S,4), t (R,S) -> color, alpha, depth, weights."""
    raw = _f32c(raw, "raw")
    t = _f32c(t_values, "t_values")
    R, S = t.shape
    color = torch.empty((R, 3), dtype=torch.float32, device=t.device)
    alpha = torch.empty((R,), dtype=torch.float32, device=t.device)
    depth = torch.empty((R,), dtype=torch.float32, device=t.device) if include_depth else None
    weights = torch.empty((R, S), dtype=torch.float32, device=t.device) if want_weights else None
    with on_device(t.device):
        _check(lib().ffn_composite(_ptr(raw), _ptr(t), R, S, _ptr(color), _ptr(alpha), _ptr(depth),
                                   _ptr(weights), _ptr(nan_flag), _stream()), "ffn_composite")
    return color, alpha, depth, weights


def blend_weights(t_values: torch.Tensor, opacity: torch.Tensor) -> torch.Tensor:
    """``ffn_blend_weights``: t (R,S), sigma (R,S) -> weights (R,S)."""
    t = _f32c(t_values, "t_values")
    sg = _f32c(opacity, "opacity")
    R, S = t.shape
    w = torch.empty((R, S), dtype=torch.float32, device=t.device)
    with on_device(t.device):
        _check(lib().ffn_blend_weights(_ptr(t), _ptr(sg), R, S, _ptr(w), _stream()), "ffn_blend_weights")
    return w


def generate_rays(unproj: torch.Tensor, cam_pos: torch.Tensor, bounds_min, bounds_max, width: int, height: int):
    """``ffn_generate_rays``: per-camera inverse projections (C,4,4) and positions (C,3) on the device ->
    ``starts (N,3), directions (N,3), near_far (2,N), valid (N) bool`` for the N = C*H*W pixel rays."""
    u = _f32c(unproj.reshape(-1, 16), "unproj")
    pos = _f32c(cam_pos.reshape(-1, 3), "cam_pos")
    C = u.shape[0]
    n = C * width * height
    dev = u.device
    starts = torch.empty((n, 3), dtype=torch.float32, device=dev)
    directions = torch.empty((n, 3), dtype=torch.float32, device=dev)
    near_far = torch.empty((2, n), dtype=torch.float32, device=dev)
    valid = torch.empty((n,), dtype=torch.uint8, device=dev)
    lo = (ctypes.c_float * 3)(*[float(v) for v in bounds_min])
    hi = (ctypes.c_float * 3)(*[float(v) for v in bounds_max])
    with on_device(dev):
        _check(lib().ffn_generate_rays(_ptr(u), _ptr(pos), lo, hi, C, width, height, _ptr(starts), _ptr(directions),
                                       _ptr(near_far), _ptr(valid), _stream()), "ffn_generate_rays")
    return starts, directions, near_far, valid.bool()


def voxels_forward(grid_channels_last: torch.Tensor, bias4, scale: float, positions: torch.Tensor) -> torch.Tensor:
    """``ffn_voxels_forward``: (side,side,side,4) grid on the device, 4 bias floats, positions (N,3) -> (N,4)."""
    g = _f32c(grid_channels_last, "grid")
    p = _f32c(positions.reshape(-1, 3), "positions")
    n = p.shape[0]
    out = torch.empty((n, 4), dtype=torch.float32, device=p.device)
    b = (ctypes.c_float * 4)(*[float(v) for v in bias4])
    with on_device(p.device):
        _check(lib().ffn_voxels_forward(_ptr(g), b, g.shape[0], float(scale), _ptr(p), n, _ptr(out), _stream()),
               "ffn_voxels_forward")
    return out


def focus_t(raw_sigma: torch.Tensor, near, far, near_u, far_u, lin_c, lin_u, jitter_u, u_focus,
            stratified: bool, seed: int, num_samples: int, ray_offset: int = 0) -> torch.Tensor:
    """``ffn_focus_t`` on given coarse opacity logits (R, S_c) or raw outputs (R, S_c, 4)."""
    raw = _f32c(raw_sigma, "raw")
    stride = 4 if raw.dim() == 3 else 1
    R = raw.shape[0]
    t = torch.empty((R, num_samples), dtype=torch.float32, device=raw.device)
    ju = _f32c(jitter_u, "jitter_u") if jitter_u is not None else None
    uf = _f32c(u_focus, "u_focus") if u_focus is not None else None
    with on_device(raw.device):
        _check(lib().ffn_focus_t(_ptr(raw), stride, _ptr(_f32c(near, "near")), _ptr(_f32c(far, "far")),
                                 _ptr(_f32c(near_u, "near_u")), _ptr(_f32c(far_u, "far_u")), _ptr(_f32c(lin_c, "lin_c")),
                                 _ptr(_f32c(lin_u, "lin_u")), _ptr(_f32c(lin_c, "lin_c")), _ptr(ju), _ptr(uf),
                                 int(bool(stratified)), c_uint64(seed & (2**64 - 1)), ray_offset, R, num_samples, _ptr(t),
                                 _stream()), "ffn_focus_t")
    return t
