"""Binds an ``nn.Module`` (NeRF / FourierFeatureMLP) to its packed tensor-core image.

The packed image is rebuilt on the GPU (``ffn_net_pack``) whenever a parameter
changed -- detected through the tensors' in-place version counters, so an
optimiser step costs two small kernel launches and no host synchronisation.
"""
from __future__ import annotations

import os
from typing import List, Optional, Tuple

import torch

from . import _lib

_OPERAND = {"fp16": _lib.OPERAND_FP16, "bf16": _lib.OPERAND_BF16, "fp16x3": _lib.OPERAND_FP16X3}
DEFAULT_OPERAND = "fp16"


def _linear_list(model) -> List[torch.nn.Linear]:
    kind = getattr(model, "_ffn_kind", None)
    if kind == "nerf":
        return list(model.layers) + [model.opacity_out, model.bottleneck, model.hidden_view,
                                     model.color_out]
    if kind == "fourier":
        return list(model.layers)
    raise _lib.FFNError("model type %s has no libffn_b200 engine" % type(model).__name__)


def unsupported_reason(model) -> Optional[str]:
    """``None`` if the fused sm_100a kernels cover this model (kind AND shape), else why not.

    The kernels are specialised for the BASELINE configurations: 256-wide trunks (one 128x256 fp32 TMEM
    accumulator per tile), <= 10 frequencies per encoding (a 64-wide encoding chunk), 3-D inputs, 4 outputs."""
    kind = getattr(model, "_ffn_kind", None)
    if kind == "nerf":
        p = model.params
        if p["num_channels"] != 256:
            return "num_channels = %d (the kernels are built for 256)" % p["num_channels"]
        if not 1 <= p["num_layers"] <= 10:
            return "num_layers = %d (1..10)" % p["num_layers"]
        if not (0 <= p["num_freq_pos"] <= _lib.FFN_MAX_FREQS and 0 <= p["num_freq_view"] <= _lib.FFN_MAX_FREQS):
            return "more than %d frequencies per encoding" % _lib.FFN_MAX_FREQS
        if 0 in set(p["skips"]):
            return "a skip connection into layer 0"
        if 6 * p["num_freq_pos"] + (3 if p["include_inputs"] else 0) < 1:
            return "empty positional encoding"
        return None
    if kind == "fourier":
        if model.keep_activations:
            return "keep_activations is set (the fused kernels keep hidden activations on chip)"
        if not model._engine_ok():
            return "only 3 -> [256] * n -> 4 FourierFeatureMLPs with embedding size <= 256 are covered"
        return None
    return "model type %s has no libffn_b200 engine" % type(model).__name__


def supported(model) -> bool:
    """Kind and shape are covered by the fused kernels."""
    return unsupported_reason(model) is None


def note_unfused(model):
    """A NeRF / FourierFeatureMLP on a CUDA device whose SHAPE the fused kernels do not cover is evaluated by its
    plain PyTorch definition (the reference supports arbitrary widths; compositing still runs in ``ffn_composite``).
    Never silently: a ``UserWarning`` once per model, or an ``FFNError`` when ``FFN_STRICT=1``."""
    if getattr(model, "_ffn_kind", None) not in ("nerf", "fourier") or model.__dict__.get("_ffn_unfused_noted"):
        return
    why = unsupported_reason(model)
    if why is None or why.startswith("keep_activations"):
        return
    msg = ("libffn_b200: %s is outside the fused sm_100a kernels (%s); its layers run as plain PyTorch ops"
           % (type(model).__name__, why))
    if os.environ.get("FFN_STRICT", "0") not in ("", "0"):
        raise _lib.FFNError(msg + " and FFN_STRICT is set")
    import warnings
    warnings.warn(msg, UserWarning, stacklevel=3)
    model.__dict__["_ffn_unfused_noted"] = True


def _encoding_signature(model) -> Tuple:
    """(storage pointer, version) of the frozen encoding buffers: they are baked into the C handle at creation, so
    ``load_state_dict`` of a checkpoint with another B matrix / other frequencies must rebuild the handle."""
    if model._ffn_kind == "nerf":
        bufs = (model.pos_encoding, model.view_encoding)
    else:
        bufs = (model.a_values, model.b_values)
    return tuple((None, None) if b is None else (b.data_ptr(), b._version) for b in bufs)


class Engine:
    def __init__(self, model, device: torch.device, operand: str):
        self.device = device
        self.operand = operand
        self.enc_sig = _encoding_signature(model)
        kind = model._ffn_kind
        if kind == "nerf":
            p = model.params
            # pos_encoding[j, 3k+j] = f_k  (nerf_model.py:77-84)
            freq_pos = [float(model.pos_encoding[0, 3 * k]) for k in range(p["num_freq_pos"])]
            freq_view = [float(model.view_encoding[0, 3 * k]) for k in range(p["num_freq_view"])]
            self.net = _lib.Net.nerf(p["num_layers"], p["num_channels"], freq_pos, freq_view,
                                     p["skips"], p["include_inputs"], device, _OPERAND[operand])
        else:
            hidden = len(model.layers) - 1
            chans = {l.out_features for l in list(model.layers)[:-1]}
            if chans != {256} or model.layers[-1].out_features != 4 or model.num_inputs != 3:
                raise _lib.FFNError("libffn_b200 supports 3 -> [256]*n -> 4 FourierFeatureMLPs")
            self.net = _lib.Net.ffmlp(hidden, 256, model.a_values, model.b_values, device,
                                      _OPERAND[operand])
        self._sig: Optional[Tuple] = None

    def sync_weights(self, model):
        """Re-pack when a parameter changed: storage pointer, tensor version counter, or -- because fused /
        capturable optimizers update parameters without bumping the version counters -- any optimizer step
        anywhere since the last pack."""
        lins = _linear_list(model)
        sig = (_OPT_GENERATION[0],) + tuple(
            (l.weight.data_ptr(), l.weight._version, l.bias.data_ptr(), l.bias._version) for l in lins)
        if sig != self._sig:
            self.net.pack([l.weight for l in lins], [l.bias for l in lins])
            self._sig = sig


_OPT_GENERATION = [0]


def _note_optimizer_step(*_args, **_kwargs):
    _OPT_GENERATION[0] += 1


try:        # global hook: runs after every torch.optim.Optimizer.step()
    from torch.optim.optimizer import register_optimizer_step_post_hook
    register_optimizer_step_post_hook(_note_optimizer_step)
except ImportError:      # very old torch: version counters only
    pass


def mark_weights_changed():
    """Call after changing parameters through a path that neither bumps tensor versions nor is an optimizer step
    (e.g. a custom fused update kernel)."""
    _OPT_GENERATION[0] += 1


def default_operand(model) -> str:
    """``model.ffn_operand`` if set, else the ``FFN_OPERAND`` environment variable (lets the reference's unchanged
    scripts select a mode), else fp16.  "fp16x3" = the precise inference mode (NeRF handles)."""
    return getattr(model, "ffn_operand", None) or os.environ.get("FFN_OPERAND") or DEFAULT_OPERAND


def coarse_operand(model) -> str:
    """Operand mode of the coarse opacity pass of hierarchical sampling (``model.ffn_coarse_operand`` /
    ``FFN_COARSE_OPERAND``; default: the model's own mode).  Inverse-transform sampling amplifies differences of the
    coarse opacities (ray_sampler.py:325-355), so "fp16x3" here makes the sample positions -- and with them the frames --
    follow the reference's fp32 path closely at ~1.8x the frame time."""
    return getattr(model, "ffn_coarse_operand", None) or os.environ.get("FFN_COARSE_OPERAND") or default_operand(model)


def get_engine(model, device: torch.device, operand: Optional[str] = None) -> Engine:
    """The engine (C handle + packed weights) of ``model`` for one operand mode; one per mode is kept."""
    operand = operand or default_operand(model)
    if operand == "fp16x3" and getattr(model, "_ffn_kind", None) != "nerf":
        operand = DEFAULT_OPERAND          # the precise mode covers NeRF handles
    engines = model.__dict__.setdefault("_ffn_engines", {})
    eng = engines.get(operand)
    if eng is None or eng.device != device or eng.enc_sig != _encoding_signature(model):
        eng = engines[operand] = Engine(model, device, operand)
    model.__dict__["_ffn_engine"] = eng      # the most recently used one (NaN flag checks, trainers)
    eng.sync_weights(model)
    return eng
