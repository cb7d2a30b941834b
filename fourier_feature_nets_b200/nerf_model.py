"""``NeRF`` with the reference's constructor, parameter names and ``.pt`` format
(reference: fourier_feature_nets/nerf_model.py:9-135), evaluated on B200 by the
fused tcgen05 kernel when called on CUDA tensors without autograd."""
from typing import Sequence

import torch
import torch.nn as nn

from . import engine as _engine


def positional_frequencies(max_log_scale: float, num_freq: int, num_inputs: int) -> torch.Tensor:
    """(num_inputs, num_freq*num_inputs) block-diagonal frequency matrix: column
    ``num_inputs*k + j`` carries ``2**linspace(0, max_log_scale, num_freq)[k]`` on row j
    (same buffer contents as nerf_model.py:77-84 so state dicts interchange)."""
    freqs = 2.0 ** torch.linspace(0, max_log_scale, num_freq)
    mat = torch.zeros(num_inputs, num_freq * num_inputs)
    for j in range(num_inputs):
        mat[j, j::num_inputs] = freqs
    return mat


class NeRF(nn.Module):
    """The full NeRF model (8x256 trunk, skip, sigma head, view branch)."""

    _ffn_kind = "nerf"

    def __init__(self, num_layers: int, num_channels: int,
                 max_log_scale_pos: float, num_freq_pos: int,
                 max_log_scale_view: float, num_freq_view: int,
                 skips: Sequence[int], include_inputs: bool):
        super().__init__()
        self.params = {
            "num_layers": num_layers,
            "num_channels": num_channels,
            "max_log_scale_pos": max_log_scale_pos,
            "num_freq_pos": num_freq_pos,
            "max_log_scale_view": max_log_scale_view,
            "num_freq_view": num_freq_view,
            "skips": list(skips),
            "include_inputs": include_inputs,
        }
        self.pos_encoding = nn.Parameter(
            positional_frequencies(max_log_scale_pos, num_freq_pos, 3), requires_grad=False)
        self.view_encoding = nn.Parameter(
            positional_frequencies(max_log_scale_view, num_freq_view, 3), requires_grad=False)
        self.skips = set(skips)
        self.include_inputs = include_inputs
        self.use_view = True

        enc_width = 2 * self.pos_encoding.shape[-1] + (3 if include_inputs else 0)
        self.layers = nn.ModuleList()
        width = enc_width
        for i in range(num_layers):
            if i in self.skips:
                width += enc_width
            self.layers.append(nn.Linear(width, num_channels))
            width = num_channels
        self.opacity_out = nn.Linear(width, 1)
        self.bottleneck = nn.Linear(width, num_channels)
        view_width = num_channels + 2 * self.view_encoding.shape[-1] + (3 if include_inputs else 0)
        self.hidden_view = nn.Linear(view_width, num_channels // 2)
        self.color_out = nn.Linear(num_channels // 2, 3)

    # -- plain differentiable definition (CPU, or CUDA under autograd) -----------------
    def _encode(self, x: torch.Tensor, matrix: torch.Tensor) -> torch.Tensor:
        e = x @ matrix
        parts = [e.cos(), e.sin()]
        if self.include_inputs:
            parts.append(x)
        return torch.cat(parts, dim=-1)

    def forward_torch(self, position: torch.Tensor, view: torch.Tensor) -> torch.Tensor:
        enc_p = self._encode(position, self.pos_encoding)
        enc_v = self._encode(view, self.view_encoding)
        h = enc_p
        for i, layer in enumerate(self.layers):
            if i in self.skips:
                h = torch.cat([h, enc_p], dim=-1)
            h = torch.relu(layer(h))
        sigma = self.opacity_out(h)
        h = torch.relu(self.hidden_view(torch.cat([self.bottleneck(h), enc_v], dim=-1)))
        return torch.cat([self.color_out(h), sigma], dim=-1)

    def forward(self, position: torch.Tensor, view: torch.Tensor) -> torch.Tensor:
        """(N,3) positions, (N,3) unit view directions -> (N,4) [rgb_raw | sigma_raw]."""
        if position.is_cuda and not _needs_grad(self, position, view) and _engine.supported(self):
            eng = _engine.get_engine(self, position.device)
            return eng.net.mlp_forward(position.reshape(-1, 3), view.reshape(-1, 3))
        return self.forward_torch(position, view)

    def __getstate__(self):
        """``copy.deepcopy`` / pickling: the C handles bound to this instance (engine, trainer, flat gradient buffer)
        stay behind; the copy builds its own on first use."""
        state = self.__dict__.copy()
        for key in [k for k in state if k.startswith("_ffn_")]:
            del state[key]
        return state

    def save(self, path: str):
        """Same ``.pt`` layout as nerf_model.py:126-135."""
        state_dict = self.state_dict()
        state_dict["type"] = "nerf"
        state_dict["params"] = self.params
        torch.save(state_dict, path)


def _needs_grad(model: nn.Module, *inputs) -> bool:
    if not torch.is_grad_enabled():
        return False
    if any(t is not None and t.requires_grad for t in inputs):
        return True
    return any(p.requires_grad for p in model.parameters())
