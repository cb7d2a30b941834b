"""CPU oracle, torch flavour: the reference's op sequence for  RaySampler.sample -> NeRF.forward -> Raycaster.render
executed with the SAME ATen CPU kernels the reference itself runs (torch.mm / cos / sin / cat / relu / sigmoid /
softplus / cumprod ...), multi-threaded.

THIS IS TEST / BASELINE INFRASTRUCTURE (see oracle/__init__.py).  It exists because the reference is a pure-Python
package that cannot travel to the GPU box: ``bench.py --impl reference`` and ``cpu_baseline`` time THIS (it is ~1.5x
faster than the numpy restatement on the same cores, i.e. the fairer stand-in for the reference's CPU path); parity is
still judged against ``ffn_oracle.py``.  Pinned to the real reference by ``tests/test_oracle_golden.py``
(``test_torch_oracle_*``: bit-identical pixels to the reference's recorded fp32 outputs where the op order is the same).
Every function cites the reference ``file:line`` it follows.
"""
from __future__ import annotations

from typing import Dict, NamedTuple, Optional

import torch
import torch.nn.functional as F

TorchRender = NamedTuple("TorchRender", [("color", torch.Tensor), ("alpha", torch.Tensor),
                                         ("depth", Optional[torch.Tensor])])


def encoding_matrix(max_log_scale: float, num_freq: int) -> torch.Tensor:
    """nerf_model.py:77-84: (3, 3F) with P[j, 3k + j] = 2 ** linspace(0, max_log_scale, F)[k]."""
    freqs = torch.pow(2., torch.linspace(0., max_log_scale, num_freq))
    mat = torch.zeros((3, 3 * num_freq), dtype=torch.float32)
    for k in range(num_freq):
        for j in range(3):
            mat[j, 3 * k + j] = freqs[k]
    return mat


def sample_rays(starts, directions, near, far, num_samples: int, u: Optional[torch.Tensor]):
    """ray_sampler.py:380-397 (uniform / stratified, no annealing, no focus sampling)."""
    lin = torch.linspace(0, 1, num_samples)                                        # utils.py:192
    t = near.unsqueeze(-1) + lin.unsqueeze(0) * (far - near).unsqueeze(-1)         # utils.py:193-194
    if u is not None:
        t = t + u * ((far - near) / num_samples).unsqueeze(-1)                      # :381-386
    n = len(starts)
    dirs = directions.reshape(n, 1, 3).repeat(1, num_samples, 1)                    # :396
    positions = starts.reshape(n, 1, 3) + t.unsqueeze(-1) * dirs                    # :397
    return positions, dirs, t


def nerf_forward(p: Dict[str, torch.Tensor], position, view, num_layers=8, skips=(4,), include_inputs=True,
                 pos_enc=None, view_enc=None) -> torch.Tensor:
    """nerf_model.py:86-124 -> (N,4) = [rgb_raw | sigma_raw]."""
    def encode(x, mat):                                                             # :97-109
        e = torch.mm(x, mat)
        parts = [e.cos(), e.sin()]
        if include_inputs:
            parts.append(x)
        return torch.cat(parts, -1)
    enc_p, enc_v = encode(position, pos_enc), encode(view, view_enc)
    out = enc_p
    for i in range(num_layers):                                                     # :111-116
        if i in skips:
            out = torch.cat([out, enc_p], -1)
        out = torch.relu(F.linear(out, p["layers.%d.weight" % i], p["layers.%d.bias" % i]))
    opacity = F.linear(out, p["opacity_out.weight"], p["opacity_out.bias"])         # :118
    bottleneck = F.linear(out, p["bottleneck.weight"], p["bottleneck.bias"])        # :119
    out = torch.relu(F.linear(torch.cat([bottleneck, enc_v], -1), p["hidden_view.weight"], p["hidden_view.bias"]))
    color = F.linear(out, p["color_out.weight"], p["color_out.bias"])               # :123
    return torch.cat([color, opacity], -1)


def blend_weights(t_values, opacity) -> torch.Tensor:
    """utils.py:72-97."""
    delta = t_values[:, 1:] - t_values[:, :-1]
    delta = torch.cat([delta, torch.full_like(delta[:, :1], 1e10)], -1)             # :86
    alpha = 1 - torch.exp(-(opacity * delta))
    ones = torch.ones_like(alpha[:, :1])
    trans = torch.minimum(ones, 1 - alpha + 1e-10)                                  # :92
    trans = torch.cat([ones, trans[:, :-1]], -1).cumprod(-1)
    return alpha * trans


def render(raw, t_values, include_depth=True) -> TorchRender:
    """ray_caster.py:67-93 given the model output raw (R,S,4)."""
    color = torch.sigmoid(raw[..., :3])
    opacity = F.softplus(raw[..., 3])
    assert not color.isnan().any()
    assert not opacity.isnan().any()
    w = blend_weights(t_values, opacity)
    out_color = (w.unsqueeze(-1) * color).sum(-2)
    w = w[:, :-1]
    out_alpha = w.sum(-1)
    depth = None
    if include_depth:
        cutoff = w.argmax(-1)
        cutoff[out_alpha < .1] = -1
        depth = t_values[torch.arange(len(t_values)), cutoff]
    return TorchRender(out_color, out_alpha, depth)


@torch.no_grad()
def render_rays(params: Dict[str, torch.Tensor], starts, directions, near, far, num_samples: int,
                u: Optional[torch.Tensor], include_depth=True, batch: int = 4096) -> TorchRender:
    """RaySampler.sample + Raycaster.batched_render (ray_caster.py:103-138) in ray batches of ``batch`` (the reference's
    inference batch, orbit_video.py:37).  All inputs CPU float32 tensors; NeRF(8,256,9,10,3,4,[4],True) layout."""
    pos_enc, view_enc = encoding_matrix(9.0, 10), encoding_matrix(3.0, 4)
    outs = []
    for lo in range(0, len(starts), batch):
        sl = slice(lo, lo + batch)
        positions, dirs, t = sample_rays(starts[sl], directions[sl], near[sl], far[sl], num_samples,
                                         None if u is None else u[sl])
        raw = nerf_forward(params, positions.reshape(-1, 3), dirs.reshape(-1, 3), pos_enc=pos_enc, view_enc=view_enc)
        outs.append(render(raw.reshape(len(t), num_samples, 4), t, include_depth))
    return TorchRender(torch.cat([o.color for o in outs]), torch.cat([o.alpha for o in outs]),
                       torch.cat([o.depth for o in outs]) if include_depth else None)
