"""CPU oracle for the volume-rendering hot path of matajoh/fourier_feature_nets.

THIS IS TEST INFRASTRUCTURE.  It is a numpy restatement (float32 by default,
float64 on request as the accuracy arbiter) of the reference's algorithm for

    RaySampler.sample -> NeRF / FourierFeatureMLP forward -> Raycaster.render

Every function cites the reference ``file:line`` it follows (paths relative to
the reference checkout).  Nothing in the shipped package imports it.

Parity pinning: the reference ships no tests for this path; the
oracle is pinned against (1) golden vectors produced by importing and running
the *real* reference in the build container (``tests/golden/make_golden.py``,
fixtures committed under ``tests/golden/``) and (2) the one numeric fixture the
reference holds, ``docs/ray_data.tsv`` (transmittance trace, copied values in
``tests/golden/ray_data_kat.npz``).  See ``tests/test_oracle_golden.py``.
"""
from __future__ import annotations

import math
from typing import Dict, List, NamedTuple, Optional, Sequence

import numpy as np

__all__ = [
    "torch_linspace", "linspace", "calculate_blend_weights", "determine_cdf",
    "sample_t_values", "unproject", "raycast", "near_far", "sample_rays",
    "nerf_encoding_matrix", "nerf_forward", "nerf_hidden", "positional_b_values",
    "ffmlp_forward", "render", "render_rays", "softplus", "sigmoid",
    "OracleSamples", "OracleRender", "nerf_param_shapes", "init_nerf_params",
    "init_ffmlp_params", "image_loss",
]


# --------------------------------------------------------------------------
# small torch-semantics helpers
# --------------------------------------------------------------------------

def torch_linspace(start: float, end: float, steps: int, dtype=np.float32) -> np.ndarray:
    """``torch.linspace(start, end, steps)`` value-for-value.

    ATen computes ``step = (end-start)/(steps-1)`` in the output dtype and fills
    symmetrically: ``fma(step, i, start)`` for ``i < steps//2`` and
    ``fma(-step, steps-1-i, end)`` otherwise.  Used at utils.py:192,
    ray_sampler.py:314, nerf_model.py:79, fourier_feature_models.py:161.
    """
    dt = np.dtype(dtype).type
    if steps == 1:
        return np.array([start], dtype=dtype)
    s, e = dt(start), dt(end)
    step = dt((e - s) / dt(steps - 1))
    idx = np.arange(steps)
    half = steps // 2
    # ATen's vectorised CPU kernel evaluates start + step*i with a fused
    # multiply-add (one rounding): reproduce by forming the exact product in
    # float64 and rounding once (verified bit-exact vs torch for n in 7..128).
    lo = (np.float64(s) + np.float64(step) * idx).astype(dtype)
    hi = (np.float64(e) - np.float64(step) * (steps - 1 - idx)).astype(dtype)
    return np.where(idx < half, lo, hi).astype(dtype)


def linspace(start: np.ndarray, stop: np.ndarray, num_samples: int) -> np.ndarray:
    """utils.py:179-194 -- ``start[:,None] + linspace(0,1,S)[None,:] * (stop-start)[:,None]``."""
    dtype = start.dtype
    diff = (stop - start).astype(dtype)
    samples = torch_linspace(0.0, 1.0, num_samples, dtype)
    return (start[:, None] + (samples[None, :] * diff[:, None]).astype(dtype)).astype(dtype)


def softplus(x: np.ndarray) -> np.ndarray:
    """``F.softplus`` (beta=1, threshold=20) as called at ray_caster.py:71."""
    with np.errstate(over="ignore"):
        soft = np.log1p(np.exp(x))
    return np.where(x > 20, x, soft).astype(x.dtype)


def sigmoid(x: np.ndarray) -> np.ndarray:
    """``torch.sigmoid`` as called at ray_caster.py:70."""
    with np.errstate(over="ignore"):
        return (1 / (1 + np.exp(-x))).astype(x.dtype)


# --------------------------------------------------------------------------
# compositing
# --------------------------------------------------------------------------

def calculate_blend_weights(t_values: np.ndarray, opacity: np.ndarray) -> np.ndarray:
    """utils.py:72-97.

    delta_i = t_{i+1}-t_i, delta_last = 1e10; alpha = 1-exp(-sigma*delta);
    T = exclusive cumprod(min(1, 1-alpha+1e-10)); w = alpha*T.
    """
    dtype = t_values.dtype
    deltas = t_values[:, 1:] - t_values[:, :-1]
    max_dist = np.full_like(deltas[:, :1], 1e10)
    deltas = np.concatenate([deltas, max_dist], axis=-1)
    with np.errstate(over="ignore"):
        alpha = (1 - np.exp(-(opacity * deltas))).astype(dtype)
    trans = np.minimum(np.ones_like(alpha), (1 - alpha + dtype.type(1e-10)).astype(dtype))
    trans = trans[:, :-1]
    trans = np.concatenate([np.ones_like(trans[:, :1]), trans], axis=-1)
    trans = np.cumprod(trans, axis=-1, dtype=dtype)
    return (alpha * trans).astype(dtype)


def determine_cdf(t_values: np.ndarray, opacity: np.ndarray) -> np.ndarray:
    """ray_sampler.py:59-67."""
    dtype = t_values.dtype
    weights = calculate_blend_weights(t_values, opacity)
    weights = weights[:, 1:-1] + dtype.type(1e-5)
    cdf = np.cumsum(weights, axis=-1, dtype=dtype)
    cdf = (cdf / cdf[:, -1:]).astype(dtype)
    return np.concatenate([np.zeros_like(cdf[:, :1]), cdf], axis=-1)


def sample_t_values(near: np.ndarray, far: np.ndarray, cdf: np.ndarray,
                    num_samples: int, u: Optional[np.ndarray]) -> np.ndarray:
    """ray_sampler.py:301-357 (inverse-transform sampling of the coarse CDF).

    ``u`` is the (R, num_samples) uniform draw the reference takes from
    ``torch.rand`` when stratified (:313); ``None`` = the deterministic
    ``linspace(0,1,num_samples)`` branch (:315-316).
    """
    dtype = near.dtype
    num_rays = len(near)
    t_values = linspace(near, far, num_samples)
    t_values = (dtype.type(0.5) * (t_values[..., :-1] + t_values[..., 1:])).astype(dtype)
    if u is None:
        samples = np.repeat(torch_linspace(0.0, 1.0, num_samples, dtype)[None, :], num_rays, 0)
    else:
        samples = u.astype(dtype)
    # torch.searchsorted(cdf, samples, right=True), row-wise
    index = np.stack([np.searchsorted(cdf[r], samples[r], side="right")
                      for r in range(num_rays)]).astype(np.int64)
    i = np.maximum(0, index - 1)
    j = np.minimum(cdf.shape[-1] - 1, index)
    cdf_i = np.take_along_axis(cdf, i, 1)
    cdf_j = np.take_along_axis(cdf, j, 1)
    t_i = np.take_along_axis(t_values, i, 1)
    t_j = np.take_along_axis(t_values, j, 1)
    denominator = cdf_j - cdf_i
    denominator = np.where(denominator < 1e-5, np.ones_like(denominator), denominator)
    t_diff = ((samples - cdf_i) / denominator).astype(dtype)
    t_scale = t_j - t_i
    return (t_i + t_diff * t_scale).astype(dtype)


# --------------------------------------------------------------------------
# ray generation
# --------------------------------------------------------------------------

def unproject(intrinsics: np.ndarray, extrinsics: np.ndarray, points: np.ndarray) -> np.ndarray:
    """camera_info.py:66-74."""
    projection = np.eye(4, dtype=np.float32)
    projection[:3, :3] = intrinsics
    projection = projection @ np.linalg.inv(extrinsics)
    unprojection = np.linalg.inv(projection)
    h_coords = points.reshape(-1, 2)
    h_coords = np.concatenate([h_coords, np.ones((h_coords.shape[0], 2), np.float32)], axis=-1)
    return (unprojection @ h_coords.T).T


def raycast(intrinsics: np.ndarray, extrinsics: np.ndarray, points: np.ndarray):
    """camera_info.py:99-109 -> (origins (N,3), unit directions (N,3))."""
    points = points.astype(np.float32)
    world_coords = unproject(intrinsics, extrinsics, points)
    camera_pos = extrinsics[:3, 3].reshape(1, 3)
    ray_dir = world_coords[:, :3] - camera_pos
    ray_dir = ray_dir / np.linalg.norm(ray_dir, axis=-1, keepdims=True)
    return camera_pos + 0 * ray_dir, ray_dir


def near_far(bounds: np.ndarray, starts: np.ndarray, directions: np.ndarray):
    """ray_sampler.py:98-102,202-232 -> (near_far (2,N), valid (N,) bool)."""
    bounds_min = (bounds @ np.array([-0.5, -0.5, -0.5, 1], np.float32))[np.newaxis, :3]
    bounds_max = (bounds @ np.array([0.5, 0.5, 0.5, 1], np.float32))[np.newaxis, :3]
    with np.errstate(divide="ignore", invalid="ignore"):
        test0 = (bounds_min - starts) / directions
        test1 = (bounds_max - starts) / directions
    near = np.where(test0 < test1, test0, test1)
    far = np.where(test0 > test1, test0, test1)
    near = near.max(-1)
    far = far.min(-1)
    valid = near < far
    near[valid] = np.maximum(0.1, near[valid])
    return np.stack([near, far]), valid


class OracleSamples(NamedTuple):
    positions: np.ndarray        # (R,S,3)
    view_directions: np.ndarray  # (R,S,3)
    t_values: np.ndarray         # (R,S)


def sample_rays(starts: np.ndarray, directions: np.ndarray, near: np.ndarray, far: np.ndarray,
                num_samples: int, u: Optional[np.ndarray] = None,
                step: Optional[int] = None, num_anneal_steps: int = 0,
                anneal_start: float = 0.5,
                cdf: Optional[np.ndarray] = None, u_focus: Optional[np.ndarray] = None,
                focus_stratified: bool = False) -> OracleSamples:
    """ray_sampler.py:359-403.

    ``u`` is the (R,S') stratified jitter (``torch.rand`` at :383); ``None`` =
    not stratified.  With ``cdf`` given (focus sampling) S' = S//2 uniform
    samples are merged and sorted with S-S//2 CDF samples (:388-392).
    """
    dtype = starts.dtype
    num_rays = len(starts)
    ns = num_samples // 2 if cdf is not None else num_samples
    near = near.astype(dtype)
    far = far.astype(dtype)
    focus_near, focus_far = near, far          # _sample_t_values re-reads near_far (:305)
    if step is not None and step < num_anneal_steps:
        progress = step / num_anneal_steps
        anneal = dtype.type(min(max(progress, anneal_start), 1))
        midpoint = ((near + far) * dtype.type(0.5)).astype(dtype)
        near = (midpoint + (near - midpoint) * anneal).astype(dtype)
        far = (midpoint + (far - midpoint) * anneal).astype(dtype)
    t_values = linspace(near, far, ns)
    if u is not None:
        scale = ((far - near) / dtype.type(ns)).astype(dtype)
        permute = (u.astype(dtype) * scale[:, None]).astype(dtype)
        t_values = (t_values + permute).astype(dtype)
    if cdf is not None:
        nf = num_samples - ns
        focus = sample_t_values(focus_near, focus_far, cdf, nf,
                                u_focus if focus_stratified else None)
        t_values = np.sort(np.concatenate([t_values, focus], -1), axis=-1)
    dirs = np.repeat(directions.reshape(num_rays, 1, 3), num_samples, 1)
    positions = (starts.reshape(num_rays, 1, 3) + (t_values[..., None] * dirs).astype(dtype)).astype(dtype)
    return OracleSamples(positions, dirs, t_values)


# --------------------------------------------------------------------------
# models
# --------------------------------------------------------------------------

def nerf_encoding_matrix(max_log_scale: float, num_freq: int, num_inputs: int = 3,
                         dtype=np.float32) -> np.ndarray:
    """nerf_model.py:77-84 (== fourier_feature_models.py:157-166 with
    ``embedding_size//num_inputs`` frequencies).  Returns (num_inputs, F*num_inputs)
    with column ``F_k*num_inputs + j`` = ``2**lin_k`` on row j."""
    lin = torch_linspace(0.0, max_log_scale, num_freq, np.float32)
    freqs = np.power(np.float32(2.0), lin).astype(np.float32)
    mat = np.zeros((num_inputs, num_freq * num_inputs), dtype=dtype)
    for k in range(num_freq):
        for j in range(num_inputs):
            mat[j, k * num_inputs + j] = freqs[k]
    return mat


def positional_b_values(max_log_scale: float, embedding_size: int, num_inputs: int) -> np.ndarray:
    """fourier_feature_models.py:157-166."""
    return nerf_encoding_matrix(max_log_scale, embedding_size // num_inputs, num_inputs)


def _linear(x: np.ndarray, w: np.ndarray, b: np.ndarray) -> np.ndarray:
    """``nn.Linear``: y = x @ W.T + b, W stored (out, in)."""
    return (x @ w.T.astype(x.dtype) + b.astype(x.dtype)).astype(x.dtype)


def nerf_param_shapes(num_layers=8, num_channels=256, num_freq_pos=10, num_freq_view=4,
                      skips=(4,), include_inputs=True) -> Dict[str, tuple]:
    """Parameter names/shapes of ``NeRF`` (nerf_model.py:45-75)."""
    n_in = 2 * 3 * num_freq_pos + (3 if include_inputs else 0)
    shapes = {}
    li = n_in
    for i in range(num_layers):
        if i in set(skips):
            li += n_in
        shapes[f"layers.{i}.weight"] = (num_channels, li)
        shapes[f"layers.{i}.bias"] = (num_channels,)
        li = num_channels
    shapes["opacity_out.weight"] = (1, li)
    shapes["opacity_out.bias"] = (1,)
    shapes["bottleneck.weight"] = (num_channels, li)
    shapes["bottleneck.bias"] = (num_channels,)
    lv = num_channels + 2 * 3 * num_freq_view + (3 if include_inputs else 0)
    shapes["hidden_view.weight"] = (num_channels // 2, lv)
    shapes["hidden_view.bias"] = (num_channels // 2,)
    shapes["color_out.weight"] = (3, num_channels // 2)
    shapes["color_out.bias"] = (3,)
    return shapes


def _init_linear(rng: np.random.Generator, out_f: int, in_f: int, gain: float = 1.0):
    """Same *distribution* as ``nn.Linear.reset_parameters`` (U(-1/sqrt(in), 1/sqrt(in)));
    not the same stream as torch -- tests that need reference weights load them
    from the golden fixtures."""
    bound = gain / math.sqrt(in_f)
    w = rng.uniform(-bound, bound, size=(out_f, in_f)).astype(np.float32)
    b = rng.uniform(-bound, bound, size=(out_f,)).astype(np.float32)
    return w, b


def init_nerf_params(seed: int = 0, gain: float = 1.0, **kw) -> Dict[str, np.ndarray]:
    rng = np.random.default_rng(seed)
    params = {}
    shapes = nerf_param_shapes(**kw)
    for name, shp in shapes.items():
        if name.endswith(".weight"):
            w, b = _init_linear(rng, shp[0], shp[1], gain)
            params[name] = w
            params[name[:-6] + "bias"] = b
    return params


def init_ffmlp_params(seed: int, num_inputs_enc: int, num_outputs: int,
                      layer_channels: Sequence[int], gain: float = 1.0) -> Dict[str, np.ndarray]:
    rng = np.random.default_rng(seed)
    params = {}
    n_in = num_inputs_enc
    chans = list(layer_channels) + [num_outputs]
    for i, c in enumerate(chans):
        w, b = _init_linear(rng, c, n_in, gain)
        params[f"layers.{i}.weight"] = w
        params[f"layers.{i}.bias"] = b
        n_in = c
    return params


def nerf_forward(params: Dict[str, np.ndarray], position: np.ndarray, view: np.ndarray,
                 num_layers=8, max_log_scale_pos=9.0, num_freq_pos=10,
                 max_log_scale_view=3.0, num_freq_view=4, skips=(4,),
                 include_inputs=True) -> np.ndarray:
    """nerf_model.py:86-124 -> (N,4) = [rgb_raw(3) | sigma_raw(1)]."""
    dtype = position.dtype
    skips = set(skips)
    pos_enc = nerf_encoding_matrix(max_log_scale_pos, num_freq_pos, 3, dtype)
    view_enc = nerf_encoding_matrix(max_log_scale_view, num_freq_view, 3, dtype)
    e = (position @ pos_enc).astype(dtype)
    enc_p = [np.cos(e), np.sin(e)]
    if include_inputs:
        enc_p.append(position)
    enc_p = np.concatenate(enc_p, -1).astype(dtype)
    e = (view @ view_enc).astype(dtype)
    enc_v = [np.cos(e), np.sin(e)]
    if include_inputs:
        enc_v.append(view)
    enc_v = np.concatenate(enc_v, -1).astype(dtype)
    out = enc_p
    for i in range(num_layers):
        if i in skips:
            out = np.concatenate([out, enc_p], -1)
        out = np.maximum(_linear(out, params[f"layers.{i}.weight"], params[f"layers.{i}.bias"]), 0)
    opacity = _linear(out, params["opacity_out.weight"], params["opacity_out.bias"])
    bottleneck = _linear(out, params["bottleneck.weight"], params["bottleneck.bias"])
    out = np.concatenate([bottleneck, enc_v], -1)
    out = np.maximum(_linear(out, params["hidden_view.weight"], params["hidden_view.bias"]), 0)
    color = _linear(out, params["color_out.weight"], params["color_out.bias"])
    return np.concatenate([color, opacity], -1).astype(dtype)


def nerf_hidden(params: Dict[str, np.ndarray], position: np.ndarray, view: np.ndarray,
                num_layers=8, max_log_scale_pos=9.0, num_freq_pos=10,
                max_log_scale_view=3.0, num_freq_view=4, skips=(4,),
                include_inputs=True) -> List[np.ndarray]:
    """Post-activation outputs of every dense layer of nerf_model.py:111-122 in order:
    layers.0..L-1 (ReLU), bottleneck (linear), hidden_view (ReLU).  Bring-up aid."""
    dtype = position.dtype
    skips = set(skips)
    pos_enc = nerf_encoding_matrix(max_log_scale_pos, num_freq_pos, 3, dtype)
    view_enc = nerf_encoding_matrix(max_log_scale_view, num_freq_view, 3, dtype)
    e = (position @ pos_enc).astype(dtype)
    enc_p = np.concatenate([np.cos(e), np.sin(e)] + ([position] if include_inputs else []), -1)
    e = (view @ view_enc).astype(dtype)
    enc_v = np.concatenate([np.cos(e), np.sin(e)] + ([view] if include_inputs else []), -1)
    outs = []
    out = enc_p
    for i in range(num_layers):
        if i in skips:
            out = np.concatenate([out, enc_p], -1)
        out = np.maximum(_linear(out, params[f"layers.{i}.weight"], params[f"layers.{i}.bias"]), 0)
        outs.append(out)
    b = _linear(out, params["bottleneck.weight"], params["bottleneck.bias"])
    outs.append(b)
    out = np.maximum(_linear(np.concatenate([b, enc_v], -1), params["hidden_view.weight"],
                             params["hidden_view.bias"]), 0)
    outs.append(out)
    return outs


def ffmlp_forward(params: Dict[str, np.ndarray], inputs: np.ndarray,
                  a_values: Optional[np.ndarray], b_values: Optional[np.ndarray]) -> np.ndarray:
    """fourier_feature_models.py:57-78: ``[a cos(pi x B) | a sin(pi x B)]`` ->
    ReLU MLP -> final Linear (no activation)."""
    dtype = inputs.dtype
    if b_values is None:
        out = inputs
    else:
        enc = ((dtype.type(math.pi) * inputs).astype(dtype) @ b_values.astype(dtype)).astype(dtype)
        a = a_values.astype(dtype)
        out = np.concatenate([a * np.cos(enc), a * np.sin(enc)], -1).astype(dtype)
    n_layers = len([k for k in params if k.endswith(".weight")])
    for i in range(n_layers - 1):
        out = np.maximum(_linear(out, params[f"layers.{i}.weight"], params[f"layers.{i}.bias"]), 0)
    return _linear(out, params[f"layers.{n_layers-1}.weight"], params[f"layers.{n_layers-1}.bias"])


# --------------------------------------------------------------------------
# render
# --------------------------------------------------------------------------

class OracleRender(NamedTuple):
    color: np.ndarray   # (R,3)
    alpha: np.ndarray   # (R,)
    depth: Optional[np.ndarray]  # (R,)
    weights: np.ndarray  # (R,S) blend weights (for diagnostics)


def render(color_o: np.ndarray, t_values: np.ndarray, include_depth: bool = True) -> OracleRender:
    """ray_caster.py:67-93 given the model output ``color_o`` (R,S,4)."""
    dtype = t_values.dtype
    color = sigmoid(color_o[..., :3])
    opacity = softplus(color_o[..., 3])
    assert not np.isnan(color).any()
    assert not np.isnan(opacity).any()
    weights = calculate_blend_weights(t_values, opacity)
    out_color = (weights[..., None] * color).sum(-2, dtype=dtype)
    w1 = weights[:, :-1]
    out_alpha = w1.sum(-1, dtype=dtype)
    depth = None
    if include_depth:
        cutoff = w1.argmax(-1)
        cutoff[out_alpha < .1] = -1
        depth = t_values[np.arange(len(t_values)), cutoff]
    return OracleRender(out_color.astype(dtype), out_alpha.astype(dtype), depth, weights)


def render_rays(model_fn, samples: OracleSamples, include_depth: bool = True,
                use_view: bool = True, chunk: int = 8192) -> OracleRender:
    """ray_caster.py:48-93: flatten -> model -> reshape -> ``render``."""
    R, S = samples.positions.shape[:2]
    pos = samples.positions.reshape(-1, 3)
    views = samples.view_directions.reshape(-1, 3)
    # the model is evaluated in row chunks that fit the CPU caches (the result does not
    # depend on the chunking: every row is independent)
    outs = []
    for lo in range(0, len(pos), chunk):
        if use_view:
            outs.append(model_fn(pos[lo:lo + chunk], views[lo:lo + chunk]))
        else:
            outs.append(model_fn(pos[lo:lo + chunk]))
    color_o = np.concatenate(outs, 0)
    return render(color_o.reshape(R, S, 4), samples.t_values, include_depth)


def image_loss(color: np.ndarray, alpha: np.ndarray, gt_color: np.ndarray,
               gt_alpha: Optional[np.ndarray]) -> float:
    """image_dataset.py:224-242: MSE(color) + 0.1*MSE(alpha) (alpha term only
    when the dataset carries alpha)."""
    loss = np.mean((color - gt_color) ** 2, dtype=np.float64)
    if gt_alpha is not None:
        loss = loss + 0.1 * np.mean((alpha - gt_alpha) ** 2, dtype=np.float64)
    return float(loss)
