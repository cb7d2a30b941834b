"""Hot-path utilities with the reference's names (fourier_feature_nets/utils.py):
``calculate_blend_weights`` (:72), ``linspace`` (:179), ``exponential_lr_decay`` (:422),
``load_model`` (:448), ``RenderResult`` (:506), ``orbit`` (:244)."""
import math
import os
from typing import List, NamedTuple

import numpy as np
import torch

from . import _lib
from .camera_info import CameraInfo, Resolution


class RenderResult(NamedTuple("RenderResult", [("color", torch.Tensor),
                                               ("alpha", torch.Tensor),
                                               ("depth", torch.Tensor)])):
    """Per-ray colour, alpha and (optionally) depth."""

    @property
    def device(self) -> torch.device:
        return self.color.device

    def to(self, *args, **kwargs) -> "RenderResult":
        return RenderResult(*[None if t is None else t.to(*args, **kwargs) for t in self])

    def numpy(self) -> "RenderResult":
        return RenderResult(*[None if t is None else t.cpu().numpy() for t in self])


def blend_weights_torch(t_values: torch.Tensor, opacity: torch.Tensor) -> torch.Tensor:
    """Differentiable definition: delta_last = 1e10, alpha = 1 - exp(-sigma delta),
    T = exclusive cumprod(min(1, 1 - alpha + 1e-10)), w = alpha T   (utils.py:84-97)."""
    deltas = t_values[:, 1:] - t_values[:, :-1]
    deltas = torch.cat([deltas, torch.full_like(deltas[:, :1], 1e10)], dim=-1)
    alpha = 1 - torch.exp(-(opacity * deltas))
    trans = torch.minimum(torch.ones_like(alpha), 1 - alpha + 1e-10)[:, :-1]
    trans = torch.cat([torch.ones_like(trans[:, :1]), trans], dim=-1)
    return alpha * torch.cumprod(trans, -1)


def calculate_blend_weights(t_values: torch.Tensor, opacity: torch.Tensor) -> torch.Tensor:
    """(R,S) t values and opacities -> (R,S) blend weights."""
    if t_values.is_cuda and not (torch.is_grad_enabled() and (t_values.requires_grad or opacity.requires_grad)):
        return _lib.blend_weights(t_values, opacity)
    return blend_weights_torch(t_values, opacity)


def linspace(start: torch.Tensor, stop: torch.Tensor, num_samples: int) -> torch.Tensor:
    """Row-wise linspace: (D,) , (D,) -> (D, num_samples), end point included (utils.py:179-194)."""
    steps = torch.linspace(0, 1, num_samples, device=start.device)
    return start.unsqueeze(-1) + steps.unsqueeze(0) * (stop - start).unsqueeze(-1)


def exponential_lr_decay(optim: torch.optim.Optimizer, initial_learning_rate: float,
                         step: int, decay_rate: float, decay_steps: float):
    """lr = lr_0 * decay_rate ** (step / decay_steps) on every parameter group."""
    lr = initial_learning_rate * decay_rate ** (step / decay_steps)
    for group in optim.param_groups:
        group["lr"] = lr


def _safe_load(path: str):
    allow = [np.dtype, np.ndarray]
    for mod in ("numpy._core.multiarray", "numpy.core.multiarray"):
        try:
            m = __import__(mod, fromlist=["scalar"])
            allow += [m.scalar, m._reconstruct]
            break
        except (ImportError, AttributeError):
            continue
    allow += [type(np.dtype(t)) for t in (np.float32, np.float64, np.int32, np.int64, np.bool_)]
    with torch.serialization.safe_globals(allow):
        return torch.load(path, map_location="cpu", weights_only=True)


def load_model(path: str) -> torch.nn.Module:
    """Load a ``.pt`` written by ``model.save`` (reference format: state dict + "type" +
    "params", utils.py:448-503) and return it in eval mode."""
    from .fourier_feature_models import FourierFeatureMLP
    from .nerf_model import NeRF
    if not os.path.exists(path):
        alt = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "models", path))
        if not os.path.exists(alt):
            print("Unable to find model", path, "(no network: assets cannot be downloaded)")
            return None
        path = alt
    # tensors + a "type" string + a "params" dict of lists / numbers: loads under the safe unpickler.  Checkpoints
    # written by the reference may carry numpy scalars in "params" (train_voxels.py:99-101): allow exactly the numpy
    # scalar / dtype reconstructors, nothing else
    state = _safe_load(path)
    kind, params = state.pop("type"), state.pop("params")
    if kind == "fourier":
        for key in ("a_values", "b_values"):
            if params[key] is not None:
                params[key] = torch.FloatTensor(params[key])
        model = FourierFeatureMLP(**params)
    elif kind == "nerf":
        model = NeRF(**params)
    elif kind == "voxels":
        from .voxels_model import Voxels
        model = Voxels(**params)
    else:
        raise ValueError("Unrecognized model type: %s" % kind)
    model.load_state_dict(state)
    model.eval()
    return model


def _rotation(axis: np.ndarray, angle: float) -> np.ndarray:
    axis = np.asarray(axis, np.float64)
    axis = axis / np.linalg.norm(axis)
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    R = np.eye(3) + math.sin(angle) * K + (1 - math.cos(angle)) * (K @ K)
    out = np.eye(4)
    out[:3, :3] = R
    return out


def look_at_extrinsics(center: np.ndarray, up_dir: np.ndarray, target=(0, 0, 0)) -> np.ndarray:
    """Camera-to-world matrix, +z forward / +y down (OpenCV convention)."""
    center = np.asarray(center, np.float64)
    fwd = np.asarray(target, np.float64) - center
    fwd /= np.linalg.norm(fwd)
    right = np.cross(fwd, np.asarray(up_dir, np.float64))
    right /= np.linalg.norm(right)
    down = np.cross(fwd, right)
    ext = np.eye(4)
    ext[:3, 0], ext[:3, 1], ext[:3, 2], ext[:3, 3] = right, down, fwd, center
    return ext


def orbit(up_dir: np.ndarray, forward_dir: np.ndarray, num_frames: int,
          fov_y_degrees: float, resolution: Resolution, distance: float,
          min_altitude=np.pi / 12, max_altitude=np.pi / 4) -> List[CameraInfo]:
    """Two-turn orbit around the origin with an altitude sweep (utils.py:244-300).
    The reference builds the start pose with scenepic, which is absent here: orbit poses
    are "parity unpinned" (SURVEY.md section 8c); pixel parity is defined for given cameras."""
    up_dir = np.asarray(up_dir, np.float64)
    forward_dir = np.asarray(forward_dir, np.float64)
    right_dir = np.cross(up_dir, forward_dir)
    azimuth = np.linspace(0, 4 * np.pi, num_frames, endpoint=False)
    altitude = np.zeros_like(azimuth)
    half = num_frames // 2
    altitude[:half] = np.linspace(min_altitude, max_altitude, half, endpoint=False)
    altitude[half:] = np.linspace(max_altitude, min_altitude, num_frames - half, endpoint=False)
    focal = .5 * resolution.width / np.tan(.5 * fov_y_degrees * np.pi / 180)
    intrinsics = np.array([focal, 0, resolution.width / 2, 0, focal, resolution.height / 2, 0, 0, 1],
                          np.float32).reshape(3, 3)
    init_ext = look_at_extrinsics(-forward_dir * distance, up_dir)
    cameras = []
    for azi, alt in zip(azimuth, altitude):
        ext = _rotation(up_dir, azi) @ _rotation(right_dir, alt) @ init_ext
        cameras.append(CameraInfo.create("cam%d" % len(cameras), resolution, intrinsics,
                                         ext.astype(np.float32)))
    return cameras


class ETABar:
    """Console progress bar with the interface the scripts use (``next``/``info``/``finish``); the reference
    derives it from the ``progress`` package (utils.py:36-69)."""

    def __init__(self, message: str, max: int = 100):
        self.message, self.max, self.index, self.suffix = message, max, 0, ""
        import time
        self._t0 = time.time()

    def next(self, n: int = 1):
        import sys
        import time
        self.index += n
        if self.index == self.max or self.index % builtin_max(1, self.max // 50) == 0:
            el = time.time() - self._t0
            eta = el / builtin_max(1, self.index) * (self.max - self.index)
            sys.stdout.write("\r%s %5.1f%% - %ds %s" % (self.message, 100.0 * self.index / builtin_max(1, self.max), eta, self.suffix))
            sys.stdout.flush()

    def info(self, text: str):
        self.suffix = text

    def writeln(self, line: str):
        print(line)

    def finish(self):
        print()


builtin_max = max
