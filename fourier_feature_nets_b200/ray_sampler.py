"""``RaySampler`` / ``RaySamples`` with the reference's API
(fourier_feature_nets/ray_sampler.py:15-403).

B200-first differences (same results, different plumbing):

* ``sample()`` returns a :class:`RayBundle` -- a ``RaySamples`` that carries only
  per-ray (origin, direction, near, far[, jitter]) = 32 B/ray and materialises the
  (R,S,3) position / direction / (R,S) t tensors only if somebody reads them.
  ``Raycaster.render`` hands a bundle straight to the fused kernel, which evaluates
  ray_sampler.py:380-397 in registers; the 1.8 KB/ray of samples never exist in HBM.
* ray tables may be made device-resident (``RaySampler.to(device)``), so indexing, the
  valid-ray filter and the image scatter are tensor ops instead of Python lists/sets.
"""
import os
from typing import List, NamedTuple, Optional, Union

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from .camera_info import CameraInfo
from .utils import blend_weights_torch, linspace


class RaySamples(NamedTuple("RaySamples", [("positions", torch.Tensor),
                                           ("view_directions", torch.Tensor),
                                           ("t_values", torch.Tensor),
                                           ("rays", torch.Tensor)])):
    """Point samples ``start + direction * t`` grouped by ray: positions (R,S,3),
    view_directions (R,S,3), t_values (R,S), rays (R,) indices."""

    def to(self, *args, **kwargs) -> "RaySamples":
        return RaySamples(*[None if t is None else t.to(*args, **kwargs) for t in self])

    def pin_memory(self) -> "RaySamples":
        return RaySamples(*[None if t is None else t.pin_memory() for t in self])

    def subset(self, index: List[int]) -> "RaySamples":
        return RaySamples(*[None if t is None else t[index] for t in self])

    def numpy(self) -> "RaySamples":
        return RaySamples(*[None if t is None else t.cpu().numpy() for t in self])


class RayBundle(RaySamples):
    """Compact ``RaySamples``: per-ray segment + sampling recipe.

    ``jitter`` (R,S) holds explicit uniform draws (reference-exact stratified sampling,
    ray_sampler.py:383); when it is ``None`` and ``stratified`` is set the kernel draws
    Philox(seed, rays[i], sample) itself.
    """

    def __new__(cls, starts, directions, near, far, rays, num_samples: int,
                stratified: bool = False, jitter: Optional[torch.Tensor] = None, seed: int = 0,
                ray_offset: int = 0):
        self = super().__new__(cls, None, None, None, rays)
        self.starts = starts
        self.directions = directions
        self.near = near
        self.far = far
        self.num_samples = int(num_samples)
        self.stratified = bool(stratified)
        self.jitter = jitter
        self.seed = int(seed)
        # position of ray 0 of this bundle inside the bundle ``sample()`` returned: in-kernel jitter is
        # Philox(seed, ray_offset + i, sample), so the sub-batches of ``batched_render`` draw independent jitter
        self.ray_offset = int(ray_offset)
        self._stage = None          # pinned staging set these tensors live in (RaySampler.enable_pinned_staging)
        self._cache = None
        return self

    def __len__(self):  # NamedTuple length, kept for tuple compatibility
        return 4

    # ---- materialisation: exactly ray_sampler.py:380-397 ------------------------------
    def materialize(self) -> RaySamples:
        if self._cache is None:
            S = self.num_samples
            t = linspace(self.near, self.far, S)
            if self.stratified:
                jitter = self.jitter
                if jitter is None:
                    gen = torch.Generator(device=self.near.device).manual_seed(self.seed + 7919 * self.ray_offset)
                    jitter = torch.rand((len(self.near), S), dtype=torch.float32,
                                        device=self.near.device, generator=gen)
                scale = (self.far - self.near) / S
                t = t + jitter * scale.unsqueeze(-1)
            n = len(self.near)
            dirs = self.directions.reshape(n, 1, 3).repeat(1, S, 1)
            pos = self.starts.reshape(n, 1, 3) + t.unsqueeze(-1) * dirs
            self._cache = RaySamples(pos, dirs, t, super().__getitem__(3))
        return self._cache

    positions = property(lambda self: self.materialize()[0])
    view_directions = property(lambda self: self.materialize()[1])
    t_values = property(lambda self: self.materialize()[2])
    rays = property(lambda self: tuple.__getitem__(self, 3))

    def __iter__(self):
        return iter(self.materialize())

    def __getitem__(self, i):
        return self.materialize()[i]

    # ---- RaySamples API, staying compact ----------------------------------------------
    def _map(self, fn, ray_offset: Optional[int] = None) -> "RayBundle":
        return RayBundle(fn(self.starts), fn(self.directions), fn(self.near), fn(self.far),
                         None if self.rays is None else fn(self.rays), self.num_samples,
                         self.stratified, None if self.jitter is None else fn(self.jitter), self.seed,
                         self.ray_offset if ray_offset is None else ray_offset)

    def to(self, *args, **kwargs) -> "RayBundle":
        out = self._map(lambda t: t.to(*args, **kwargs))
        if self._stage is not None and out.starts.is_cuda:
            # the copies out of the pinned staging set are in flight on the current stream: sample() waits for this
            # event before it overwrites the set
            self._stage.event = torch.cuda.Event()
            self._stage.event.record()
        return out

    def pin_memory(self) -> "RayBundle":
        if self._stage is not None:         # already lives in pinned memory
            return self
        return self._map(lambda t: t.pin_memory())

    def _subset_offset(self, index):
        """(index, ray_offset of the subset): contiguous ranges become slices and keep their position; a gather
        gets a hashed offset so that its in-kernel jitter differs from the parent's."""
        if isinstance(index, (list, range)) and len(index) > 0:
            lo, hi = index[0], index[-1] + 1
            if hi - lo == len(index):         # contiguous: a view, not a gather
                return slice(lo, hi), self.ray_offset + lo
        if isinstance(index, slice):
            return index, self.ray_offset + (index.start or 0)
        return index, self.ray_offset

    def subset(self, index) -> "RayBundle":
        index, off = self._subset_offset(index)
        return self._map(lambda t: t[index], off)

    def numpy(self) -> RaySamples:
        return self.materialize().numpy()

    @property
    def num_rays(self) -> int:
        return len(self.near)


class FocusBundle(RayBundle):
    """``RayBundle`` whose samples are S//2 (stratified) uniform + S - S//2 CDF-focused t values, sorted
    (ray_sampler.py:388-392).  The coarse sigma pass, the CDF and the inverse-transform sampling run on the
    GPU per batch (``ffn_focus_sample``) instead of once per dataset on the host."""

    def __new__(cls, starts, directions, near, far, near_raw, far_raw, rays, num_samples, stratified,
                jitter_u, u_focus, seed, coarse_model):
        self = super().__new__(cls, starts, directions, near, far, rays, num_samples, stratified, jitter_u, seed)
        self.near_raw, self.far_raw = near_raw, far_raw     # un-annealed segment: the CDF lives on it
        self.u_focus = u_focus
        self.coarse_model = coarse_model
        return self

    def focus_t(self) -> torch.Tensor:
        """Sorted t values (R,S) on the device of the rays (must be CUDA)."""
        from . import engine as _engine
        S = self.num_samples
        n_u, n_f = S // 2, S - S // 2
        dev = self.starts.device
        lin_c, lin_u = torch.linspace(0, 1, n_f).to(dev), torch.linspace(0, 1, n_u).to(dev)
        model = self.coarse_model
        if _engine.supported(model) and getattr(model, "use_view", False):
            eng = _engine.get_engine(model, dev, _engine.coarse_operand(model))
            return eng.net.focus_sample(self.starts, self.directions, self.near_raw, self.far_raw, self.near,
                                        self.far, lin_c, lin_u, self.jitter, self.u_focus, self.stratified,
                                        self.seed, S, self.ray_offset)
        # any other opacity model on the device (e.g. ``Voxels``): its own forward gives the coarse raw outputs at
        # t_c = linspace(near, far, S_c) (ray_sampler.py:246-263), the CDF / inverse transform / sort run in
        # ``ffn_focus_t``
        from . import _lib
        n = len(self.near_raw)
        t_c = self.near_raw.unsqueeze(-1) + lin_c.unsqueeze(0) * (self.far_raw - self.near_raw).unsqueeze(-1)
        dirs = self.directions.reshape(n, 1, 3)
        pos = (self.starts.reshape(n, 1, 3) + t_c.unsqueeze(-1) * dirs).reshape(-1, 3)
        with torch.no_grad():
            if getattr(model, "use_view", False):
                raw = model(pos, dirs.expand(-1, n_f, -1).reshape(-1, 3))
            else:
                raw = model(pos)
        return _lib.focus_t(raw.reshape(n, n_f, -1)[..., -1].contiguous(), self.near_raw, self.far_raw, self.near,
                            self.far, lin_c, lin_u, self.jitter, self.u_focus, self.stratified, self.seed, S,
                            self.ray_offset)

    def materialize(self) -> RaySamples:
        if self._cache is None:
            t = self.focus_t()
            n, S = len(self.near), self.num_samples
            dirs = self.directions.reshape(n, 1, 3).repeat(1, S, 1)
            pos = self.starts.reshape(n, 1, 3) + t.unsqueeze(-1) * dirs
            self._cache = RaySamples(pos, dirs, t, tuple.__getitem__(self, 3))
        return self._cache

    def _map(self, fn, ray_offset: Optional[int] = None) -> "FocusBundle":
        out = FocusBundle(fn(self.starts), fn(self.directions), fn(self.near), fn(self.far), fn(self.near_raw),
                          fn(self.far_raw), None if self.rays is None else fn(self.rays), self.num_samples,
                          self.stratified, None if self.jitter is None else fn(self.jitter),
                          None if self.u_focus is None else fn(self.u_focus), self.seed, self.coarse_model)
        out.ray_offset = self.ray_offset if ray_offset is None else ray_offset
        return out


def _determine_cdf(t_values: torch.Tensor, opacity: torch.Tensor) -> torch.Tensor:
    """Coarse weights -> CDF over the S-2 interior bins (ray_sampler.py:59-67)."""
    weights = blend_weights_torch(t_values, opacity)[:, 1:-1] + 1e-5
    cdf = weights.cumsum(-1)
    cdf = cdf / cdf[:, -1:]
    return torch.cat([torch.zeros_like(cdf[:, :1]), cdf], dim=-1)


class RaySampler:
    """Samples points along the rays of a set of cameras inside a bounding volume."""

    def __init__(self, bounds: np.ndarray, cameras: List[CameraInfo], num_samples: int,
                 stratified=False, opacity_model: nn.Module = None, batch_size=4096,
                 anneal_start=0.5, num_anneal_steps=0, device=None):
        """Same arguments as the reference (ray_sampler.py:110-131) plus ``device``: a CUDA device (or the
        ``FFN_RAY_DEVICE`` environment variable) builds the ray tables there with ``ffn_generate_rays`` instead
        of numpy on the host (the reference spends ~0.75 s per 160 k rays in that loop)."""
        if device is None and os.environ.get("FFN_RAY_DEVICE"):
            device = os.environ["FFN_RAY_DEVICE"]
        self.bounds = bounds
        self.bounds_min = (bounds @ np.array([-0.5, -0.5, -0.5, 1], np.float32))[np.newaxis, :3]
        self.bounds_max = (bounds @ np.array([0.5, 0.5, 0.5, 1], np.float32))[np.newaxis, :3]
        self.image_width, self.image_height = cameras[0].resolution
        self.rays_per_camera = self.image_width * self.image_height
        self.num_rays = len(cameras) * self.rays_per_camera
        self.num_cameras = len(cameras)
        self.num_samples = num_samples
        self.anneal_start = anneal_start
        self.num_anneal_steps = num_anneal_steps
        self.cameras = cameras
        self.stratified = stratified
        self.opacity_model = opacity_model
        self.focus_sampling = opacity_model is not None
        # an opacity model that lives on a CUDA device is sampled lazily, per batch, on the GPU: no
        # constructor-time sigma pass over every ray, no (num_rays, S_c-1) CDF table
        self.lazy_focus = False
        if self.focus_sampling:
            self.opacity_model.eval()
            params = list(self.opacity_model.parameters())
            self.lazy_focus = bool(params and params[0].is_cuda)
        self.batch_size = batch_size
        self.seed = 20080524
        self._draws = 0
        # True: stratified jitter is drawn inside the kernel (Philox); False: on the host with
        # torch.rand like the reference (ray_sampler.py:383).  None = decide by table residency.
        self.device_jitter = None

        xs, ys = np.meshgrid(np.arange(self.image_width), np.arange(self.image_height))
        self.points = np.stack([xs, ys], -1).reshape(-1, 2)

        num_focus = num_samples - (num_samples // 2)
        self._invalid_set = None
        if device is not None and torch.device(device).type == "cuda":
            self._generate_on_device(torch.device(device), num_focus)
            return
        starts, directions, near_far, cdfs, valid = [], [], [], [], []
        for camera in cameras:
            o, d = camera.raycast(self.points)
            nf, ok = self._near_far(o, d)
            o, d, nf = torch.from_numpy(o), torch.from_numpy(d), torch.from_numpy(nf)
            starts.append(o)
            directions.append(d)
            near_far.append(nf)
            valid.append(torch.from_numpy(ok))
            if self.focus_sampling and not self.lazy_focus:
                t = linspace(nf[0], nf[1], num_focus)
                cdfs.append(_determine_cdf(t, self._determine_opacity(t, o, d)))
        self.starts = torch.cat(starts)
        self.directions = torch.cat(directions)
        self.near_far = torch.cat(near_far, -1)
        self.valid_mask = torch.cat(valid)
        if self.focus_sampling and not self.lazy_focus:
            self.cdfs = torch.cat(cdfs)

    def _generate_on_device(self, device: torch.device, num_focus: int):
        """Ray tables straight into HBM: the 4x4 inverses are taken on the host
        exactly like camera_info.py:101, everything per pixel runs in ``generate_rays_kernel``."""
        from . import _lib
        unproj = np.stack([np.linalg.inv(cam._projection()) for cam in self.cameras]).astype(np.float32)
        pos = np.stack([np.asarray(cam.position, np.float32).reshape(3) for cam in self.cameras])
        self.starts, self.directions, self.near_far, self.valid_mask = _lib.generate_rays(
            torch.from_numpy(unproj).to(device), torch.from_numpy(pos).to(device), self.bounds_min[0],
            self.bounds_max[0], self.image_width, self.image_height)
        if self.focus_sampling and not self.lazy_focus:
            cdfs = []
            for lo in range(0, self.num_rays, self.rays_per_camera):
                sl = slice(lo, lo + self.rays_per_camera)
                t = linspace(self.near_far[0, sl], self.near_far[1, sl], num_focus)
                cdfs.append(_determine_cdf(t, self._determine_opacity(t, self.starts[sl], self.directions[sl])))
            self.cdfs = torch.cat(cdfs)

    # ---- host tables: pinned staging ---------------------------------------------------------
    def enable_pinned_staging(self, max_rays: int, depth: int = 2) -> "RaySampler":
        """Host-resident ray tables: ``sample()`` gathers each batch (<= ``max_rays`` rays) straight into one of
        ``depth`` rotating sets of page-locked buffers, so ``bundle.to(device, non_blocking=True)`` is an async DMA
        and no pinned memory is allocated per call.  A set is rewritten only after the copies out of it have
        finished (event recorded by ``RayBundle.to``)."""
        class _Set:
            def __init__(self, n):
                self.starts = torch.empty((n, 3), dtype=torch.float32).pin_memory()
                self.directions = torch.empty((n, 3), dtype=torch.float32).pin_memory()
                self.near = torch.empty((n,), dtype=torch.float32).pin_memory()
                self.far = torch.empty((n,), dtype=torch.float32).pin_memory()
                self.rays = torch.empty((n,), dtype=torch.int64).pin_memory()
                self.event = None
        self._stage_sets = [_Set(int(max_rays)) for _ in range(depth)]
        self._stage_next = 0
        return self

    def _stage_gather(self, idx: torch.Tensor):
        st = self._stage_sets[self._stage_next]
        self._stage_next = (self._stage_next + 1) % len(self._stage_sets)
        if st.event is not None:
            st.event.synchronize()
            st.event = None
        n = len(idx)
        torch.index_select(self.starts, 0, idx, out=st.starts[:n])
        torch.index_select(self.directions, 0, idx, out=st.directions[:n])
        torch.index_select(self.near_far[0], 0, idx, out=st.near[:n])
        torch.index_select(self.near_far[1], 0, idx, out=st.far[:n])
        st.rays[:n].copy_(idx)
        return st, st.starts[:n], st.directions[:n], st.near[:n], st.far[:n], st.rays[:n]

    # ---- device residency (section 8f-1): keep the ray tables in HBM ----------------------
    def to(self, device) -> "RaySampler":
        self.starts = self.starts.to(device)
        self.directions = self.directions.to(device)
        self.near_far = self.near_far.to(device)
        self.valid_mask = self.valid_mask.to(device)
        if self.focus_sampling and not self.lazy_focus:
            self.cdfs = self.cdfs.to(device)
        return self

    @property
    def device(self) -> torch.device:
        return self.starts.device

    @property
    def invalid_rays(self) -> set:
        """Indices of rays missing the volume (a Python ``set`` in the reference,
        ray_sampler.py:140,230-231); built on first use from the boolean mask."""
        if self._invalid_set is None:
            self._invalid_set = set(torch.nonzero(~self.valid_mask).flatten().tolist())
        return self._invalid_set

    def _near_far(self, starts: np.ndarray, directions: np.ndarray):
        """Slab test against the AABB (ray_sampler.py:202-232): near = max of the per-axis
        entry distances (clamped to 0.1 for hits), far = min of the exits, hit iff near < far."""
        with np.errstate(divide="ignore", invalid="ignore"):
            t0 = (self.bounds_min - starts) / directions
            t1 = (self.bounds_max - starts) / directions
        near = np.where(t0 < t1, t0, t1).max(-1)
        far = np.where(t0 > t1, t0, t1).min(-1)
        hit = near < far
        near[hit] = np.maximum(0.1, near[hit])
        return np.stack([near, far]), hit

    def _determine_opacity(self, t_values: torch.Tensor, starts: torch.Tensor,
                           directions: torch.Tensor) -> torch.Tensor:
        """sigma = softplus(opacity_model(o + t d)[:, -1]) in ray batches (ray_sampler.py:234-269)."""
        num_rays, n = t_values.shape
        device = next(self.opacity_model.parameters()).device
        out = []
        with torch.no_grad():
            for lo in range(0, num_rays, self.batch_size):
                hi = min(lo + self.batch_size, num_rays)
                o = starts[lo:hi].to(device).unsqueeze(1)
                d = directions[lo:hi].to(device).unsqueeze(1)
                pos = (o + t_values[lo:hi].to(device).unsqueeze(2) * d).reshape(-1, 3)
                if self.opacity_model.use_view:
                    logits = self.opacity_model(pos, d.expand(-1, n, -1).reshape(-1, 3))[:, -1]
                else:
                    logits = self.opacity_model(pos)[:, -1]
                out.append(F.softplus(logits))
        return torch.cat(out).reshape(num_rays, -1).to(t_values.device)

    # ---- index plumbing -------------------------------------------------------------
    def _valid_for_camera(self, camera: int) -> torch.Tensor:
        lo = camera * self.rays_per_camera
        mask = self.valid_mask[lo:lo + self.rays_per_camera]
        return torch.nonzero(mask).flatten() + lo

    def rays_for_camera(self, camera: int) -> RaySamples:
        return self.sample(self._valid_for_camera(camera), None)

    def valid_mask_host(self) -> np.ndarray:
        """Boolean numpy copy of ``valid_mask`` (cached; the mask never changes after construction)."""
        if getattr(self, "_valid_np", None) is None:
            self._valid_np = self.valid_mask.cpu().numpy()
        return self._valid_np

    def to_valid(self, idx: Union[List[int], torch.Tensor]) -> Union[List[int], torch.Tensor]:
        """Keep the rays that intersect the volume (order preserved)."""
        if isinstance(idx, torch.Tensor):
            return idx[self.valid_mask[idx.to(self.valid_mask.device)].to(idx.device)]
        t = torch.as_tensor(idx, dtype=torch.long)
        return t[self.valid_mask.cpu()[t]].tolist() if len(t) else []

    def __len__(self) -> int:
        return self.num_rays

    def to_image(self, camera: int, colors, color_space: str) -> np.ndarray:
        """Scatter valid-ray colours into an (H,W,3) uint8 image (ray_sampler.py:177-200)."""
        idx = (self._valid_for_camera(camera) - camera * self.rays_per_camera).cpu().numpy()
        if isinstance(colors, torch.Tensor):
            colors = colors.detach().cpu().numpy()
        pixels = np.zeros((self.image_height * self.image_width, 3), np.float32)
        pixels[idx] = colors
        pixels = (pixels.reshape(self.image_height, self.image_width, 3) * 255).astype(np.uint8)
        if color_space == "YCrCb":
            import cv2
            pixels = cv2.cvtColor(pixels, cv2.COLOR_YCrCb2RGB)
        return pixels

    # ---- sampling -------------------------------------------------------------------
    def _sample_t_values(self, idx, num_samples: int) -> torch.Tensor:
        """Inverse-transform samples of the coarse CDF (ray_sampler.py:301-357)."""
        near, far = self.near_far[:, idx]
        n = len(near)
        t = linspace(near, far, num_samples)
        t = 0.5 * (t[..., :-1] + t[..., 1:])
        if self.stratified:
            u = torch.rand((n, num_samples), dtype=torch.float32).to(t.device)
        else:
            u = torch.linspace(0., 1., num_samples, device=t.device).unsqueeze(0).repeat(n, 1)
        cdf = self.cdfs[idx]
        k = torch.searchsorted(cdf, u, right=True)
        i = (k - 1).clamp_min(0)
        j = k.clamp_max(cdf.shape[-1] - 1)
        cdf_i, cdf_j = torch.gather(cdf, 1, i), torch.gather(cdf, 1, j)
        t_i, t_j = torch.gather(t, 1, i), torch.gather(t, 1, j)
        denom = cdf_j - cdf_i
        denom = torch.where(denom < 1e-5, torch.ones_like(denom), denom)
        return t_i + (u - cdf_i) / denom * (t_j - t_i)

    def sample(self, idx: Union[List[int], torch.Tensor], step: int) -> RaySamples:
        """Sample the requested rays (ray_sampler.py:359-403)."""
        if not isinstance(idx, torch.Tensor):
            idx = torch.as_tensor(idx, dtype=torch.long)
        idx_dev = idx.to(self.device)
        n = len(idx_dev)
        stage = None
        if (getattr(self, "_stage_sets", None) and not self.starts.is_cuda and not self.focus_sampling
                and n <= len(self._stage_sets[0].near)):
            stage, starts, directions, near, far, idx_dev = self._stage_gather(idx_dev)
        else:
            starts = self.starts[idx_dev]
            directions = self.directions[idx_dev]
            near, far = self.near_far[:, idx_dev]
        if step is not None and step < self.num_anneal_steps:
            anneal = min(max(step / self.num_anneal_steps, self.anneal_start), 1)
            mid = (near + far) * 0.5
            near = mid + (near - mid) * anneal
            far = mid + (far - mid) * anneal

        if not self.focus_sampling:
            jitter = None
            on_device = starts.is_cuda if self.device_jitter is None else self.device_jitter
            if self.stratified and not on_device:
                # host path keeps the reference's RNG stream (ray_sampler.py:383)
                jitter = torch.rand((n, self.num_samples), dtype=torch.float32)
            self._draws += 1
            bundle = RayBundle(starts, directions, near, far, idx_dev, self.num_samples,
                               self.stratified, jitter, self.seed + self._draws)
            bundle._stage = stage
            return bundle

        if self.lazy_focus:
            near_raw, far_raw = self.near_far[:, idx_dev]
            n_u = self.num_samples // 2
            jitter_u = u_focus = None
            on_device = starts.is_cuda if self.device_jitter is None else self.device_jitter
            if self.stratified and not on_device:
                # the reference's draw order: uniform part (:383) then focus part (:313)
                jitter_u = torch.rand((n, n_u), dtype=torch.float32)
                u_focus = torch.rand((n, self.num_samples - n_u), dtype=torch.float32)
            self._draws += 1
            return FocusBundle(starts, directions, near, far, near_raw, far_raw, idx_dev, self.num_samples,
                               self.stratified, jitter_u, u_focus, self.seed + self._draws, self.opacity_model)

        num_uniform = self.num_samples // 2
        t = linspace(near, far, num_uniform)
        if self.stratified:
            scale = (far - near) / num_uniform
            t = t + torch.rand((n, num_uniform), dtype=torch.float32).to(t.device) * scale.unsqueeze(-1)
        focus = self._sample_t_values(idx_dev, self.num_samples - num_uniform)
        t, _ = torch.cat([t, focus], -1).sort(-1)
        dirs = directions.reshape(n, 1, 3).repeat(1, self.num_samples, 1)
        positions = starts.reshape(n, 1, 3) + t.unsqueeze(-1) * dirs
        return RaySamples(positions, dirs, t, idx_dev)
