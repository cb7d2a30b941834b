"""``ImageDataset``: posed images -> per-ray ground truth + a ``RaySampler``
(API of the reference's fourier_feature_nets/image_dataset.py:20-472 and the
``RayDataset`` base, ray_dataset.py:16-242; scenepic visualisation omitted).

This is the caller either side of the hot path: ``get_rays`` feeds
``Raycaster.render`` and ``loss`` consumes its result.  Ground-truth colours / alphas and
the index tables can be moved to the GPU (``.to(device)``), making a training step free
of host->device traffic; on the host the behaviour (index modes, valid-ray filter,
GT zeroing where alpha == 0, MSE + 0.1 alpha-MSE) is the reference's.
"""
import os
from typing import List, Set, Union

import cv2
import numpy as np
import torch
import torch.nn as nn

from .camera_info import CameraInfo, Resolution
from .ray_dataset_modes import Mode
from .ray_sampler import RaySampler, RaySamples
from .utils import RenderResult


class RayDataset:
    """Namespace kept for ``RayDataset.Mode`` (ray_dataset.py:20)."""
    Mode = Mode


class ImageDataset(torch.utils.data.Dataset, RayDataset):
    """Dataset built from images for sampling from rays cast into a volume."""

    def __init__(self, label: str, images: np.ndarray, bounds: np.ndarray,
                 cameras: List[CameraInfo], num_samples: int, include_alpha=True, stratified=False,
                 opacity_model: nn.Module = None, batch_size=4096, color_space="RGB",
                 sparse_size=50, anneal_start=0.2, num_anneal_steps=0, alpha_weight=0.1):
        assert len(images.shape) == 4
        assert len(images) == len(cameras)
        assert images.dtype == np.uint8
        self._color_space = color_space
        self._mode = Mode.Full
        self.image_height, self.image_width = images.shape[1:3]
        self._images = images
        self._label = label
        self.include_alpha = include_alpha
        self._subsample_index = None
        self.sampler = RaySampler(bounds, cameras, num_samples, stratified, opacity_model,
                                  batch_size, anneal_start, num_anneal_steps)
        H, W = self.image_height, self.image_width
        rpc = self.sampler.rays_per_camera
        px, py = self.sampler.points[:, 0], self.sampler.points[:, 1]

        # centre crop: the middle half in each dimension (image_dataset.py:75-87)
        lo = np.array([W, H], np.float32) // 4
        hi = np.array([W, H], np.float32) - lo
        inside = ((self.sampler.points >= lo) & (self.sampler.points < hi)).all(-1)
        crop_points = torch.from_numpy(np.nonzero(inside)[0])
        self.crop_rays_per_camera = len(crop_points)

        sparse_points = torch.LongTensor(self._subsample_rays(sparse_size))
        self.sparse_size = sparse_size
        self.sparse_resolution = sparse_size * W // H, sparse_size
        self.sparse_rays_per_camera = len(sparse_points)

        radius = 8 * min(W, H) // 100
        element = cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (2 * radius + 1, 2 * radius + 1))
        self.dilate_ranges = []
        colors, alphas, crop_index, sparse_index, dilate_index = [], [], [], [], []
        num_dilate = 0
        for cam, image in enumerate(images):
            color = image[..., :3]
            if color_space == "YCrCb":
                color = cv2.cvtColor(color, cv2.COLOR_RGB2YCrCb)
            colors.append(torch.from_numpy((color.astype(np.float32) / 255)[py, px]))
            offset = cam * rpc
            if image.shape[-1] == 4:
                alpha = image[..., 3].astype(np.float32) / 255
                alphas.append(torch.from_numpy(alpha[py, px]))
                mask = cv2.dilate((alpha > 0).astype(np.uint8), element)[py, px]
                pts = torch.from_numpy(np.nonzero(mask)[0]) + offset
                dilate_index.append(pts)
                self.dilate_ranges.append((num_dilate, num_dilate + len(pts)))
                num_dilate += len(pts)
            crop_index.append(crop_points + offset)
            sparse_index.append(sparse_points + offset)
        self.crop_index = torch.cat(crop_index)
        self.sparse_index = torch.cat(sparse_index)
        self.dilate_index = torch.cat(dilate_index) if dilate_index else torch.zeros(0, dtype=torch.long)
        if alphas and include_alpha:
            self.alphas = torch.cat(alphas)
            self.alpha_weight = alpha_weight
        else:
            self.alphas = None
            self.alpha_weight = 0
        self.colors = torch.cat(colors)
        self.fused_loss = True          # False: always the element-wise PyTorch definition of ``loss``
        # host copies of the index tables: a batch given as a Python list is remapped and filtered on the host,
        # so a training step needs one small host->device copy and no device->host synchronisation
        self._host_tables = {Mode.Center: self.crop_index.numpy(), Mode.Sparse: self.sparse_index.numpy(),
                             Mode.Dilate: self.dilate_index.numpy()}

    # ---- device residency ---------------------------------------------------------------
    def to(self, device) -> "ImageDataset":
        """Move ground truth, index tables and the sampler's ray tables to ``device``."""
        self.colors = self.colors.to(device)
        if self.alphas is not None:
            self.alphas = self.alphas.to(device)
        self.crop_index = self.crop_index.to(device)
        self.sparse_index = self.sparse_index.to(device)
        self.dilate_index = self.dilate_index.to(device)
        self.sampler.to(device)
        return self

    # ---- properties of the reference ----------------------------------------------------
    color_space = property(lambda self: self._color_space)
    images = property(lambda self: self._images)
    label = property(lambda self: self._label)
    num_cameras = property(lambda self: self.sampler.num_cameras)
    num_samples = property(lambda self: self.sampler.num_samples)
    cameras = property(lambda self: self.sampler.cameras)

    @property
    def mode(self) -> Mode:
        return self._mode

    @mode.setter
    def mode(self, value: Mode):
        if value == Mode.Dilate and len(self.dilate_index) == 0:
            raise ValueError("Unable to use dilate mode: missing alpha channel")
        self._mode = value

    @property
    def subsample_index(self) -> Set[int]:
        return self._subsample_index

    @subsample_index.setter
    def subsample_index(self, index: Set[int]):
        self._subsample_index = index

    def to_valid(self, idx):
        return self.sampler.to_valid(idx)

    # ---- the step either side of the hot path --------------------------------------------
    def render(self, samples: RaySamples) -> RenderResult:
        """Ground-truth render of the rays (colour zeroed where GT alpha == 0)."""
        rays = samples.rays.to(self.colors.device)
        color = self.colors[rays]
        if self.alphas is None or self.mode == Mode.Dilate:
            alpha = None
        else:
            alpha = self.alphas[rays]
            color = torch.where(alpha.unsqueeze(1) > 0, color, torch.zeros_like(color))
        return RenderResult(color, alpha, None)

    def loss(self, _: int, rays: RaySamples, render: RenderResult) -> torch.Tensor:
        """MSE(colour) + alpha_weight * MSE(alpha)   (image_dataset.py:224-242).  With predictions and ground truth in
        HBM this is one launch (value + gradient, ``autograd.MSELoss``); otherwise the PyTorch ops of the reference."""
        pred = render.color
        if (self.fused_loss and pred.is_cuda and pred.dtype == torch.float32 and self.colors.device == pred.device
                and rays.rays.device == pred.device and rays.rays.dtype == torch.int64 and len(pred) > 0):
            from .autograd import MSELoss
            colors, alphas, weight = self.loss_tables()
            return MSELoss.apply(pred, render.alpha if alphas is not None else None, colors, alphas,
                                 rays.rays.contiguous(), weight)
        actual = self.render(rays).to(render.device)
        color_loss = (actual.color - render.color).square().mean()
        if self.alpha_weight > 0 and actual.alpha is not None:
            return color_loss + self.alpha_weight * (actual.alpha - render.alpha).square().mean()
        return color_loss

    def loss_tables(self):
        """(colors (N,3), alphas (N,) or None, alpha_weight) as ``loss`` uses them: no alpha term and no zeroing of
        the background colour without an alpha channel or in Dilate mode (image_dataset.py:253-260)."""
        use_alpha = self.alphas is not None and self.mode != Mode.Dilate
        return self.colors, (self.alphas if use_alpha else None), self.alpha_weight

    def _mode_index(self):
        return {Mode.Center: self.crop_index, Mode.Sparse: self.sparse_index,
                Mode.Dilate: self.dilate_index}.get(self.mode)

    def get_rays(self, idx: Union[List[int], torch.Tensor], step: int = None) -> RaySamples:
        """Samples of the selected rays (mode remap -> optional pixel subsample -> valid filter)."""
        if not torch.is_tensor(idx):
            # host path (the batch lists of Raycaster.fit / _validate): numpy on host copies of the tables
            idx = np.asarray(idx, dtype=np.int64).reshape(-1)
            table = self._host_tables.get(self.mode)
            if table is not None:
                idx = table[idx]
            if self.subsample_index:
                keep = np.fromiter(self.subsample_index, dtype=np.int64)
                idx = idx[np.isin(idx % self.sampler.rays_per_camera, keep)]
            idx = idx[self.sampler.valid_mask_host()[idx]]
            return self.sampler.sample(torch.from_numpy(idx), step)
        table = self._mode_index()
        if table is not None:
            idx = table[idx.to(table.device)]
        if self.subsample_index:
            keep = torch.as_tensor(sorted(self.subsample_index), dtype=torch.long, device=idx.device)
            idx = idx[torch.isin(idx % self.sampler.rays_per_camera, keep)]
        return self.sampler.sample(self.sampler.to_valid(idx), step)

    def _camera_span(self, camera: int):
        if self.mode == Mode.Center:
            return camera * self.crop_rays_per_camera, (camera + 1) * self.crop_rays_per_camera
        if self.mode == Mode.Sparse:
            return camera * self.sparse_rays_per_camera, (camera + 1) * self.sparse_rays_per_camera
        if self.mode == Mode.Dilate:
            return self.dilate_ranges[camera]
        if self.mode == Mode.Full:
            return camera * self.sampler.rays_per_camera, (camera + 1) * self.sampler.rays_per_camera
        raise NotImplementedError("Unsupported sampling mode")

    def index_for_camera(self, camera: int) -> List[int]:
        """Pixel indices (within the image) of the rays ``rays_for_camera`` returns."""
        lo, hi = self._camera_span(camera)
        table = self._mode_index()
        idx = torch.arange(lo, hi) if table is None else table[lo:hi].cpu()
        idx = torch.as_tensor(self.sampler.to_valid(idx), dtype=torch.long)
        return (idx - camera * self.sampler.rays_per_camera).tolist()

    def rays_for_camera(self, camera: int) -> RaySamples:
        lo, hi = self._camera_span(camera)
        return self.get_rays(torch.arange(lo, hi), None)

    def __len__(self) -> int:
        table = self._mode_index()
        return len(self.sampler) if table is None else len(table)

    def to_image(self, camera: int, colors: np.ndarray) -> np.ndarray:
        """(H,W,3) uint8 image from per-ray values in dataset order (ray_dataset.py:159-183)."""
        if isinstance(colors, torch.Tensor):
            colors = colors.detach().cpu().numpy()
        if len(colors.shape) == 1:
            colors = colors[..., np.newaxis]
        res = self.cameras[camera].resolution
        pixels = np.zeros((res.width * res.height, 3), np.float32)
        pixels[self.index_for_camera(camera)] = colors
        pixels = (pixels.reshape(res.height, res.width, 3) * 255).astype(np.uint8)
        if self._color_space == "YCrCb":
            pixels = cv2.cvtColor(pixels, cv2.COLOR_YCrCb2RGB)
        return pixels

    # ---- subsets ---------------------------------------------------------------------------
    def subset(self, cameras: List[int], num_samples: int, stratified: bool, label: str) -> "ImageDataset":
        s = self.sampler
        return ImageDataset(label, self.images[cameras], s.bounds, [s.cameras[i] for i in cameras],
                            num_samples, self.include_alpha, stratified, s.opacity_model, s.batch_size,
                            self.color_space, self.sparse_size, s.anneal_start, s.num_anneal_steps,
                            self.alpha_weight if self.alphas is not None else 0.1)

    def sample_cameras(self, num_cameras: int, num_samples: int, stratified: bool) -> "ImageDataset":
        """Farthest-point subset of the cameras (ray_dataset.py:185-215)."""
        if self.num_cameras < num_cameras:
            chosen = list(range(self.num_cameras))
        else:
            positions = np.concatenate([cam.position for cam in self.sampler.cameras])
            chosen = [0]
            while len(chosen) < num_cameras:
                d2 = np.square(positions[:, None, :] - positions[chosen][None, :, :]).sum(-1).min(-1)
                d2[chosen] = -1
                chosen.append(int(np.asarray(d2, np.float32).argmax()))
            chosen = sorted(set(chosen), key=chosen.index)
        return self.subset(list(chosen), num_samples, stratified, self.label)

    def _subsample_rays(self, resolution: int) -> List[int]:
        nx = resolution * self.image_width // self.image_height
        xs = (np.linspace(0, self.image_width - 1, nx) + 0.5).astype(np.int32)
        ys = (np.linspace(0, self.image_height - 1, resolution) + 0.5).astype(np.int32)
        xs, ys = np.meshgrid(xs, ys)
        return (ys.reshape(-1) * self.image_width + xs.reshape(-1)).tolist()

    @staticmethod
    def load(path: str, split: str, num_samples: int, include_alpha: bool, stratified: bool,
             opacity_model: nn.Module = None, batch_size=4096, color_space="RGB", sparse_size=50,
             anneal_start=0.2, num_anneal_steps=0) -> "ImageDataset":
        """Load a split from an NPZ with images (N,H,W,3|4) uint8, bounds (4,4), intrinsics (N,3,3),
        extrinsics (N,4,4), split_counts (3,)   (image_dataset.py:389-472; no downloads)."""
        if not os.path.exists(path):
            alt = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "data", path))
            if not os.path.exists(alt):
                print("Unable to find dataset", path, "(no network: assets cannot be downloaded)")
                return None
            path = alt
        data = np.load(path)
        total, height, width = data["images"].shape[:3]
        train_end = int(data["split_counts"][0])
        val_end = train_end + int(data["split_counts"][1])
        spans = {"train": (0, train_end), "val": (train_end, val_end), "test": (val_end, total)}
        if split not in spans:
            print("Unrecognized split:", split)
            return None
        idx = list(range(*spans[split]))
        cameras = [CameraInfo.create("{}{:03}".format(split, i), Resolution(width, height), intr, extr)
                   for i, (intr, extr) in enumerate(zip(data["intrinsics"][idx], data["extrinsics"][idx]))]
        return ImageDataset(split, data["images"][idx], data["bounds"], cameras, num_samples,
                            include_alpha, stratified, opacity_model, batch_size, color_space,
                            sparse_size, anneal_start, num_anneal_steps)
