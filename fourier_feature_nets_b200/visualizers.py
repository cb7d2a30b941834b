"""Training-time image writers with the reference's constructor signatures
(fourier_feature_nets/visualizers.py:34-263).  They only *call* the render path (``render`` =
``Raycaster.batched_render``), so they are callers of the hot path, not part of it."""
import os
from typing import Callable

import cv2
import numpy as np

from .camera_info import Resolution
from .image_dataset import ImageDataset
from .ray_sampler import RaySampler, RaySamples
from .utils import RenderResult, orbit

ImageRender = Callable[[RaySamples, bool], RenderResult]
ActivationRender = Callable[[RaySampler, int], np.ndarray]


class Visualizer:
    def visualize(self, step: int, render: ImageRender, act_render: ActivationRender):
        raise NotImplementedError


class EvaluationVisualizer(Visualizer):
    """Every ``interval`` steps: a 2x2 PNG (prediction | depth / ground truth | error) of one camera."""

    def __init__(self, results_dir: str, dataset: ImageDataset, interval: int, max_depth=10):
        self._output_dir = os.path.join(results_dir, dataset.label)
        os.makedirs(self._output_dir, exist_ok=True)
        self._dataset, self._interval, self._index, self._max_depth = dataset, interval, 0, max_depth

    def visualize(self, step: int, render: ImageRender, _: ActivationRender):
        if step % self._interval != 0:
            return
        ds = self._dataset
        camera = self._index % ds.num_cameras
        samples = ds.rays_for_camera(camera)
        act = ds.render(samples).numpy()
        pred = render(samples, True)
        error = np.square(act.color - pred.color).sum(-1)
        if act.alpha is not None:
            error = (3 * error + np.square(act.alpha - pred.alpha)) / 4
        width, height = ds.cameras[camera].resolution
        tiles = [ds.to_image(camera, np.clip(pred.color, 0, 1)),
                 ds.to_image(camera, np.clip(pred.depth, 0, self._max_depth) / self._max_depth),
                 ds.to_image(camera, act.color * act.alpha[..., np.newaxis] if act.alpha is not None else act.color),
                 ds.to_image(camera, np.sqrt(error) / max(np.sqrt(error).max(), 1e-12))]
        grid = np.zeros((height * 2, width * 2, 3), np.uint8)
        grid[:height, :width], grid[:height, width:] = tiles[0], tiles[1]
        grid[height:, :width], grid[height:, width:] = tiles[2], tiles[3]
        cv2.imwrite(os.path.join(self._output_dir, "s{:07}_c{:03}.png".format(step, camera)),
                    cv2.cvtColor(grid, cv2.COLOR_RGB2BGR))
        self._index += 1


class OrbitVideoVisualizer(Visualizer):
    """One orbit frame every ``num_steps // num_frames`` steps."""

    def __init__(self, results_dir: str, num_steps: int, resolution: Resolution, num_frames: int,
                 num_samples: int, color_space: str):
        self._output_dir = os.path.join(results_dir, "video")
        os.makedirs(self._output_dir, exist_ok=True)
        cameras = orbit(np.array([0, 1, 0]), np.array([0, 0, -1]), num_frames, 40, resolution.square(), 4)
        self._sampler = RaySampler(np.eye(4, dtype=np.float32) * 2, cameras, num_samples)
        self._interval = max(1, num_steps // num_frames)
        self._index, self._color_space = 0, color_space

    def visualize(self, step: int, render: ImageRender, _: ActivationRender):
        if step % self._interval != 0:
            return
        camera = self._index % self._sampler.num_cameras
        pred = render(self._sampler.rays_for_camera(camera), False)
        image = self._sampler.to_image(camera, pred.color, self._color_space)
        cv2.imwrite(os.path.join(self._output_dir, "frame_{:05d}.png".format(self._index)),
                    cv2.cvtColor(image, cv2.COLOR_RGB2BGR))
        self._index += 1


class ActivationVisualizer(Visualizer):
    """Lecture visualisation of hidden activations: out of scope; accepted and ignored so
    that train_tiny_nerf.py keeps running."""

    def __init__(self, *args, **kwargs):
        print("ActivationVisualizer: activation grids are not rendered by the B200 build")

    def visualize(self, step, render, act_render):
        return


class ComparisonVisualizer(ActivationVisualizer):
    pass
