"""``PixelDataset``: 2-D image regression data (config 0 of BASELINE.json, ``train_image_regression.py``), mirroring
fourier_feature_nets/pixel_dataset.py:26-199.  Host-side numpy/OpenCV only; this configuration runs on the CPU in the
reference and here (``FourierFeatureMLP.forward`` as plain PyTorch), it is not part of the CUDA hot path."""
import math
import os
from typing import NamedTuple

import cv2
import numpy as np
import torch

PixelData = NamedTuple("PixelData", [("uv", torch.Tensor), ("color", torch.Tensor)])


def _uv_grid(size: int) -> np.ndarray:
    """(size, size, 2) float32 grid over [0, 2) x [0, 2), x fastest (pixel_dataset.py:93-99,176-178)."""
    axis = np.linspace(0, 2, size, endpoint=False, dtype=np.float32)
    return np.stack(np.meshgrid(axis, axis), axis=-1)


class PixelDataset:
    """Square image as (uv, colour) pairs: every second pixel trains, all pixels validate."""

    def __init__(self, size: int, color_space: str, train_data: PixelData, val_data: PixelData):
        self.size = size
        self.color_space = color_space
        self.train_uv, self.train_color = train_data
        self.val_uv, self.val_color = val_data
        self.image = self.to_image(self.val_color)

    @staticmethod
    def create(path: str, color_space: str, size=512) -> "PixelDataset":
        """Centre-crop to a square, resize to ``size``, convert to RGB / YCrCb in [0, 1] (float64, as ``uint8 / 255``
        gives in the reference, pixel_dataset.py:62-88)."""
        if not os.path.exists(path):
            path = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "data", path))
        bgr = cv2.imread(path)
        if bgr is None:
            print("Unable to load image at", path)
            return None
        h, w = bgr.shape[:2]
        side = min(h, w)
        top, left = (h - side) // 2, (w - side) // 2
        bgr = bgr[top:top + side, left:left + side]
        if side != size:
            # third positional argument of cv2.resize is ``dst``: the reference's INTER_AREA lands there and the
            # interpolation stays at its default (bilinear); same call so the pixels are identical
            bgr = cv2.resize(bgr, (size, size), cv2.INTER_AREA)
        codes = {"YCrCb": cv2.COLOR_BGR2YCrCb, "RGB": cv2.COLOR_BGR2RGB}
        if color_space not in codes:
            raise NotImplementedError("Unsupported color space: {}".format(color_space))
        pixels = cv2.cvtColor(bgr, codes[color_space]) / 255
        train = PixelData(torch.from_numpy(_uv_grid(size // 2)), torch.from_numpy(pixels[::2, ::2, :]))
        val = PixelData(torch.from_numpy(_uv_grid(size)), torch.from_numpy(pixels))
        return PixelDataset(size, color_space, train, val)

    def to(self, *args) -> "PixelDataset":
        return PixelDataset(self.size, self.color_space,
                            PixelData(self.train_uv.to(*args), self.train_color.to(*args)),
                            PixelData(self.val_uv.to(*args), self.val_color.to(*args)))

    @staticmethod
    def generate_uvs(size: int, device) -> torch.Tensor:
        return torch.from_numpy(_uv_grid(size)).to(device=device)

    def to_image(self, colors: torch.Tensor, size=0) -> np.ndarray:
        """Predicted colours -> (size, size, 3) uint8 (truncating cast, pixel_dataset.py:159-163)."""
        size = size or self.size
        pixels = (colors * 255).reshape(size, size, 3).cpu().numpy().astype(np.uint8)
        if self.color_space == "YCrCb":
            pixels = cv2.cvtColor(pixels, cv2.COLOR_YCrCb2RGB)
        return pixels

    def psnr(self, colors: torch.Tensor) -> float:
        return -10 * math.log10(torch.square(colors - self.val_color).mean().item())

    def to_act_image(self, model, size: int) -> np.ndarray:
        """8 x 8 mosaic: tile k shows sigmoid(activation_k * W_out[:, k] + b_out) of the last hidden layer
        (pixel_dataset.py:114-150)."""
        grid, tile = 8, size // 8
        uvs = self.generate_uvs(tile, next(model.parameters()).device).reshape(-1, 2)
        model.keep_activations = True
        with torch.no_grad():
            model(uvs)
        model.keep_activations = False
        w_out = model.layers[-1].weight.detach().cpu().numpy()          # (3, C)
        b_out = model.layers[-1].bias.detach().cpu().numpy()
        act = np.asarray(model.activations[-1])                         # (tile*tile, C)
        mosaic = np.zeros((size, size, 3), np.float32)
        for k in range(grid * grid):
            logits = torch.from_numpy(act[:, k, None] * w_out[None, :, k] + b_out)
            r, c = divmod(k, grid)
            mosaic[r * tile:(r + 1) * tile, c * tile:(c + 1) * tile] = \
                torch.sigmoid(logits).numpy().reshape(tile, tile, 3)
        mosaic = (mosaic * 255).astype(np.uint8)
        if self.color_space == "YCrCb":
            mosaic = cv2.cvtColor(mosaic, cv2.COLOR_YCrCb2RGB)
        return mosaic
