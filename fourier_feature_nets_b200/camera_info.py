"""Pinhole camera model (API of the reference's fourier_feature_nets/camera_info.py:16-118).

Ray generation is one-off host-side setup feeding the hot path:
plain numpy, same operation order as camera_info.py:66-74,99-109 so that the ray
tables are bit-identical given identical cameras.
"""
from typing import NamedTuple

import numpy as np

Ray = NamedTuple("Ray", [("origin", np.ndarray), ("direction", np.ndarray)])


class Resolution(NamedTuple("Resolution", [("width", int), ("height", int)])):
    """Image width and height."""

    def scale_to_height(self, height: int) -> "Resolution":
        return Resolution(self.width * height // self.height, height)

    def square(self) -> "Resolution":
        side = min(self.width, self.height)
        return Resolution(side, side)

    @property
    def ratio(self) -> float:
        return self.width / self.height


class CameraInfo(NamedTuple("CameraInfo", [("name", str), ("resolution", Resolution),
                                           ("intrinsics", np.ndarray),
                                           ("extrinsics", np.ndarray)])):
    """name, resolution, 3x3 intrinsics, 4x4 camera-to-world extrinsics."""

    @staticmethod
    def create(name: str, resolution: Resolution, intrinsics: np.ndarray,
               extrinsics: np.ndarray) -> "CameraInfo":
        return CameraInfo(name, resolution, intrinsics[:3, :3], extrinsics)

    def _projection(self) -> np.ndarray:
        proj = np.eye(4, dtype=np.float32)
        proj[:3, :3] = self.intrinsics
        return proj @ np.linalg.inv(self.extrinsics)

    def unproject(self, points: np.ndarray) -> np.ndarray:
        """2-D pixel positions -> homogeneous world points on the z=1 plane."""
        unproj = np.linalg.inv(self._projection())
        pts = points.reshape(-1, 2)
        pts = np.concatenate([pts, np.ones((pts.shape[0], 2), np.float32)], axis=-1)
        return (unproj @ pts.T).T

    def project(self, positions: np.ndarray) -> np.ndarray:
        """3-D world positions -> 2-D pixel positions."""
        ones = np.ones((positions.shape[0], 1), np.float32)
        pts = (self._projection() @ np.concatenate([positions, ones], -1).T).T
        return pts[:, :2] / pts[:, 2:3]

    @property
    def fov_y_degrees(self) -> float:
        return float(2 * np.arctan((0.5 * self.resolution.width) / self.intrinsics[1, 1]) * 180 / np.pi)

    @property
    def position(self) -> np.ndarray:
        return self.extrinsics[:3, 3].reshape(1, 3)

    def raycast(self, points: np.ndarray) -> Ray:
        """Pixel positions -> (origins, unit directions) in world space."""
        world = self.unproject(points.astype(np.float32))
        origin = self.position
        direction = world[:, :3] - origin
        direction = direction / np.linalg.norm(direction, axis=-1, keepdims=True)
        return Ray(origin + 0 * direction, direction)
