"""Sampling modes of a ray dataset (reference: RayDataset.Mode, fourier_feature_nets/ray_dataset.py:20-48)."""
from enum import Enum


class Mode(Enum):
    Full = 0      # all rays
    Sparse = 1    # a strided subset of pixels
    Center = 2    # a centre crop (first `crop_steps` of training)
    Dilate = 3    # the dilated object mask
    Patch = 4
