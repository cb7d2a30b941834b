"""Minimal stand-in for the two scenepic calls the reference's render scripts make outside the package
(orbit_video.py:64 ``sp.Transforms.scale(2)``); scenepic itself is not installable offline."""
import numpy as np


class Transforms:
    @staticmethod
    def scale(s):
        m = np.eye(4, dtype=np.float32)
        m[:3, :3] *= np.asarray(s, np.float32)
        return m
