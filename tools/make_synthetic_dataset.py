"""Procedural lego_400-shaped dataset (README.md:125-139 of the reference: images (N,H,W,4) uint8,
intrinsics (N,3,3), extrinsics (N,4,4), bounds (4,4), split_counts (3,)): an analytic emissive volume
(two coloured blobs and a box frame inside [-1,1]^3) ray-marched with 256 steps.  No network needed.

    python tools/make_synthetic_dataset.py out.npz --resolution 400 --train 100 --val 7 --test 13
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fourier_feature_nets_b200 as ffn  # noqa: E402
from fourier_feature_nets_b200.utils import look_at_extrinsics  # noqa: E402


def scene_smooth(p):
    """A scene a NeRF fits to > 30 dB within 20 k steps: two soft-edged blobs, smooth colours, no thin structures."""
    c1, c2 = np.array([0.3, 0.05, 0.0]), np.array([-0.3, -0.15, 0.2])
    d1 = np.linalg.norm(p - c1, axis=-1)
    d2 = np.linalg.norm(p - c2, axis=-1)
    sigma = 30 * (1 / (1 + np.exp((d1 - 0.42) * 12))) + 20 * (1 / (1 + np.exp((d2 - 0.34) * 12)))
    rgb = np.stack([0.5 + 0.4 * np.sin(3 * p[..., 0] + 1), 0.5 + 0.4 * np.sin(2.5 * p[..., 1] + 2),
                    0.5 + 0.4 * np.sin(3.5 * p[..., 2])], -1)
    return sigma, rgb


def scene_solid(p):
    """Opaque objects with (nearly) binary alpha like the reference's lego renders: the reference compares the
    composited colour with the image's straight RGB (image_dataset.py:253-260), which only agree where alpha is 0 or 1.
    Two hard spheres and a rounded box, smooth surface colours."""
    c1, c2 = np.array([0.3, 0.05, 0.0]), np.array([-0.3, -0.15, 0.2])
    d1 = np.linalg.norm(p - c1, axis=-1) - 0.40
    d2 = np.linalg.norm(p - c2, axis=-1) - 0.32
    q = np.abs(p - np.array([0.0, 0.45, -0.35])) - np.array([0.22, 0.12, 0.18])
    d3 = np.linalg.norm(np.maximum(q, 0), axis=-1) + np.minimum(q.max(-1), 0) - 0.04
    d = np.minimum(np.minimum(d1, d2), d3)
    sigma = 400.0 / (1 + np.exp(np.clip(d * 150, -60, 60)))
    rgb = np.stack([0.5 + 0.4 * np.sin(3 * p[..., 0] + 1), 0.5 + 0.4 * np.sin(2.5 * p[..., 1] + 2),
                    0.5 + 0.4 * np.sin(3.5 * p[..., 2])], -1)
    return sigma, rgb


def scene(p):
    """p (...,3) -> sigma (...), rgb (...,3)."""
    c1, c2 = np.array([0.35, 0.1, 0.0]), np.array([-0.3, -0.2, 0.25])
    d1 = np.linalg.norm(p - c1, axis=-1)
    d2 = np.linalg.norm(p - c2, axis=-1)
    sigma = 40 * (1 / (1 + np.exp((d1 - 0.38) * 40))) + 25 * (1 / (1 + np.exp((d2 - 0.3) * 40)))
    q = np.abs(p)
    edge = (np.sort(q, -1)[..., 1] > 0.62) & (q.max(-1) < 0.7)      # box frame
    sigma = sigma + 30 * edge
    rgb = np.stack([0.5 + 0.5 * np.sin(6 * p[..., 0] + 1), 0.5 + 0.5 * np.sin(5 * p[..., 1] + 2),
                    0.5 + 0.5 * np.sin(7 * p[..., 2])], -1)
    rgb = np.where(edge[..., None], np.array([0.9, 0.85, 0.2]), rgb)
    return sigma, rgb


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("path")
    ap.add_argument("--resolution", type=int, default=64)
    ap.add_argument("--train", type=int, default=14)
    ap.add_argument("--val", type=int, default=3)
    ap.add_argument("--test", type=int, default=3)
    ap.add_argument("--steps", type=int, default=192)
    ap.add_argument("--scene", default="frame", choices=["frame", "smooth", "solid"])
    ap.add_argument("--binary-alpha", action="store_true",
                    help="alpha in {0, 1} like a mask render (no fractional silhouette pixels)")
    args = ap.parse_args()
    n = args.train + args.val + args.test
    res = args.resolution
    focal = .5 * res / np.tan(.5 * 40 * np.pi / 180)
    K = np.array([[focal, 0, res / 2], [0, focal, res / 2], [0, 0, 1]], np.float32)
    bounds = np.diag([2, 2, 2, 1]).astype(np.float32)
    rng = np.random.default_rng(0)
    images, intr, extr = [], [], []
    for i in range(n):
        z = rng.uniform(0.05, 0.9)
        a = rng.uniform(0, 2 * np.pi)
        r = np.sqrt(1 - z * z)
        pos = 4.0 * np.array([r * np.cos(a), z, r * np.sin(a)])
        E = look_at_extrinsics(pos, np.array([0, 1.0, 0])).astype(np.float32)
        cam = ffn.CameraInfo.create("c", ffn.Resolution(res, res), K, E)
        sampler = ffn.RaySampler(bounds, [cam], args.steps)
        o, d = sampler.starts.numpy(), sampler.directions.numpy()
        near, far = sampler.near_far.numpy()
        valid = sampler.valid_mask.numpy()
        img = np.zeros((res * res, 4), np.float32)
        t = near[valid, None] + np.linspace(0, 1, args.steps)[None, :] * (far - near)[valid, None]
        p = o[valid, None, :] + t[..., None] * d[valid, None, :]
        sigma, rgb = {"smooth": scene_smooth, "solid": scene_solid, "frame": scene}[args.scene](p)
        delta = np.diff(t, axis=1, append=t[:, -1:] + 1e-3)
        alpha = 1 - np.exp(-sigma * delta)
        T = np.cumprod(np.concatenate([np.ones_like(alpha[:, :1]), 1 - alpha[:, :-1] + 1e-10], 1), 1)
        w = alpha * T
        img[valid, :3] = (w[..., None] * rgb).sum(1)
        img[valid, 3] = w.sum(1)
        a_ = np.clip(img[:, 3:4], 1e-6, 1)
        img[:, :3] = np.where(img[:, 3:4] > 1e-3, img[:, :3] / a_, 0)       # un-premultiplied colour
        if args.binary_alpha:
            solid = img[:, 3] > 0.5
            img[:, 3] = solid
            img[~solid, :3] = 0
        images.append((np.clip(img, 0, 1) * 255).astype(np.uint8).reshape(res, res, 4))
        intr.append(K)
        extr.append(E)
    np.savez_compressed(args.path, images=np.stack(images), intrinsics=np.stack(intr), extrinsics=np.stack(extr),
                        bounds=bounds, split_counts=np.array([args.train, args.val, args.test]))
    print("wrote", args.path, np.stack(images).shape)


if __name__ == "__main__":
    main()
