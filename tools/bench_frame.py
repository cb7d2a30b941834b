"""End-to-end frame time of the orbit_video.py inner loop (orbit_video.py:76-84): ``Raycaster.render_image`` for
400x400 (or --res) look-at cameras on one B200, sampler tables built on the device.  Reports wall ms/frame, the GPU-busy
part, and the same with hierarchical (coarse -> fine) sampling as train_nerf.py / orbit_video.py use it.
    python tools/bench_frame.py [--res 400] [--frames 8] [--samples 128]"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fourier_feature_nets_b200 as ffn  # noqa: E402


def look_at(name, position, res, fov_y_deg=40.0):
    position = np.asarray(position, np.float32)
    fwd = -position / np.linalg.norm(position)
    up = np.array([0, 1, 0], np.float32)
    right = np.cross(fwd, up)
    right /= np.linalg.norm(right)
    down = np.cross(fwd, right)
    ext = np.eye(4, dtype=np.float32)
    ext[:3, 0], ext[:3, 1], ext[:3, 2], ext[:3, 3] = right, down, fwd, position
    f = 0.5 * res / np.tan(np.radians(fov_y_deg) / 2)
    intr = np.array([[f, 0, res / 2], [0, f, res / 2], [0, 0, 1]], np.float32)
    return ffn.CameraInfo.create(name, ffn.Resolution(res, res), intr, ext)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--res", type=int, default=400)
    ap.add_argument("--frames", type=int, default=8)
    ap.add_argument("--samples", type=int, default=128)
    ap.add_argument("--batch", type=int, default=4096)       # orbit_video.py:37
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    torch.manual_seed(20080524)
    model = ffn.NeRF(8, 256, 9, 10, 3, 4, [4], True).to(dev).eval()
    coarse = ffn.NeRF(8, 256, 9, 10, 3, 4, [4], True).to(dev).eval()
    cams = [look_at("f%d" % i, (4 * np.cos(a), 1.0, 4 * np.sin(a)), args.res)
            for i, a in enumerate(np.linspace(0, 2 * np.pi, args.frames, endpoint=False))]
    bounds = np.diag([2, 2, 2, 1]).astype(np.float32)
    out = {"res": args.res, "samples": args.samples, "frames": args.frames}
    for label, opacity, batch in (("uniform", None, args.batch), ("uniform_big_batch", None, 1 << 20),
                                  ("coarse_fine", coarse, args.batch), ("coarse_fine_big_batch", coarse, 1 << 20)):
        t0 = time.perf_counter()
        sampler = ffn.RaySampler(bounds, cams, args.samples, False, opacity, batch, device=dev)
        torch.cuda.synchronize()
        t_build = time.perf_counter() - t0
        rc = ffn.Raycaster(model)
        rc.render_image(sampler, 0, batch)                     # warm-up
        torch.cuda.synchronize()
        n_valid = int(sampler.valid_mask[:sampler.rays_per_camera].sum())
        t0 = time.perf_counter()
        for i in range(args.frames):
            img = rc.render_image(sampler, i, batch)
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) / args.frames
        assert img.shape == (args.res, args.res, 3) and img.dtype == np.uint8
        out[label] = {"sampler_build_s": round(t_build, 3), "ms_per_frame": round(wall * 1e3, 2),
                      "valid_rays_frame0": n_valid, "rays_per_s": round(n_valid / wall), "batch": batch}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
