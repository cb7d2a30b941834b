"""orbit_video.py on N GPUs of one box (BASELINE.json configs[4]: "360-frame inference render, 8xB200 ray-split"):
same arguments and output files as the reference script (orbit_video.py:14-97), frames dealt round-robin to the ranks --
frames (and their rays) are independent, so there is no collective on the render path; every rank builds the sampler
tables for ITS frames only (on the GPU), renders them with the fused kernels and writes its PNGs.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29521 \\
        tools/orbit_video_multi_gpu.py model.pt 800 out_dir --num-frames 360
A single process (no torchrun) renders every frame.  ``--device cpu`` + gloo is the CPU test configuration."""
import argparse
import json
import os
import sys
import time

import cv2
import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fourier_feature_nets_b200 as ffn  # noqa: E402

VECTORS = {name: np.array(v, np.float32) for name, v in {
    "x+": (1, 0, 0), "x-": (-1, 0, 0), "y+": (0, 1, 0), "y-": (0, -1, 0), "z+": (0, 0, 1), "z-": (0, 0, -1)}.items()}


def main():
    ap = argparse.ArgumentParser("Orbit Video Maker (frames sharded over the ranks)")
    ap.add_argument("model_path")
    ap.add_argument("resolution", type=int)
    ap.add_argument("output_dir")
    ap.add_argument("--opacity-model")
    ap.add_argument("--distance", type=float, default=4)
    ap.add_argument("--fov-y-degrees", type=float, default=40)
    ap.add_argument("--num-frames", type=int, default=200)
    ap.add_argument("--up-dir", default="y+", choices=list(VECTORS))
    ap.add_argument("--forward-dir", default="z-", choices=list(VECTORS))
    ap.add_argument("--num-samples", type=int, default=128)
    ap.add_argument("--batch_size", type=int, default=4096)
    ap.add_argument("--device", default="cuda")
    args = ap.parse_args()

    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    on_cuda = args.device.startswith("cuda")
    device = torch.device("cuda", local) if on_cuda else torch.device("cpu")
    if on_cuda:
        torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl" if on_cuda else "gloo")

    cameras = ffn.orbit(VECTORS[args.up_dir], VECTORS[args.forward_dir], args.num_frames, args.fov_y_degrees,
                        ffn.Resolution(args.resolution, args.resolution), args.distance)
    mine = list(range(rank, args.num_frames, world))                  # this rank's frames
    bounds = np.diag([2, 2, 2, 1]).astype(np.float32)                 # sp.Transforms.scale(2), orbit_video.py:64
    model = ffn.load_model(args.model_path).to(device)
    opacity = ffn.load_model(args.opacity_model).to(device) if args.opacity_model else model
    raycaster = ffn.Raycaster(model)
    sampler = ffn.RaySampler(bounds, [cameras[f] for f in mine], args.num_samples, False, opacity, args.batch_size,
                             device=device if on_cuda else None)
    os.makedirs(args.output_dir, exist_ok=True)
    if on_cuda:
        torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    rays = 0
    with torch.no_grad():
        for i, frame in enumerate(mine):
            image = raycaster.render_image(sampler, i, args.batch_size)
            rays += int(sampler.valid_mask[i * sampler.rays_per_camera:(i + 1) * sampler.rays_per_camera].sum())
            cv2.imwrite(os.path.join(args.output_dir, "frame_{:05d}.png".format(frame)),
                        cv2.cvtColor(image, cv2.COLOR_RGB2BGR))
    if on_cuda:
        torch.cuda.synchronize()
    stats = torch.tensor([time.perf_counter() - t0, float(rays)], dtype=torch.float64, device=device)
    if world > 1:
        wall = stats[:1].clone()
        dist.all_reduce(wall, op=dist.ReduceOp.MAX)                   # slowest rank
        total = stats[1:].clone()
        dist.all_reduce(total, op=dist.ReduceOp.SUM)
        stats = torch.cat([wall, total])
    if rank == 0:
        wall, total = stats.tolist()
        print(json.dumps({"n_ranks": world, "frames": args.num_frames, "resolution": args.resolution,
                          "samples_per_ray": args.num_samples, "seconds": round(wall, 3),
                          "frames_per_s": round(args.num_frames / wall, 2), "valid_rays_per_s": round(total / wall)}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
