"""Data-parallel training on N GPUs of one box (config 4 of BASELINE.json: ray-sharded DDP): one process per GPU,
replicated model, every rank draws its own 1024-ray batches, ONE flat NCCL all-reduce of the 595,844 gradients per step
(``parallel.allreduce_gradients`` through ``Raycaster.fit(grad_sync=...)``), identical ClipAdam update on every rank.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \\
        tools/ddp_train.py [--steps 600]
Rank 0 prints one JSON line: global rays/s, ms/step, final validation PSNR, max parameter difference across ranks."""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fourier_feature_nets_b200 as ffn  # noqa: E402
from fourier_feature_nets_b200 import parallel  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=600)
    ap.add_argument("--res", type=int, default=100)
    ap.add_argument("--samples", type=int, default=64)
    ap.add_argument("--batch", type=int, default=1024)
    args = ap.parse_args()
    rank, ws, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if ws > 1:
        dist.init_process_group("nccl", device_id=dev)
    data = os.path.join(tempfile.gettempdir(), "ffn_ddp_scene.npz")
    if rank == 0:
        subprocess.run([sys.executable, os.path.join(ROOT, "tools", "make_synthetic_dataset.py"), data, "--resolution",
                        str(args.res), "--train", "30", "--val", "12", "--test", "2", "--steps", "128"], check=True,
                       capture_output=True)
    if ws > 1:
        dist.barrier()
    torch.manual_seed(20080524)
    model = ffn.NeRF(8, 256, 9, 10, 3, 4, [4], True).to(dev)
    parallel.broadcast_parameters(model)
    np.random.seed(1000 + rank)                       # every rank shuffles its own ray order
    train = ffn.ImageDataset.load(data, "train", args.samples, True, True).to(dev)
    train.sampler.seed += 7919 * rank                 # ... and draws its own stratified jitter
    val = ffn.ImageDataset.load(data, "val", args.samples, True, False).to(dev)
    rc = ffn.Raycaster(model)
    sync = parallel.allreduce_gradients if ws > 1 else None
    rc.fit(train, val, args.batch, 5e-4, 12, 0, 1000000, 0.1, 250000, 0.0, [], grad_sync=sync)      # warm-up
    validate = rc._validate
    rc._validate = lambda *a, **k: float("nan")       # time the optimisation steps only
    torch.cuda.synchronize()
    if ws > 1:
        dist.barrier()
    t0 = time.perf_counter()
    rc.fit(train, val, args.batch, 5e-4, args.steps, 0, 1000000, 0.1, 250000, 0.0, [], grad_sync=sync)
    torch.cuda.synchronize()
    wall = torch.tensor([time.perf_counter() - t0], device=dev)
    if ws > 1:
        dist.all_reduce(wall, op=dist.ReduceOp.MAX)
    psnr = validate(val, args.batch, args.steps)
    flat = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
    spread = torch.zeros((1,), device=dev)
    if ws > 1:
        lo, hi = flat.clone(), flat.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        spread = (hi - lo).abs().max().reshape(1)
    if rank == 0:
        steps = args.steps + 1
        print(json.dumps({"n_gpus": ws, "steps": steps, "ms_per_step": round(wall.item() / steps * 1e3, 3),
                          "global_batch_rays": args.batch * ws, "samples_per_ray": args.samples,
                          "train_rays_per_s": round(args.batch * ws * steps / wall.item()),
                          "val_psnr_after": round(psnr, 3), "max_param_spread_across_ranks": spread.item(),
                          "collective": "one flat NCCL all-reduce of 595,844 fp32 gradients per step"}))
    if ws > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
