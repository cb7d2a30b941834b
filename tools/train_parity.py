"""Matched-PSNR evidence for the training path (north_star: "at matched PSNR"): train the same NeRF from the same seed on
the same procedural lego-shaped dataset twice through ``Raycaster.fit`` -- once on the CUDA kernels (fused forward,
dgrad, ffn_wgrad, ClipAdam), once on fp32 PyTorch autograd of the reference definition -- and compare
  * validation PSNR of both runs (reference's own metric, ray_caster.py:220-246),
  * wall time per optimisation step of both runs,
  * the trained model rendered by the fused inference kernel vs the fp32 PyTorch definition (pixel max-abs, PSNR).
    python tools/train_parity.py [--steps 1500] [--res 100] [--out profiles/r01_train_parity.json]"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fourier_feature_nets_b200 as ffn  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=1500)
    ap.add_argument("--res", type=int, default=100)
    ap.add_argument("--cams", type=int, default=30)
    ap.add_argument("--samples", type=int, default=64)
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--out", default="")
    ap.add_argument("--operand", default="fp16")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    tmp = tempfile.mkdtemp()
    data = os.path.join(tmp, "scene.npz")
    subprocess.run([sys.executable, os.path.join(ROOT, "tools", "make_synthetic_dataset.py"), data, "--resolution",
                    str(args.res), "--train", str(args.cams), "--val", "12", "--test", "2", "--steps", "128"], check=True)
    # (12 validation cameras: > 102,400 validation rays, so that _validate takes the to_valid branch of
    # ray_caster.py:228-233; with fewer rays batches made of background-only rays average an empty loss = NaN, in the
    # reference as well)
    res = {"config": vars(args)}
    models = {}
    for label, kernels in (("libffn_b200", True), ("torch_fp32_autograd", False)):
        torch.manual_seed(20080524)
        np.random.seed(20080524)
        train = ffn.ImageDataset.load(data, "train", args.samples, True, True).to(dev)
        val = ffn.ImageDataset.load(data, "val", args.samples, True, False).to(dev)
        model = ffn.NeRF(8, 256, 9, 10, 3, 4, [4], True).to(dev)
        model.ffn_operand = args.operand
        rc = ffn.Raycaster(model)
        rc.train_kernels = kernels
        warm = ffn.Raycaster(ffn.NeRF(8, 256, 9, 10, 3, 4, [4], True).to(dev))     # one-time costs (module load, cuBLAS
        warm.train_kernels = kernels                                              # init) stay out of the timing
        warm.fit(train, val, args.batch, 5e-4, 12, 0, 1000000, 0.1, 250000, 0.0, [])
        torch.manual_seed(20080524)
        np.random.seed(20080524)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        log = rc.fit(train, val, args.batch, 5e-4, args.steps, 0, args.steps // 3, 0.1, 250000, 0.0, [])
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        res[label] = {"val_psnr": [round(e.val_psnr, 3) for e in log], "train_psnr": [round(e.train_psnr, 3) for e in log],
                      "steps": [e.step for e in log], "fit_wall_s": round(wall, 2),
                      "ms_per_step_incl_validation": round(wall / args.steps * 1e3, 3)}
        models[label] = model
        print(label, res[label], flush=True)
    # the kernel-trained model rendered both ways (inference parity on a TRAINED net)
    model = models["libffn_b200"].eval()
    val = ffn.ImageDataset.load(data, "val", args.samples, True, False).to(dev)
    rc = ffn.Raycaster(model)
    worst, mse, n = 0.0, 0.0, 0
    before = ffn._lib.launch_count()
    with torch.no_grad():
        for cam in range(val.num_cameras):
            rays = val.rays_for_camera(cam).to(dev)
            ours = rc.render(rays, True)
            mat = rays.materialize() if hasattr(rays, "materialize") else rays
            model.forward = model.forward_torch          # the reference definition as plain fp32 PyTorch ops
            launched = ffn._lib.launch_count()
            ref = rc._render_torch(mat, True)
            assert ffn._lib.launch_count() == launched, "the fp32 reference must not touch libffn_b200"
            del model.forward
            d = (ours.color - ref.color)
            worst = max(worst, d.abs().max().item(), (ours.alpha - ref.alpha).abs().max().item())
            mse += d.square().sum().item()
            n += d.numel()
    res["trained_model_inference_parity"] = {"pixel_max_abs": worst, "psnr_vs_fp32_db": round(-10 * np.log10(mse / n), 2),
                                             "pixels": n // 3, "kernel_launches": ffn._lib.launch_count() - before}
    print(json.dumps(res))
    if args.out:
        with open(args.out, "w") as f:
            json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
