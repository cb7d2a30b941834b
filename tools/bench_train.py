"""Training-step timing: CUDA-kernel path vs. fp32 PyTorch autograd over the same definition (what the
reference runs on a GPU).  One step = render (with depth) + loss + backward + clip + Adam, as
Raycaster.fit (ray_caster.py:319-329).   python tools/bench_train.py [--rays 1024] [--samples 128]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fourier_feature_nets_b200 as ffn  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rays", type=int, default=1024)
ap.add_argument("--samples", type=int, default=128)
ap.add_argument("--steps", type=int, default=30)
args = ap.parse_args()
dev = torch.device("cuda:0")
R, S = args.rays, args.samples
res = {}
for name, use_kernels in (("torch_fp32_autograd", False), ("libffn_b200", True)):
    torch.manual_seed(20080524)
    model = ffn.NeRF(8, 256, 9, 10, 3, 4, [4], True).to(dev)
    rc = ffn.Raycaster(model)
    rc.train_kernels = use_kernels
    # the optimiser of Raycaster.fit: ClipAdam (clips + Adam in two launches) on the kernel path, PyTorch calls otherwise
    opt = ffn.ClipAdam(model.parameters(), 5e-4) if use_kernels else torch.optim.Adam(model.parameters(), 5e-4)
    g = torch.Generator(device=dev).manual_seed(0)
    o = torch.tensor([0.0, 0.3, -4.0], device=dev).repeat(R, 1)
    d = torch.nn.functional.normalize(torch.randn((R, 3), device=dev, generator=g) * 0.15
                                      + torch.tensor([0, 0, 1.0], device=dev), dim=-1)
    near = torch.full((R,), 3.0, device=dev)
    far = torch.full((R,), 5.0, device=dev)
    gt_c = torch.rand((R, 3), device=dev, generator=g)
    gt_a = torch.rand((R,), device=dev, generator=g)

    def step(i):
        b = ffn.RayBundle(o, d, near, far, None, S, True, None, seed=i)
        opt.zero_grad()
        out = rc.render(b, True)
        loss = (out.color - gt_c).square().mean() + 0.1 * (out.alpha - gt_a).square().mean()
        loss.backward()
        if not use_kernels:
            torch.nn.utils.clip_grad_value_(model.parameters(), 0.1)
            torch.nn.utils.clip_grad_norm_(model.parameters(), 0.1)
        opt.step()
        return loss

    for i in range(5):
        step(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        step(5 + i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    res[name] = {"ms_per_step": ms, "rays_per_s": R / ms * 1e3, "peak_mem_gb": torch.cuda.max_memory_allocated() / 2**30}
    torch.cuda.reset_peak_memory_stats()
res["speedup"] = res["torch_fp32_autograd"]["ms_per_step"] / res["libffn_b200"]["ms_per_step"]
res["config"] = {"rays": R, "samples": S, "model": "NeRF(8,256,9,10,3,4,[4],True)"}
print(json.dumps(res))
