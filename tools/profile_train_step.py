"""Where a training step goes: GPU-busy time by kernel and the host-side wall time per phase.
    python tools/profile_train_step.py [--rays 1024] [--samples 128]
"""
import argparse
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fourier_feature_nets_b200 as ffn  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rays", type=int, default=1024)
ap.add_argument("--samples", type=int, default=128)
ap.add_argument("--steps", type=int, default=20)
ap.add_argument("--fused-adam", action="store_true")
args = ap.parse_args()
dev = torch.device("cuda:0")
R, S = args.rays, args.samples
torch.manual_seed(20080524)
model = ffn.NeRF(8, 256, 9, 10, 3, 4, [4], True).to(dev)
rc = ffn.Raycaster(model)
opt = torch.optim.Adam(model.parameters(), 5e-4, fused=True) if args.fused_adam else ffn.ClipAdam(model.parameters(), 5e-4)
g = torch.Generator(device=dev).manual_seed(0)
o = torch.tensor([0.0, 0.3, -4.0], device=dev).repeat(R, 1)
d = torch.nn.functional.normalize(torch.randn((R, 3), device=dev, generator=g) * 0.15
                                  + torch.tensor([0, 0, 1.0], device=dev), dim=-1)
near, far = torch.full((R,), 3.0, device=dev), torch.full((R,), 5.0, device=dev)
gt_c, gt_a = torch.rand((R, 3), device=dev, generator=g), torch.rand((R,), device=dev, generator=g)
phases = {"forward": 0.0, "backward": 0.0, "clip": 0.0, "adam": 0.0}


def step(i, timed=False):
    def mark(name, t0):
        if timed:
            torch.cuda.synchronize()
            phases[name] += time.perf_counter() - t0
        return time.perf_counter()
    t = time.perf_counter()
    b = ffn.RayBundle(o, d, near, far, None, S, True, None, seed=i)
    opt.zero_grad()
    out = rc.render(b, True)
    loss = (out.color - gt_c).square().mean() + 0.1 * (out.alpha - gt_a).square().mean()
    t = mark("forward", t)
    loss.backward()
    t = mark("backward", t)
    if args.fused_adam:
        torch.nn.utils.clip_grad_value_(model.parameters(), 0.1)
        torch.nn.utils.clip_grad_norm_(model.parameters(), 0.1)
    t = mark("clip", t)
    opt.step()
    mark("adam", t)


for i in range(5):
    step(i)
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(args.steps):
    step(10 + i)
torch.cuda.synchronize()
print("untimed-phase step: %.3f ms" % ((time.perf_counter() - t0) / args.steps * 1e3))
for i in range(args.steps):
    step(100 + i, True)
print("phases with a sync after each (ms/step):", {k: round(v / args.steps * 1e3, 3) for k, v in phases.items()})
from torch.profiler import ProfilerActivity, profile  # noqa: E402
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for i in range(args.steps):
        step(200 + i)
    torch.cuda.synchronize()
rows = [e for e in prof.key_averages() if e.device_time_total > 0]
rows.sort(key=lambda e: -e.device_time_total)
tot = sum(e.device_time_total for e in rows)
print("GPU busy per step: %.3f ms in %d launches" % (tot / args.steps / 1e3, sum(e.count for e in rows) / args.steps))
for e in rows[:60]:
    print("  %7.1f us  x%-3d %s" % (e.device_time_total / args.steps, e.count // args.steps, e.key[:90]))
