"""Tiny driver for ncu: a few FusedTrainer steps (1024 rays x 128 samples, NeRF 8x256) on synthetic rays.
    ncu --set full --clock-control none -k regex:'ffn_render_kernel|ffn_wgrad_kernel' -s 6 -c 3 \\
        -o gpurun_out/train python tools/profile_train_kernels.py
(per step the library launches: train-forward render kernel, dgrad render kernel, ffn_wgrad_kernel)"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fourier_feature_nets_b200 as ffn  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rays", type=int, default=1024)
ap.add_argument("--samples", type=int, default=128)
ap.add_argument("--steps", type=int, default=4)
ap.add_argument("--operand", default="fp16")
args = ap.parse_args()
dev = torch.device("cuda:0")
torch.manual_seed(20080524)
model = ffn.NeRF(8, 256, 9, 10, 3, 4, [4], True).to(dev)
model.ffn_operand = args.operand
R, S = args.rays, args.samples
g = torch.Generator(device=dev).manual_seed(0)
o = torch.tensor([0.0, 0.3, -4.0], device=dev).repeat(R, 1)
d = torch.nn.functional.normalize(torch.randn((R, 3), device=dev, generator=g) * 0.15
                                  + torch.tensor([0, 0, 1.0], device=dev), dim=-1)
near, far = torch.full((R,), 3.0, device=dev), torch.full((R,), 5.0, device=dev)
gt_c, gt_a = torch.rand((R, 3), device=dev, generator=g), torch.rand((R,), device=dev, generator=g)
trainer = ffn.FusedTrainer(model, 5e-4)
lin = torch.linspace(0, 1, S).to(dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(args.steps):
    if i == args.steps - 1:
        e0.record()
    b = ffn.RayBundle(o, d, near, far, torch.arange(R, device=dev), S, True, None, seed=i)
    trainer.backward(b, gt_c, gt_a, 0.1, lin)
    trainer.update()
e1.record()
torch.cuda.synchronize()
print("last step: %.3f ms" % e0.elapsed_time(e1))
if os.environ.get("FFN_STATS"):
    # cycle counters of the render-kernel passes (forward-with-saves and dgrad accumulate into the same array)
    st = trainer.net.debug_stats()
    tot, wa, ww, n = st[:4]
    print("epilogue warp 4 (per CTA launch, cycles): wait-acc %.0f  convert+store %.0f  front %.0f  back %.0f" % (
        st[4] / n, st[5] / n, st[6] / n, st[7] / n))
    tiles_per_slot = args.steps * R * S / 128 / 2
    print("epilogue cycles per tile and layer (fwd + bwd summed):", " ".join("%.0f" % (x / tiles_per_slot) for x in st[8:8 + 12]))
    print("issuer warp: total %.0f cyc/CTA launch, wait-epilogue %.1f%%, wait-weights %.1f%%, issuing %.1f%%" % (
        tot / n, 100 * wa / tot, 100 * ww / tot, 100 * (tot - wa - ww) / tot))
