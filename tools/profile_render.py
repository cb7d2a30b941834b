"""Tiny driver for ncu: a few launches of the fused render kernel on synthetic rays.
    ncu --set full --clock-control none --import-source on -k regex:ffn_infer_kernel -s 2 -c 1 \
        -o gpurun_out/render python tools/profile_render.py --rays 262144 --iters 4
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fourier_feature_nets_b200 as ffn  # noqa: E402
from fourier_feature_nets_b200 import engine  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rays", type=int, default=262144)
ap.add_argument("--samples", type=int, default=64)
ap.add_argument("--iters", type=int, default=4)
ap.add_argument("--operand", default="fp16", choices=["fp16", "bf16", "fp16x3"])
args = ap.parse_args()
dev = torch.device("cuda:0")
torch.manual_seed(20080524)
model = ffn.NeRF(8, 256, 9, 10, 3, 4, [4], True).to(dev).eval()
eng = engine.get_engine(model, dev, args.operand)
R, S = args.rays, args.samples
gen = torch.Generator(device=dev).manual_seed(0)
o = torch.tensor([0.0, 0.3, -4.0], device=dev).repeat(R, 1)
d = torch.nn.functional.normalize(torch.randn((R, 3), device=dev, generator=gen) * 0.15
                                  + torch.tensor([0, 0, 1.0], device=dev), dim=-1)
near = torch.full((R,), 3.0, device=dev)
far = torch.full((R,), 5.0, device=dev)
lin = torch.linspace(0, 1, S).to(dev)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(args.iters):
    if i == args.iters - 1:
        e0.record()
    eng.net.render_rays(o, d, near, far, lin, None, True, 1, 0, S, True)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print("last launch %.3f ms -> %.2f M rays/s" % (ms, R / ms / 1e3))

if os.environ.get("FFN_STATS"):
    st = eng.net.debug_stats()
    tot, wa, ww, n = st[:4]
    if st[4] + st[5]:
        print("epilogue warp 4 (per CTA, cycles): wait-acc %.0f  convert+store %.0f | aux warp 2: encode %.0f  composite %.0f  "
              "wait-enc-free %.0f  wait-raw %.0f" % (st[4] / n, st[5] / n, st[6] / n, st[7] / n, st[30] / n, st[31] / n))
    # counters [8 + l]: summed over warp 4 (slot 0) of all 148 CTAs and over all launches
    tiles_per_slot = args.iters * R * S / 128 / 2
    print("epilogue cycles per tile and layer (acc-full -> A-ready):",
          " ".join("%.0f" % (x / tiles_per_slot) for x in st[8:8 + 12]))
    print("issuer warp: total %.0f cyc/CTA, wait-epilogue %.1f%%, wait-weights %.1f%%, issuing %.1f%%" % (
        tot / n, 100 * wa / tot, 100 * ww / tot, 100 * (tot - wa - ww) / tot))
