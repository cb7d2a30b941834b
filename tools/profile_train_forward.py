"""The training forward (ffn_train_forward: fused render + saves for the backward) alone, with the in-kernel cycle
counters (FFN_STATS=1) and the save switches of FFN_DBG_FLAGS (16: no sign words, 32: no activation saves):
    FFN_STATS=1 python tools/profile_train_forward.py [--rays 1024 --samples 128 --iters 20]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fourier_feature_nets_b200 as ffn  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rays", type=int, default=1024)
ap.add_argument("--samples", type=int, default=128)
ap.add_argument("--iters", type=int, default=20)
args = ap.parse_args()
dev = torch.device("cuda:0")
torch.manual_seed(20080524)
model = ffn.NeRF(8, 256, 9, 10, 3, 4, [4], True).to(dev).train()
rc = ffn.Raycaster(model)
R, S = args.rays, args.samples
g = torch.Generator(device=dev).manual_seed(0)
o = torch.tensor([0.0, 0.3, -4.0], device=dev).repeat(R, 1)
d = torch.nn.functional.normalize(torch.randn((R, 3), device=dev, generator=g) * 0.15
                                  + torch.tensor([0, 0, 1.0], device=dev), dim=-1)
near, far = torch.full((R,), 3.0, device=dev), torch.full((R,), 5.0, device=dev)
b = ffn.RayBundle(o, d, near, far, torch.arange(R, device=dev), S, True, None, seed=1)
for _ in range(3):
    out = rc.render(b, False)          # grad mode: autograd.RenderNeRF.forward = ffn_train_forward
assert out.color.requires_grad
eng = model.__dict__["_ffn_engine"]
if os.environ.get("FFN_STATS"):
    eng.net.debug_stats()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.iters):
    rc.render(b, False)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / args.iters
print("train forward: %.4f ms per call (%d x %d), FFN_DBG_FLAGS=%s" % (ms, R, S, os.environ.get("FFN_DBG_FLAGS", "0")))
if os.environ.get("FFN_STATS"):
    st = eng.net.debug_stats()
    tot, wa, ww, n = st[:4]
    print("epilogue warp 4 (per CTA launch, cycles): wait-acc %.0f  convert+store %.0f  front %.0f  back %.0f" % (
        st[4] / n, st[5] / n, st[6] / n, st[7] / n))
    tiles_per_slot = args.iters * R * S / 128 / 2
    print("epilogue cycles per tile and layer:", " ".join("%.0f" % (x / tiles_per_slot) for x in st[8:8 + 12]))
    print("issuer warp: total %.0f cyc/CTA launch, wait-epilogue %.1f%%, wait-weights %.1f%%, issuing %.1f%%" % (
        tot / n, 100 * wa / tot, 100 * ww / tot, 100 * (tot - wa - ww) / tot))
