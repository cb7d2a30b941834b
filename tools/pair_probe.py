"""Compare the render kernel variants: cta_group::1 (FFN_PAIR=0), pair (cta_group::2, the default) and pair with
the compositing at the end of each tile (FFN_DBG_FLAGS=8): outputs against the cta_group::1 kernel, launch time
and -- the stable A/B signal, wall time moves with the power cap -- issuer-warp cycles (FFN_STATS=1).
    timeout -s KILL 180 python tools/pair_probe.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fourier_feature_nets_b200 as ffn  # noqa: E402
from fourier_feature_nets_b200 import engine  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(20080524)
model = ffn.NeRF(8, 256, 9, 10, 3, 4, [4], True).to(dev).eval()
with torch.no_grad():
    model.opacity_out.weight.mul_(20.0)
eng = engine.get_engine(model, dev, "fp16")
VARIANTS = (("single", "0", "0"), ("pair", "1", "0"), ("pair/nodefer", "1", "8"))


def run(R, S, iters):
    gen = torch.Generator(device=dev).manual_seed(R)
    o = torch.tensor([0.0, 0.3, -4.0], device=dev).repeat(R, 1)
    d = torch.nn.functional.normalize(torch.randn((R, 3), device=dev, generator=gen) * 0.15
                                      + torch.tensor([0, 0, 1.0], device=dev), dim=-1)
    near = torch.full((R,), 3.0, device=dev)
    far = torch.full((R,), 5.0, device=dev)
    lin = torch.linspace(0, 1, S).to(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    out = None
    for i in range(iters):
        if i == iters - 1:
            e0.record()
        out = eng.net.render_rays(o, d, near, far, lin, None, True, 1, 0, S, True)
    e1.record()
    torch.cuda.synchronize()
    return out, e0.elapsed_time(e1)


for R, S in ((3, 64), (100, 64), (1000, 48), (4096, 64), (262144, 64)):
    print("R=%d S=%d" % (R, S), flush=True)
    base = None
    for name, pair, flags in VARIANTS:
        os.environ["FFN_PAIR"], os.environ["FFN_DBG_FLAGS"] = pair, flags
        out, ms = run(R, S, 3)
        if base is None:
            base = out
        err = max((x - y).abs().max().item() for x, y in zip(base, out) if x is not None)
        print("   %-13s %.3f ms (%.2f M rays/s)   max|x - default| = %.3g" % (name, ms, R / ms / 1e3, err), flush=True)
        if os.environ.get("FFN_STATS") and R > 100000:
            st = eng.net.debug_stats()
            tot, wa, ww, n = st[:4]
            print("      issuer: %.0f cycles/CTA  wait-epilogue %.1f%%  wait-weights %.1f%%  issuing %.1f%%" % (
                tot / n, 100 * wa / tot, 100 * ww / tot, 100 * (tot - wa - ww) / tot), flush=True)
