"""GPU bring-up: layer-by-layer comparison of the fused kernel against the oracle.
Run under gpurun:  python tools/bringup.py
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402  (test infrastructure)
from fourier_feature_nets_b200 import _lib  # noqa: E402
from fourier_feature_nets_b200.nerf_model import NeRF  # noqa: E402


def main():
    g = np.load(os.path.join(ROOT, "tests", "golden", "nerf_render.npz"))
    w = {k[2:]: g[k] for k in g.files if k.startswith("w.")}
    dev = torch.device("cuda:0")
    print(torch.cuda.get_device_name(0))
    model = NeRF(8, 256, 9, 10, 3, 4, [4], True)
    model.load_state_dict({k: torch.from_numpy(v) for k, v in w.items()})
    model = model.to(dev).eval()
    pos = torch.from_numpy(g["positions"]).reshape(-1, 3).to(dev)
    view = torch.from_numpy(g["view_directions"]).reshape(-1, 3).to(dev)
    hidden = oracle.nerf_hidden(w, g["positions"].reshape(-1, 3), g["view_directions"].reshape(-1, 3))

    from fourier_feature_nets_b200 import engine
    for operand in ("fp16", "bf16"):
        eng = engine.get_engine(model, dev, operand)
        torch.cuda.synchronize()
        print("== operand", operand)
        for l in range(10):
            t0 = time.time()
            out = eng.net.debug_layer(pos, view, l)
            torch.cuda.synchronize()
            out = out.cpu().numpy()
            ref = hidden[l]
            n = ref.shape[1]
            err = np.abs(out[:, :n] - ref)
            print("layer %d: max|ref| %.3f  max err %.3e  mean err %.3e  (%.1f ms)" % (
                l, np.abs(ref).max(), err.max(), err.mean(), 1e3 * (time.time() - t0)), flush=True)
            if err.max() > 0.05 * max(1.0, np.abs(ref).max()):
                bad = np.argwhere(err > 0.05 * max(1.0, np.abs(ref).max()))
                print("   first bad (row, col):", bad[:8].tolist(), " rows affected:",
                      len(set(bad[:, 0])), "cols affected:", len(set(bad[:, 1])))
                print("   got", out[bad[0][0], :8], "\n   ref", ref[bad[0][0], :8])
        raw = eng.net.mlp_forward(pos, view).cpu().numpy()
        print("raw: max err %.3e (max |ref| %.2f)" % (np.abs(raw - g["raw"]).max(), np.abs(g["raw"]).max()))
        c, a, d = eng.net.render_samples(torch.from_numpy(g["positions"]).to(dev),
                                         torch.from_numpy(g["view_directions"]).to(dev),
                                         torch.from_numpy(g["t_values"]).to(dev), True)
        torch.cuda.synchronize()
        print("render: color err %.3e alpha err %.3e depth mismatches %d/%d nan_flag %d" % (
            np.abs(c.cpu().numpy() - g["color"]).max(), np.abs(a.cpu().numpy() - g["alpha"]).max(),
            int((d.cpu().numpy() != g["depth"]).sum()), len(g["depth"]), eng.net.nan_flag()))

    # quick throughput probe
    eng = engine.get_engine(model, dev, "fp16")
    R, S = 262144, 64
    gen = torch.Generator(device=dev).manual_seed(0)
    o = torch.randn((R, 3), device=dev, generator=gen) * 0.1 + torch.tensor([0, 0, -4.0], device=dev)
    d = torch.nn.functional.normalize(torch.randn((R, 3), device=dev, generator=gen) * 0.1
                                      + torch.tensor([0, 0, 1.0], device=dev), dim=-1)
    near = torch.full((R,), 3.0, device=dev)
    far = torch.full((R,), 5.0, device=dev)
    lin = torch.linspace(0, 1, S).to(dev)
    for _ in range(3):
        eng.net.render_rays(o, d, near, far, lin, None, True, 1, 0, S, True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        eng.net.render_rays(o, d, near, far, lin, None, True, 1, 0, S, True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print("render_rays %d x %d: %.3f ms -> %.2f M rays/s, %.1f TFLOP/s" % (
        R, S, ms, R / ms / 1e3, R * S * 1186816 / ms / 1e9))


if __name__ == "__main__":
    main()
