"""Where do orbit-frame differences vs the reference come from?  For one orbit frame of a checkpoint: hierarchical t
values from (a) ffn_focus_sample (fp16 coarse pass) and (b) the reference algorithm in fp32 torch on the same GPU;
then pixels of the fused fine pass vs the fp32 torch definition on (a)'s and on (b)'s samples.
    python tools/frame_parity_probe.py model.pt [--res 40] [--samples 32]"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fourier_feature_nets_b200 as ffn  # noqa: E402
from fourier_feature_nets_b200.ray_sampler import _determine_cdf  # noqa: E402
from fourier_feature_nets_b200.utils import linspace  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("model")
    ap.add_argument("--res", type=int, default=40)
    ap.add_argument("--samples", type=int, default=32)
    ap.add_argument("--frames", type=int, default=3)
    ap.add_argument("--operand", default="fp16", choices=["fp16", "bf16", "fp16x3"])
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    model = ffn.load_model(args.model).to(dev).eval()
    model.ffn_operand = args.operand
    cams = ffn.orbit(np.array([0, 1, 0], np.float32), np.array([0, 0, -1], np.float32), args.frames, 40,
                     ffn.Resolution(args.res, args.res), 4)
    bounds = np.diag([2, 2, 2, 1]).astype(np.float32)
    S = args.samples
    fs = ffn.RaySampler(bounds, cams, S, False, model, 4096, device=dev)
    rc = ffn.Raycaster(model)
    out = []
    with torch.no_grad():
        for cam in range(args.frames):
            b = fs.rays_for_camera(cam)
            t_ours = b.focus_t()
            # the reference algorithm (ray_sampler.py:161-166, 301-357, 367-392) in fp32 torch
            n_u, n_f = S // 2, S - S // 2
            near, far = b.near_raw, b.far_raw
            tc = linspace(near, far, n_f)
            pos = (b.starts[:, None] + tc[..., None] * b.directions[:, None]).reshape(-1, 3)
            view = b.directions[:, None].expand(-1, n_f, -1).reshape(-1, 3)
            raw = model.forward_torch(pos, view) if model.use_view else model.forward_torch(pos)
            sigma = torch.nn.functional.softplus(raw[:, -1]).reshape(-1, n_f)
            cdf = _determine_cdf(tc, sigma)
            tm = 0.5 * (tc[..., :-1] + tc[..., 1:])
            u = torch.linspace(0., 1., n_f, device=dev).unsqueeze(0).repeat(len(near), 1)
            k = torch.searchsorted(cdf, u, right=True)
            i = (k - 1).clamp_min(0)
            j = k.clamp_max(cdf.shape[-1] - 1)
            ci, cj = torch.gather(cdf, 1, i), torch.gather(cdf, 1, j)
            ti, tj = torch.gather(tm, 1, i), torch.gather(tm, 1, j)
            den = cj - ci
            den = torch.where(den < 1e-5, torch.ones_like(den), den)
            focus = ti + (u - ci) / den * (tj - ti)
            t_ref, _ = torch.cat([linspace(b.near, b.far, n_u), focus], -1).sort(-1)
            dt = (t_ours - t_ref).abs()

            def torch_render(t):
                n = len(t)
                dirs = b.directions.reshape(n, 1, 3).repeat(1, S, 1)
                p = b.starts.reshape(n, 1, 3) + t.unsqueeze(-1) * dirs
                model.forward = model.forward_torch
                try:
                    return rc._render_torch(ffn.RaySamples(p, dirs, t, b.rays), False).color
                finally:
                    del model.forward
            c_fused_ours = rc.render(b, False).color
            c_fused_tref = model._ffn_engine.net.render_rays_t(b.starts, b.directions, t_ref, False)[0]
            c_torch_tref = torch_render(t_ref)
            c_torch_tours = torch_render(t_ours)
            out.append({
                "camera": cam, "rays": len(near),
                "t_max_abs": float(dt.max()), "t_rays_over_1e-3": int((dt.max(-1)[0] > 1e-3).sum()),
                "pix_fused(t_ours)_vs_torch(t_ref)": float((c_fused_ours - c_torch_tref).abs().max()),
                "pix_fused(t_ref)_vs_torch(t_ref)": float((c_fused_tref - c_torch_tref).abs().max()),
                "pix_torch(t_ours)_vs_torch(t_ref)": float((c_torch_tours - c_torch_tref).abs().max()),
                "sigma_max": float(sigma.max())})
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
