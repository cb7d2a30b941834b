"""Trained-model parity.  On one B200:

1. Train the same NeRF (train_nerf.py's architecture and hyper-parameters) on the same procedural lego-shaped dataset
   with BOTH arms through ``Raycaster.fit`` (ray_caster.py:248-376 semantics):
     ours       fourier_feature_nets_b200 on the CUDA kernels (FusedTrainer)
     reference  the unmodified reference package (oracle/_ref) on the same GPU, fp32 PyTorch autograd
   and record the validation-PSNR curves (the reference's own metric, ray_caster.py:220-246).
2. Render the converged model with the fused inference kernel and with the fp32 definition (and, for good measure, the
   reference's own ``Raycaster.render`` on the same samples): pixel max-abs, PSNR, depth mismatches.
3. The same after scaling ``opacity_out`` until sigma reaches ~350 (the dynamic range of docs/ray_data.tsv).
4. Gradients of one training batch against fp64 autograd of the reference definition: relative L2 error per parameter of
   the kernel gradients and of fp32 autograd.

    python tools/train_parity_r2.py --steps 20000 --out gpurun_out/r02_train_parity.json"""
import argparse
import contextlib
import copy
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
from oracle import reference as refmod  # noqa: E402


def fit_arm(pkg, data, args, dev, label):
    torch.manual_seed(20080524)
    np.random.seed(20080524)
    with contextlib.redirect_stdout(sys.stderr):
        train = pkg.ImageDataset.load(data, "train", args.samples, True, True, None, args.batch, "RGB",
                                      anneal_start=0.2, num_anneal_steps=2000)
        val = pkg.ImageDataset.load(data, "val", args.samples, True, False, None, args.batch, "RGB")
        model = pkg.NeRF(8, 256, 9, 10, 3, 4, [4], True).to(dev)
        rc = pkg.Raycaster(model)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        log = rc.fit(train, val, args.batch, 5e-4, args.steps, 1000, args.report, 0.1, 250000, 0.0, [])
        torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    out = {"steps": [e.step for e in log], "val_psnr": [round(float(e.val_psnr), 3) for e in log],
           "train_psnr": [round(float(e.train_psnr), 3) for e in log], "fit_wall_s": round(wall, 1),
           "ms_per_step_incl_validation": round(wall / args.steps * 1e3, 3)}
    print(label, json.dumps(out), flush=True)
    return model, out


def inference_parity(model, ref_pkg, ref_model, val, dev, operand="fp16"):
    """Our fused kernel vs the fp32 PyTorch definition (and the reference's own render) on every validation camera."""
    model = copy.deepcopy(model)
    model.ffn_operand = operand
    rc = ffn.Raycaster(model.eval())
    worst_c = worst_a = 0.0
    mse, n, mism, rays_n, worst_ref = 0.0, 0, 0, 0, 0.0
    sig_max = 0.0
    with torch.no_grad():
        rc.render(val.rays_for_camera(0).to(dev), True)        # (re-packs the weight image once)
        for cam in range(val.num_cameras):
            rays = val.rays_for_camera(cam).to(dev)
            before = ffn._lib.launch_count()
            ours = rc.render(rays, True)
            assert ffn._lib.launch_count() == before + 1
            mat = rays.materialize()
            model.forward = model.forward_torch          # the definition as plain fp32 PyTorch ops
            try:
                ref = rc._render_torch(mat, True)
                raw = model.forward_torch(mat.positions.reshape(-1, 3), mat.view_directions.reshape(-1, 3))
            finally:
                del model.forward
            sig_max = max(sig_max, float(torch.nn.functional.softplus(raw[:, 3]).max()))
            worst_c = max(worst_c, float((ours.color - ref.color).abs().max()))
            worst_a = max(worst_a, float((ours.alpha - ref.alpha).abs().max()))
            mse += float((ours.color - ref.color).square().sum())
            n += ref.color.numel()
            mism += int((ours.depth != ref.depth).sum())
            rays_n += len(ref.depth)
            if ref_model is not None:       # the reference package's own render on the same materialised samples
                rs = ref_pkg.RaySamples(mat.positions, mat.view_directions, mat.t_values, mat.rays)
                rr = ref_pkg.Raycaster(ref_model).render(rs, True)
                worst_ref = max(worst_ref, float((rr.color - ref.color).abs().max()))
    return {"rays": rays_n, "color_max_abs": worst_c, "alpha_max_abs": worst_a,
            "psnr_vs_fp32_db": round(float(-10 * np.log10(max(mse / n, 1e-20))), 2),
            "depth_mismatch_frac": mism / rays_n, "sigma_max": round(sig_max, 1),
            "fp32_definition_vs_reference_package_max_abs": worst_ref}


def gradient_check(model, train, dev, R=512, S=64):
    """Kernel gradients and fp32-autograd gradients of one batch, both against fp64 autograd of the definition."""
    torch.manual_seed(5)
    idx = train.sampler.to_valid(torch.randperm(len(train.sampler))[:4 * R].to(train.sampler.device))[:R]
    bundle = train.sampler.sample(idx, None).to(dev)
    jitter = torch.rand((len(idx), S), device=dev)
    bundle = ffn.RayBundle(bundle.starts, bundle.directions, bundle.near, bundle.far, bundle.rays, S, True, jitter)
    g = torch.Generator(device=dev).manual_seed(1)
    gt_c, gt_a = torch.rand((len(idx), 3), device=dev, generator=g), torch.rand((len(idx),), device=dev, generator=g)

    def loss_of(out, dtype):
        return ((out.color - gt_c.to(dtype)) ** 2).mean() + 0.1 * ((out.alpha - gt_a.to(dtype)) ** 2).mean()

    def grads_of(m, kernels, dtype):
        m = m.train()
        rc = ffn.Raycaster(m)
        rc.train_kernels = kernels
        m.zero_grad()
        if dtype == torch.float64:
            mat = bundle.materialize()
            rs = ffn.RaySamples(mat.positions.double(), mat.view_directions.double(), mat.t_values.double(), mat.rays)
            loss_of(rc._render_torch(rs, False), dtype).backward()
        else:
            loss_of(rc.render(bundle, False), dtype).backward()
        return {n: p.grad.detach().double().clone() for n, p in m.named_parameters() if p.grad is not None}

    ref64 = grads_of(copy.deepcopy(model).double(), False, torch.float64)
    g32 = grads_of(copy.deepcopy(model), False, torch.float32)
    gk = grads_of(copy.deepcopy(model), True, torch.float32)
    out = {}
    for n, r in ref64.items():
        den = r.norm().item() + 1e-30
        a, b = gk[n], g32[n]
        out[n] = {"kernels_rel_l2": float((a - r).norm() / den), "fp32_autograd_rel_l2": float((b - r).norm() / den),
                  "kernels_cos": float((a.flatten() @ r.flatten()) / (a.norm() * r.norm() + 1e-30))}
    out["worst_kernels_rel_l2"] = max(v["kernels_rel_l2"] for v in out.values() if isinstance(v, dict))
    out["worst_fp32_rel_l2"] = max(v["fp32_autograd_rel_l2"] for v in out.values() if isinstance(v, dict))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20000)
    ap.add_argument("--report", type=int, default=2000)
    ap.add_argument("--res", type=int, default=128)
    ap.add_argument("--cams", type=int, default=40)
    ap.add_argument("--samples", type=int, default=64)
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--out", default="")
    ap.add_argument("--skip-reference", action="store_true")
    ap.add_argument("--scene", default="solid", choices=["frame", "smooth", "solid"])
    ap.add_argument("--train-only", action="store_true", help="fit arms only (scene scans)")
    ap.add_argument("--binary-alpha", action="store_true")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    tmp = tempfile.mkdtemp()
    data = os.path.join(tmp, "scene.npz")
    # 12 validation cameras x res^2 > 102,400 rays: the reference's _validate only works on its to_valid branch
    # (ray_caster.py:228-233: with fewer rays it indexes a Python set with a numpy array and raises)
    subprocess.run([sys.executable, os.path.join(ROOT, "tools", "make_synthetic_dataset.py"), data, "--resolution",
                    str(args.res), "--train", str(args.cams), "--val", "12", "--test", "2", "--steps", "192", "--scene", args.scene] + (["--binary-alpha"] if args.binary_alpha else []),
                   check=True,
                   stdout=sys.stderr)
    res = {"config": vars(args), "gpu": torch.cuda.get_device_name(0)}
    ours_model, res["ours"] = fit_arm(ffn, data, args, dev, "ours")
    ref_pkg = ref_model = None
    if refmod.available() and not args.skip_reference:
        ref_pkg = refmod.import_reference()
        ref_model, res["reference_cuda_fp32"] = fit_arm(ref_pkg, data, args, dev, "reference")
        res["final_val_psnr_gap_db"] = round(res["ours"]["val_psnr"][-1] - res["reference_cuda_fp32"]["val_psnr"][-1], 3)
    if args.train_only:
        print(json.dumps(res))
        return
    val = ffn.ImageDataset.load(data, "val", args.samples, True, False).to(dev)
    train = ffn.ImageDataset.load(data, "train", args.samples, True, True).to(dev)
    res["inference_parity_converged_model"] = inference_parity(ours_model, None, None, val, dev)
    res["inference_parity_converged_model_fp16x3"] = inference_parity(ours_model, None, None, val, dev, "fp16x3")
    if ref_model is not None:
        # the reference-trained weights in our model: the other direction of the drop-in
        m2 = ffn.NeRF(8, 256, 9, 10, 3, 4, [4], True).to(dev)
        m2.load_state_dict(ref_model.state_dict())
        res["inference_parity_reference_trained_model"] = inference_parity(m2, ref_pkg, ref_model.eval(), val, dev)
    # sigma ~ 350 stress (docs/ray_data.tsv reaches sigma 350): scale the opacity head of the converged model
    stress = copy.deepcopy(ours_model)
    base = res["inference_parity_converged_model"]["sigma_max"]
    k = max(1.0, 350.0 / max(base, 1e-3))
    with torch.no_grad():
        stress.opacity_out.weight.mul_(k)
        stress.opacity_out.bias.mul_(k)
    res["inference_parity_sigma350_stress"] = dict(inference_parity(stress, None, None, val, dev), opacity_scale=round(k, 2))
    res["inference_parity_sigma350_stress_fp16x3"] = inference_parity(stress, None, None, val, dev, "fp16x3")
    res["gradients_vs_fp64"] = gradient_check(ours_model, train, dev)
    print(json.dumps(res))
    if args.out:
        with open(args.out, "w") as f:
            json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
