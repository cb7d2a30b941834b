"""GPU tests (``-m gpu``) of ``Voxels`` (voxels_model.py:9-57) on the device and of hierarchical sampling with a
coarse opacity model the MLP engine cannot evaluate.

Tolerances: the interpolation is fp32 end to end.  Against the fixture computed by the reference on the CPU the
kernel differs only in the rounding of the 8-term weighted sum (2e-6 x max|out|).  torch's CUDA ``positions / scale``
multiplies by the reciprocal where the kernel (like the CPU reference) divides: a 1-ulp coordinate difference,
amplified by side/2 cells per unit, so against torch-on-GPU the bound carries a 2e-7 x side term."""
import os

import numpy as np
import pytest
import torch

import oracle
import fourier_feature_nets_b200 as ffn
from fourier_feature_nets_b200 import _lib
from conftest import GOLDEN

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_voxels_kernel_matches_the_reference_fixture():
    g = np.load(os.path.join(GOLDEN, "checkpoints.npz"))
    vox = ffn.load_model(os.path.join(GOLDEN, "ref_voxels_small.pt")).to(DEV)
    before = _lib.launch_count()
    with torch.no_grad():
        out = vox(torch.from_numpy(g["pos_vox"]).to(DEV))
    assert _lib.launch_count() == before + 1            # the CUDA kernel ran, not grid_sample
    scale = np.abs(g["out_vox"]).max()
    assert np.abs(out.cpu().numpy() - g["out_vox"]).max() <= 2e-6 * scale


@pytest.mark.parametrize("side,n", [(1, 100), (2, 1000), (33, 100000), (128, 1 << 20)])
def test_voxels_kernel_matches_grid_sample(side, n):
    torch.manual_seed(side)
    vox = ffn.Voxels(side, 1.3).to(DEV)
    with torch.no_grad():
        vox.voxels.normal_()
        vox.bias.normal_()
    pos = (torch.rand((n, 3), device=DEV) * 3.2 - 1.6)
    with torch.no_grad():
        out = vox(pos)
        ref = vox.forward_torch(pos)
    tol = (2e-6 + 2e-7 * side) * max(1.0, ref.abs().max().item())
    assert (out - ref).abs().max().item() <= tol
    # parameters changed in place -> the channels-last copy is refreshed
    with torch.no_grad():
        vox.voxels.mul_(2.0)
        assert (vox(pos) - vox.forward_torch(pos)).abs().max().item() <= 2 * tol
    # with gradients the differentiable definition is used
    loss = vox(pos[:64]).square().sum()
    loss.backward()
    assert vox.voxels.grad is not None and vox.voxels.grad.abs().sum().item() > 0


def look_at(pos, res):
    pos = np.asarray(pos, np.float32)
    fwd = -pos / np.linalg.norm(pos)
    right = np.cross(np.array([0, 1, 0], np.float32), fwd)
    right /= np.linalg.norm(right)
    up = np.cross(fwd, right)
    E = np.eye(4, dtype=np.float32)
    E[:3, 0], E[:3, 1], E[:3, 2], E[:3, 3] = right, up, fwd, pos
    f = 0.5 * res / np.tan(np.radians(20.0))
    K = np.array([[f, 0, res / 2], [0, f, res / 2], [0, 0, 1]], np.float32)
    return ffn.CameraInfo("c", ffn.Resolution(res, res), K, E)


def blob_voxels(side=24):
    """A smooth density blob: sigma logits well away from the flat-CDF regime on rays through the middle."""
    vox = ffn.Voxels(side, 1.0)
    ax = (torch.arange(side, dtype=torch.float32) + 0.5) / side * 2 - 1
    z, y, x = torch.meshgrid(ax, ax, ax, indexing="ij")
    r2 = (x - 0.1) ** 2 + (y + 0.05) ** 2 + z ** 2
    with torch.no_grad():
        vox.voxels[0, 3] = 6.0 * torch.exp(-r2 / 0.15) + 1.0
        vox.voxels[0, :3] = torch.stack([x, y, z]) * 2
    return vox


@pytest.mark.parametrize("S", [32, 64])
def test_focus_sampling_with_a_voxel_opacity_model(S):
    """Per-batch device path (FocusBundle) vs the reference-style host path (constructor CDF table) of the same
    sampler, same torch.rand draws; fine model = the golden NeRF."""
    res = 20
    cams = [look_at((0.3, 0.4, -4.0), res), look_at((3.6, 0.5, 1.5), res)]
    bounds = np.diag([2, 2, 2, 1]).astype(np.float32)
    vox_host = blob_voxels()
    vox_dev = blob_voxels().to(DEV)
    host = ffn.RaySampler(bounds, cams, S, True, vox_host, 4096)
    dev = ffn.RaySampler(bounds, cams, S, True, vox_dev, 4096)
    assert not host.lazy_focus and dev.lazy_focus and not hasattr(dev, "cdfs")
    idx = host.to_valid(list(range(len(host))))[::3]
    torch.manual_seed(5)
    ref = host.sample(idx, None)
    torch.manual_seed(5)
    dev.device_jitter = False            # host torch.rand draws, the reference's order
    bundle = dev.sample(idx, None)
    assert isinstance(bundle, ffn.FocusBundle)
    t = bundle.to(DEV).focus_t().cpu().numpy()
    t_ref = ref.t_values.numpy()
    assert np.all(np.diff(t, axis=1) >= 0)
    # slope-aware tolerance as in tests/test_gpu_focus.py
    n_f = S - S // 2
    near, far = host.near_far[:, idx].numpy()
    cdf = host.cdfs[idx].numpy()
    tm = 0.5 * (oracle.linspace(near, far, n_f)[:, :-1] + oracle.linspace(near, far, n_f)[:, 1:])
    slope = (np.diff(tm, axis=1) / np.maximum(np.diff(cdf, axis=1), 1e-5)).max(1, keepdims=True)
    tol = 2e-5 + 4e-7 * slope
    bad = np.abs(t - t_ref) > tol
    assert bad.mean() <= 2e-3, (bad.mean(), np.abs(t - t_ref).max())
    # and the bundle renders through the fused kernel
    g = np.load(os.path.join(GOLDEN, "nerf_render.npz"))
    fine = ffn.NeRF(8, 256, 9, 10, 3, 4, [4], True)
    fine.load_state_dict({k[2:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("w.")})
    rc = ffn.Raycaster(fine.to(DEV).eval())
    with torch.no_grad():
        a = rc.render(bundle.to(DEV), True)
        b = rc.render(ref.to(DEV), True)
    assert (a.color - b.color).abs().max().item() <= 5e-3


def test_voxels_render_and_fit_on_the_gpu(tmp_path):
    """train_voxels.py's model on a CUDA device: inference = Voxels.forward (ffn_voxels_forward) + ffn_composite, equal
    to the PyTorch definition within fp32 rounding (1e-5); training = autograd over grid_sample with the fused loss and
    ClipAdam, PSNR improves."""
    import os
    import subprocess
    import sys
    from conftest import ROOT
    torch.manual_seed(0)
    model = ffn.Voxels(24, 2.0).to(DEV)
    with torch.no_grad():
        model.voxels.normal_(0, 1.5)
    R, S = 500, 48
    g = torch.Generator().manual_seed(1)
    o = torch.tensor([0.1, 0.2, -4.0]).repeat(R, 1)
    d = torch.nn.functional.normalize(torch.randn((R, 3), generator=g) * 0.12 + torch.tensor([0, 0, 1.0]), dim=-1)
    near = 2.8 + 0.5 * torch.rand(R, generator=g)
    bundle = ffn.RayBundle(o, d, near, near + 2.0, torch.arange(R), S, True, torch.rand((R, S), generator=g)).to(DEV)
    rc = ffn.Raycaster(model.eval())
    before = _lib.launch_count()
    with torch.no_grad():
        out = rc.render(bundle, True)
        ref = rc._render_torch(bundle.materialize(), True)
    assert _lib.launch_count() - before >= 2                     # voxel gather + compositing kernels
    assert (out.color - ref.color).abs().max().item() <= 1e-5 and (out.alpha - ref.alpha).abs().max().item() <= 1e-5
    assert (out.depth != ref.depth).float().mean().item() <= 0.01
    rc.check_nan()
    data = str(tmp_path / "toy.npz")
    subprocess.run([sys.executable, os.path.join(ROOT, "tools", "make_synthetic_dataset.py"), data,
                    "--resolution", "32", "--train", "8", "--val", "2", "--test", "1", "--steps", "64"],
                   check=True, capture_output=True, timeout=300)
    train = ffn.ImageDataset.load(data, "train", 32, True, True)
    val = ffn.ImageDataset.load(data, "val", 32, True, False)
    model = ffn.Voxels(16, 2 / train.sampler.bounds[0, 0]).to(DEV)
    log = ffn.Raycaster(model).fit(train, val, 512, 0.01, 40, 0, 20, 0.9, 25000, 0.0, [])
    assert len(log) >= 2 and log[-1].val_psnr > log[0].val_psnr, [e.val_psnr for e in log]
