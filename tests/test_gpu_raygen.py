"""GPU tests (``-m gpu``) of the device-side ray tables (``ffn_generate_rays``) against the
oracle's restatement of camera_info.py:66-109 / ray_sampler.py:202-232.

Tolerances (fp32; the device evaluates the 4-term dot products with separate multiplies and adds, numpy's BLAS
may fuse or reorder them): directions 2e-6 absolute, near/far 2e-5 relative to the scene scale on rays that are
valid on both sides, and the valid masks may differ only on rays whose segment is shorter than 1e-4."""
import numpy as np
import pytest
import torch

import fourier_feature_nets_b200 as ffn
from oracle import ffn_oracle as orc

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def make_cameras(n, res, seed=0, distance=4.0):
    rng = np.random.default_rng(seed)
    f = 0.5 * res / np.tan(np.radians(20.0))
    K = np.array([[f, 0, res / 2], [0, f, res / 2], [0, 0, 1]], np.float32)
    cams = []
    for i in range(n):
        eye = rng.normal(size=3)
        eye = (eye / np.linalg.norm(eye) * distance).astype(np.float32)
        fwd = -eye / np.linalg.norm(eye)
        right = np.cross(np.array([0, 1, 0], np.float32), fwd)
        right /= np.linalg.norm(right)
        up = np.cross(fwd, right)
        E = np.eye(4, dtype=np.float32)
        E[:3, 0], E[:3, 1], E[:3, 2], E[:3, 3] = right, up, fwd, eye
        cams.append(ffn.CameraInfo("cam%d" % i, ffn.Resolution(res, res), K, E))
    return cams


@pytest.mark.parametrize("res,ncam", [(40, 3), (33, 2), (128, 1)])
def test_ray_tables_match_the_oracle(res, ncam):
    cams = make_cameras(ncam, res, seed=res)
    bounds = np.diag([2.0, 2.0, 2.0, 1.0]).astype(np.float32)
    sampler = ffn.RaySampler(bounds, cams, 64, device=DEV)
    assert sampler.starts.is_cuda and sampler.valid_mask.dtype == torch.bool
    xs, ys = np.meshgrid(np.arange(res), np.arange(res))
    points = np.stack([xs, ys], -1).reshape(-1, 2)
    o_ref, d_ref, nf_ref, ok_ref = [], [], [], []
    for cam in cams:
        o, d = orc.raycast(cam.intrinsics, cam.extrinsics, points)
        nf, ok = orc.near_far(bounds, o, d)
        o_ref.append(o), d_ref.append(d), nf_ref.append(nf), ok_ref.append(ok)
    o_ref, d_ref = np.concatenate(o_ref), np.concatenate(d_ref)
    nf_ref, ok_ref = np.concatenate(nf_ref, -1), np.concatenate(ok_ref)
    o, d = sampler.starts.cpu().numpy(), sampler.directions.cpu().numpy()
    nf, ok = sampler.near_far.cpu().numpy(), sampler.valid_mask.cpu().numpy()
    assert np.abs(o - o_ref).max() <= 1e-6
    assert np.abs(d - d_ref).max() <= 2e-6
    both = ok & ok_ref
    assert both.sum() > 0.3 * len(ok)
    assert np.abs(nf[:, both] - nf_ref[:, both]).max() <= 2e-5 * 4.0
    differ = ok != ok_ref
    if differ.any():
        seg = np.abs(nf_ref[1, differ] - nf_ref[0, differ])
        assert seg.max() <= 1e-4, seg.max()
    # the host path of the package (bit-identical to the reference, tests/test_host_logic.py) agrees as well
    host = ffn.RaySampler(bounds, cams, 64)
    assert np.abs(host.directions.numpy() - d).max() <= 2e-6
    assert (host.valid_mask.numpy() != ok).mean() <= 1e-3


def test_full_size_tables_have_the_geometric_properties():
    """lego_400-shaped: 20 cameras of 400x400 = 3.2 M rays; size-independent checks."""
    res = 400
    cams = make_cameras(20, res, seed=7, distance=4.12)
    bounds = np.diag([2.0, 2.0, 2.0, 1.0]).astype(np.float32)
    s = ffn.RaySampler(bounds, cams, 64, device=DEV)
    n = 20 * res * res
    assert s.starts.shape == (n, 3) and s.near_far.shape == (2, n)
    assert (s.directions.norm(dim=-1) - 1).abs().max().item() <= 2e-6
    near, far = s.near_far
    ok = s.valid_mask
    assert torch.equal(ok, near < far)
    frac = ok.float().mean().item()
    assert 0.5 < frac < 0.95, frac
    assert near[ok].min().item() >= 0.1
    # entry and exit points of the hits lie on the surface of the [-1,1]^3 box
    for t in (near[ok], far[ok]):
        p = s.starts[ok] + t.unsqueeze(-1) * s.directions[ok]
        assert (p.abs().max(dim=-1)[0] - 1).abs().max().item() <= 2e-5
    # origins are the camera centres
    for c in (0, 7, 19):
        lo = c * res * res
        assert torch.equal(s.starts[lo:lo + res * res], torch.from_numpy(cams[c].position).to(DEV).expand(res * res, 3))


def test_device_sampler_renders_an_image():
    cams = make_cameras(2, 48, seed=3)
    bounds = np.diag([2.0, 2.0, 2.0, 1.0]).astype(np.float32)
    torch.manual_seed(0)
    model = ffn.NeRF(8, 256, 9, 10, 3, 4, [4], True).to(DEV)
    dev_s = ffn.RaySampler(bounds, cams, 64, device=DEV)
    host_s = ffn.RaySampler(bounds, cams, 64).to(DEV)
    rc = ffn.Raycaster(model)
    a = rc.render_image(dev_s, 1, 4096)
    b = rc.render_image(host_s, 1, 4096)
    assert a.shape == (48, 48, 3) and a.dtype == np.uint8
    # identical up to rays that graze the box (mask flips) and uint8 truncation of ~1e-6 differences
    assert (np.abs(a.astype(int) - b.astype(int)) > 1).mean() <= 2e-3
