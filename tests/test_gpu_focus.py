"""GPU tests of the hierarchical ("focus") sampling path (``-m gpu``), SURVEY.md section 8 a-6.

* ``ffn_focus_t`` alone, fed the coarse opacities the REFERENCE computed (tests/golden/focus.npz): the
  merged, sorted t values match the reference's ``RaySampler.sample`` to 2e-5 (fp32 end to end).
* the lazy path (coarse sigma from the fp16 tensor-core MLP): t within 2e-3 of the oracle, rendered pixels
  within the usual pixel tolerance of an oracle render on the oracle's samples.
"""
import os

import numpy as np
import pytest
import torch

import oracle
import fourier_feature_nets_b200 as ffn
from fourier_feature_nets_b200 import _lib, engine
from conftest import GOLDEN

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def cuda(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def golden_model():
    g = np.load(os.path.join(GOLDEN, "nerf_render.npz"))
    w = {k[2:]: g[k] for k in g.files if k.startswith("w.")}
    m = ffn.NeRF(8, 256, 9, 10, 3, 4, [4], True)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in w.items()})
    return w, m.to(DEV).eval()


def coarse_raw_oracle(w, f, n_c):
    near, far = f["near_far"]
    t = oracle.linspace(near, far, n_c)
    pos = f["starts"][:, None, :] + t[..., None] * f["directions"][:, None, :]
    dirs = np.repeat(f["directions"][:, None, :], n_c, 1)
    raw = oracle.nerf_forward(w, pos.reshape(-1, 3).astype(np.float32), dirs.reshape(-1, 3))
    return raw.reshape(len(near), n_c, 4)


def test_focus_kernel_reproduces_reference_sampling():
    f = np.load(os.path.join(GOLDEN, "focus.npz"))       # reference RaySampler with opacity_model, S = 32
    w, _ = golden_model()
    S, n_u, n_f = 32, 16, 16
    near, far = f["near_far"]
    raw = coarse_raw_oracle(w, f, n_f)
    # sanity: the oracle's cdf from these opacities is the reference's table
    cdf = oracle.determine_cdf(oracle.linspace(near, far, n_f), oracle.softplus(raw[..., 3]))
    np.testing.assert_allclose(cdf, f["cdfs"], atol=5e-5)
    lin_c, lin_u = torch.linspace(0, 1, n_f).to(DEV), torch.linspace(0, 1, n_u).to(DEV)
    t = _lib.focus_t(cuda(raw), cuda(near), cuda(far), cuda(near), cuda(far), lin_c, lin_u,
                     cuda(f["u_uniform"]), cuda(f["u_focus"]), True, 0, S).cpu().numpy()
    np.testing.assert_allclose(t, f["t_values"], rtol=0, atol=2e-5)
    assert (np.diff(t, axis=1) >= 0).all()
    t = _lib.focus_t(cuda(raw[..., 3].copy()), cuda(near), cuda(far), cuda(near), cuda(far), lin_c, lin_u,
                     None, None, False, 0, S).cpu().numpy()
    np.testing.assert_allclose(t, f["t_det"], rtol=0, atol=2e-5)


@pytest.mark.parametrize("S", [32, 128, 100])
def test_lazy_focus_render_vs_oracle(S):
    w, m = golden_model()
    f = np.load(os.path.join(GOLDEN, "focus.npz"))
    near, far = f["near_far"]
    R = len(near)
    n_u, n_f = S // 2, S - S // 2
    rng = np.random.default_rng(S)
    ju, uf = rng.random((R, n_u), dtype=np.float32), rng.random((R, n_f), dtype=np.float32)
    # oracle: coarse cdf -> samples -> render
    raw_c = coarse_raw_oracle(w, f, n_f)
    cdf = oracle.determine_cdf(oracle.linspace(near, far, n_f), oracle.softplus(raw_c[..., 3]))
    ref_s = oracle.sample_rays(f["starts"], f["directions"], near, far, S, u=ju, cdf=cdf, u_focus=uf,
                               focus_stratified=True)
    ref = oracle.render_rays(lambda p, v: oracle.nerf_forward(w, p, v), ref_s, True)
    b = ffn.FocusBundle(cuda(f["starts"]), cuda(f["directions"]), cuda(near), cuda(far), cuda(near), cuda(far),
                        None, S, True, cuda(ju), cuda(uf), 0, m)
    t = b.focus_t().cpu().numpy()
    assert (np.diff(t, axis=1) >= 0).all()
    assert np.abs(t - ref_s.t_values).max() <= 2e-3
    with torch.no_grad():
        out = ffn.Raycaster(m).render(b, True).numpy()
    # sample positions differ by <= 2e-3 (fp16 coarse pass), so allow a little more than the pixel tolerance
    assert np.abs(out.color - ref.color).max() <= 6e-3
    assert np.abs(out.alpha - ref.alpha).max() <= 6e-3
    # rendering exactly on the kernel's own t values is within the usual tolerance
    samples = oracle.OracleSamples((f["starts"][:, None, :] + t[..., None] * f["directions"][:, None, :]).astype(np.float32),
                                   np.repeat(f["directions"][:, None, :], S, 1), t)
    ref2 = oracle.render_rays(lambda p, v: oracle.nerf_forward(w, p, v), samples, True)
    assert np.abs(out.color - ref2.color).max() <= 2.5e-3
    assert (out.depth != ref2.depth).mean() <= 0.05


def test_sampler_with_opacity_model_is_lazy_and_renders():
    w, m = golden_model()
    s = np.load(os.path.join(GOLDEN, "sampler.npz"))
    cams = [ffn.CameraInfo.create("c%d" % i, ffn.Resolution(24, 24), s["intrinsics"][i], s["extrinsics"][i])
            for i in range(3)]
    sampler = ffn.RaySampler(s["bounds"], cams, 64, stratified=False, opacity_model=m)
    assert sampler.lazy_focus and not hasattr(sampler, "cdfs")
    rc = ffn.Raycaster(m)
    img = rc.render_image(sampler, 0, 200)               # orbit_video.py's loop body
    assert img.shape == (24, 24, 3) and img.dtype == np.uint8 and img.max() > 0
    bundle = sampler.rays_for_camera(0)
    assert isinstance(bundle, ffn.FocusBundle)
    mat = bundle.to(DEV).materialize()
    assert mat.positions.shape[1:] == (64, 3) and (mat.t_values[:, 1:] >= mat.t_values[:, :-1]).all()
    # training through a FocusBundle (samples mode under autograd)
    m.train()
    out = rc.render(bundle.subset(range(64)).to(DEV), True)
    out.color.sum().backward()
    assert m.layers[0].weight.grad is not None and torch.isfinite(m.layers[0].weight.grad).all()
