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
    # inverse-transform sampling amplifies a 1-ulp difference of the CDF (parallel scan here, sequential
    # cumsum in the reference) by (t_j - t_i) / (cdf_j - cdf_i), up to ~1e4 on this sharpened model: the
    # tolerance is 2e-5 + 4e-7 x the steepest inverse slope of the row
    tm = 0.5 * (oracle.linspace(near, far, n_f)[:, :-1] + oracle.linspace(near, far, n_f)[:, 1:])
    slope = (np.diff(tm, axis=1) / np.maximum(np.diff(cdf, axis=1), 1e-5)).max(1, keepdims=True)
    tol = 2e-5 + 4e-7 * slope
    t = _lib.focus_t(cuda(raw), cuda(near), cuda(far), cuda(near), cuda(far), lin_c, lin_u,
                     cuda(f["u_uniform"]), cuda(f["u_focus"]), True, 0, S).cpu().numpy()
    err = np.abs(t - f["t_values"])
    assert (err <= tol).all(), (err.max(), tol.max(), (err / tol).max())
    assert (np.diff(t, axis=1) >= 0).all()
    t = _lib.focus_t(cuda(raw[..., 3].copy()), cuda(near), cuda(far), cuda(near), cuda(far), lin_c, lin_u,
                     None, None, False, 0, S).cpu().numpy()
    err = np.abs(t - f["t_det"])
    assert (err <= tol).all(), (err.max(), tol.max(), (err / tol).max())


@pytest.mark.parametrize("S", [32, 64, 128, 200])
def test_focus_kernel_well_conditioned_vs_oracle(S):
    """Smooth opacity profile, annealed uniform segment, all sizes: t matches the oracle within the
    slope-aware tolerance (fp32 end to end)."""
    rng = np.random.default_rng(S)
    R, n_u, n_f = 150, S // 2, S - S // 2
    near = rng.uniform(2.5, 3.2, R).astype(np.float32)
    far = (near + rng.uniform(1.0, 2.5, R)).astype(np.float32)
    near_u = (near + 0.1).astype(np.float32)
    far_u = (far - 0.1).astype(np.float32)
    tc = oracle.linspace(near, far, n_f)
    centre = rng.uniform(0.3, 0.7, (R, 1)).astype(np.float32)
    x = (tc - near[:, None]) / (far - near)[:, None]
    # sigma stays >= ~0.5: no near-flat CDF bins, i.e. away from the reference's `denominator < 1e-5 -> 1`
    # special case (ray_sampler.py:345-347), whose discontinuity makes bin flips visible
    raw_sigma = (1.5 - 3 * (x - centre) ** 2 + rng.normal(size=x.shape) * 0.3).astype(np.float32)
    cdf = oracle.determine_cdf(tc, oracle.softplus(raw_sigma))
    ju, uf = rng.random((R, n_u), dtype=np.float32), rng.random((R, n_f), dtype=np.float32)
    o = np.zeros((R, 3), np.float32)
    d = np.tile(np.array([[0, 0, 1]], np.float32), (R, 1))
    # oracle: uniform part on the annealed segment, focus part on the raw segment (ray_sampler.py:305,373-392)
    uni = oracle.linspace(near_u, far_u, n_u) + ju * ((far_u - near_u) / np.float32(n_u))[:, None]
    foc = oracle.sample_t_values(near, far, cdf, n_f, uf)
    ref = np.sort(np.concatenate([uni.astype(np.float32), foc], 1), axis=1)
    t = _lib.focus_t(cuda(raw_sigma), cuda(near), cuda(far), cuda(near_u), cuda(far_u),
                     torch.linspace(0, 1, n_f).to(DEV), torch.linspace(0, 1, n_u).to(DEV), cuda(ju), cuda(uf),
                     True, 0, S).cpu().numpy()
    tm = 0.5 * (tc[:, :-1] + tc[:, 1:])
    slope = (np.diff(tm, axis=1) / np.maximum(np.diff(cdf, axis=1), 1e-5)).max(1, keepdims=True)
    err = np.abs(t - ref)
    assert (err <= 2e-5 + 4e-7 * slope).all(), err.max()
    # in-kernel Philox: sorted, inside the segments, different per seed
    t1 = _lib.focus_t(cuda(raw_sigma), cuda(near), cuda(far), cuda(near_u), cuda(far_u),
                      torch.linspace(0, 1, n_f).to(DEV), torch.linspace(0, 1, n_u).to(DEV), None, None, True, 7, S)
    t2 = _lib.focus_t(cuda(raw_sigma), cuda(near), cuda(far), cuda(near_u), cuda(far_u),
                      torch.linspace(0, 1, n_f).to(DEV), torch.linspace(0, 1, n_u).to(DEV), None, None, True, 8, S)
    assert (t1[:, 1:] >= t1[:, :-1]).all() and not torch.equal(t1, t2)
    # stratified samples may overshoot the segment end by one stratum (ray_sampler.py:380-386)
    hi = np.maximum(far, far_u + (far_u - near_u) / n_u)
    assert (t1.min(1)[0].cpu().numpy() >= near - 1e-5).all() and (t1.max(1)[0].cpu().numpy() <= hi + 1e-5).all()


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
