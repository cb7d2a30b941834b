"""GPU parity tests (run with ``-m gpu`` on a B200): the CUDA path, called through the
C-ABI (``_lib.Net`` / ``Raycaster``), against (1) golden vectors produced by the real
reference and (2) the numpy oracle on seeded inputs.

Stated tolerances (fp16 tensor-core operands, fp32 everything else; "TF32-grade"):
  * pixels (colour, alpha): max-abs <= 2.5e-3 on the sharpened golden nets (sigma up to
    50, colour logits x6 -- adversarial dynamic range), <= 5e-4 on default-init nets;
  * raw network outputs: max-abs <= 4e-3 * max|raw|;
  * depth: identical except where two blend weights tie within rounding (<= 2 % of rays);
  * in-kernel sampling (t values, positions): bit-exact;
  * blend weights / compositing alone (fp32 end to end): <= 3e-7 abs.
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

PIX_TOL = 2.5e-3
DEV = "cuda:0"


def load(name):
    return np.load(os.path.join(GOLDEN, name))


def weights(g):
    return {k[2:]: g[k] for k in g.files if k.startswith("w.")}


def cuda(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


@pytest.fixture(scope="module")
def golden_nerf():
    g = load("nerf_render.npz")
    m = ffn.NeRF(8, 256, 9, 10, 3, 4, [4], True)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in weights(g).items()})
    return g, m.to(DEV).eval()


def test_library_loaded_and_counts_launches(golden_nerf):
    g, m = golden_nerf
    before = _lib.launch_count()
    with torch.no_grad():
        m(cuda(g["positions"][:4]).reshape(-1, 3), cuda(g["view_directions"][:4]).reshape(-1, 3))
    torch.cuda.synchronize()
    assert _lib.launch_count() > before


def test_nerf_raw_outputs_vs_reference(golden_nerf):
    g, m = golden_nerf
    with torch.no_grad():
        raw = m(cuda(g["positions"]).reshape(-1, 3), cuda(g["view_directions"]).reshape(-1, 3))
    err = np.abs(raw.cpu().numpy() - g["raw"]).max()
    assert err <= 4e-3 * np.abs(g["raw"]).max(), err


def test_nerf_hidden_layers_vs_oracle(golden_nerf):
    g, m = golden_nerf
    eng = engine.get_engine(m, torch.device(DEV))
    pos, view = g["positions"].reshape(-1, 3), g["view_directions"].reshape(-1, 3)
    hidden = oracle.nerf_hidden(weights(g), pos, view)
    for l, ref in enumerate(hidden):
        out = eng.net.debug_layer(cuda(pos), cuda(view), l).cpu().numpy()[:, :ref.shape[1]]
        assert np.abs(out - ref).max() <= 4e-3 * max(1.0, np.abs(ref).max()), l


def test_render_samples_vs_reference_pixels(golden_nerf):
    g, m = golden_nerf
    rc = ffn.Raycaster(m)
    samples = ffn.RaySamples(cuda(g["positions"]), cuda(g["view_directions"]), cuda(g["t_values"]),
                             cuda(g["idx"]))
    with torch.no_grad():
        out = rc.render(samples, True)
    rc.check_nan()
    assert np.abs(out.color.cpu().numpy() - g["color"]).max() <= PIX_TOL
    assert np.abs(out.alpha.cpu().numpy() - g["alpha"]).max() <= PIX_TOL
    assert (out.depth.cpu().numpy() != g["depth"]).mean() <= 0.02
    # and against the fp64 arbiter
    assert np.abs(out.color.cpu().numpy() - g["color64"]).max() <= PIX_TOL
    with torch.no_grad():
        assert rc.render(samples, False).depth is None


def test_render_rays_sampling_is_bit_exact_and_equals_samples_path(golden_nerf):
    g, m = golden_nerf
    s = load("sampler.npz")
    idx = s["idx"]
    near, far = s["near_far"][:, idx]
    eng = engine.get_engine(m, torch.device(DEV))
    lin = torch.linspace(0, 1, 64).to(DEV)
    color, alpha, depth, t_out = eng.net.render_rays(
        cuda(s["starts"][idx]), cuda(s["directions"][idx]), cuda(near), cuda(far), lin, cuda(s["u"]),
        True, 0, 0, 64, True, want_t=True)
    assert np.array_equal(t_out.cpu().numpy(), s["t_values"])          # ray_sampler.py:380-386, bitwise
    c2, a2, d2 = eng.net.render_samples(cuda(s["positions"]), cuda(s["view_directions"]), cuda(s["t_values"]), True)
    assert torch.equal(color, c2) and torch.equal(alpha, a2) and torch.equal(depth, d2)
    assert np.abs(color.cpu().numpy() - g["color"]).max() <= PIX_TOL
    # uniform (non-stratified) sampling
    _, _, _, t_u = eng.net.render_rays(cuda(s["starts"][idx]), cuda(s["directions"][idx]), cuda(near), cuda(far),
                                       lin, None, False, 0, 0, 64, False, want_t=True)
    assert np.array_equal(t_u.cpu().numpy(), s["t_uniform"])


def test_raycaster_with_raybundle_from_sampler(golden_nerf):
    g, m = golden_nerf
    s = load("sampler.npz")
    cams = [ffn.CameraInfo.create("c%d" % i, ffn.Resolution(24, 24), s["intrinsics"][i], s["extrinsics"][i])
            for i in range(3)]
    sampler = ffn.RaySampler(s["bounds"], cams, 64, stratified=True)
    torch.manual_seed(1234)
    bundle = sampler.sample(s["idx"].tolist(), None)     # host tables -> reference RNG stream
    rc = ffn.Raycaster(m)
    with torch.no_grad():
        out = rc.render(bundle.to(DEV), True).numpy()
    assert np.abs(out.color - g["color"]).max() <= PIX_TOL
    assert np.abs(out.alpha - g["alpha"]).max() <= PIX_TOL
    # batched_render / render_image plumbing
    res = rc.batched_render(bundle, 50, True)
    assert np.abs(res.color - g["color"]).max() <= PIX_TOL and res.depth.shape == (192,)
    sampler.stratified = False
    img = rc.render_image(sampler, 1, 100)
    assert img.shape == (24, 24, 3) and img.dtype == np.uint8
    valid = sampler._valid_for_camera(1)
    ref_s = oracle.sample_rays(sampler.starts[valid].numpy(), sampler.directions[valid].numpy(),
                               *sampler.near_far[:, valid].numpy(), 64)
    ref = oracle.render_rays(lambda p, v: oracle.nerf_forward(weights(g), p, v), ref_s, False)
    ref_img = sampler.to_image(1, ref.color, "RGB")
    assert np.abs(img.astype(int) - ref_img.astype(int)).max() <= 1     # uint8 truncation of <=2.5e-3


def test_device_resident_sampler_philox_jitter(golden_nerf):
    g, m = golden_nerf
    s = load("sampler.npz")
    cams = [ffn.CameraInfo.create("c%d" % i, ffn.Resolution(24, 24), s["intrinsics"][i], s["extrinsics"][i])
            for i in range(3)]
    sampler = ffn.RaySampler(s["bounds"], cams, 64, stratified=True).to(DEV)
    idx = sampler.to_valid(torch.arange(len(sampler), device=DEV))
    b = sampler.sample(idx, None)
    assert b.jitter is None and b.starts.is_cuda
    eng = engine.get_engine(m, torch.device(DEV))
    _, _, _, t = eng.net.render_rays(b.starts, b.directions, b.near, b.far, torch.linspace(0, 1, 64).to(DEV),
                                     None, True, 42, 0, 64, False, want_t=True)
    near, far = b.near[:, None], b.far[:, None]
    base = near + torch.linspace(0, 1, 64).to(DEV)[None] * (far - near)
    u = (t - base) / ((far - near) / 64)
    assert u.min() >= -1e-4 and u.max() < 1 + 1e-4                   # inside each stratum
    assert abs(u.mean().item() - 0.5) < 0.01 and abs(u.var().item() - 1 / 12) < 0.005
    _, _, _, t2 = eng.net.render_rays(b.starts, b.directions, b.near, b.far, torch.linspace(0, 1, 64).to(DEV),
                                      None, True, 43, 0, 64, False, want_t=True)
    assert not torch.equal(t, t2)                                       # seed matters


@pytest.mark.parametrize("name", ["mlp", "basic", "positional", "gaussian"])
def test_ffmlp_presets_vs_reference(name):
    g = load("ffmlp_%s.npz" % name)
    ctor = {"mlp": lambda: ffn.MLP(3, 4), "basic": lambda: ffn.BasicFourierMLP(3, 4),
            "positional": lambda: ffn.PositionalFourierMLP(3, 4, 5.5),
            "gaussian": lambda: ffn.GaussianFourierMLP(3, 4, 3.14)}[name]
    m = ctor()
    m.load_state_dict({k: torch.from_numpy(v) for k, v in weights(g).items()})
    m = m.to(DEV).eval()
    with torch.no_grad():
        raw = m(cuda(g["positions"]).reshape(-1, 3)).cpu().numpy()
        out = ffn.Raycaster(m).render(ffn.RaySamples(cuda(g["positions"]), None, cuda(g["t_values"]), None), True)
    assert np.abs(raw - g["raw"]).max() <= 4e-3 * max(1.0, np.abs(g["raw"]).max())
    assert np.abs(out.color.cpu().numpy() - g["color"]).max() <= PIX_TOL
    assert np.abs(out.alpha.cpu().numpy() - g["alpha"]).max() <= PIX_TOL
    assert (out.depth.cpu().numpy() != g["depth"]).mean() <= 0.05


@pytest.mark.parametrize("S", [1, 2, 3, 8, 16, 32, 64, 128, 48, 100, 129, 192, 256, 300])
@pytest.mark.parametrize("R", [1, 37, 300])
def test_ragged_shapes_vs_oracle(golden_nerf, R, S):
    """One fused launch for ANY samples-per-ray: rays aligned with the 128-row tiles (S | 128), rays that straddle
    tiles (finished through per-ray partials by the last CTA to arrive), partial last tiles."""
    if S == 1:
        pytest.skip("the reference's argmax over an empty weights[:, :-1] is undefined for S=1")
    g, m = golden_nerf
    rng = np.random.default_rng(R * 1000 + S)
    o = np.tile(np.array([[0.1, 0.2, -4.0]], np.float32), (R, 1))
    d = rng.normal(size=(R, 3)).astype(np.float32) * 0.12 + np.array([0, 0, 1], np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    near = rng.uniform(2.8, 3.3, R).astype(np.float32)
    far = rng.uniform(4.5, 5.2, R).astype(np.float32)
    u = rng.random((R, S), dtype=np.float32)
    ref_s = oracle.sample_rays(o, d, near, far, S, u=u)
    ref = oracle.render_rays(lambda p, v: oracle.nerf_forward(weights(g), p, v), ref_s, True)
    eng = engine.get_engine(m, torch.device(DEV))
    c, a, dep, t = eng.net.render_rays(cuda(o), cuda(d), cuda(near), cuda(far), torch.linspace(0, 1, S).to(DEV),
                                       cuda(u), True, 0, 0, S, True, want_t=True)
    assert np.array_equal(t.cpu().numpy(), ref_s.t_values)
    assert np.abs(c.cpu().numpy() - ref.color).max() <= PIX_TOL
    assert np.abs(a.cpu().numpy() - ref.alpha).max() <= PIX_TOL
    assert (dep.cpu().numpy() != ref.depth).mean() <= max(0.02, 1.5 / R)
    before = _lib.launch_count()
    c2, a2, d2 = eng.net.render_samples(cuda(ref_s.positions), cuda(ref_s.view_directions), cuda(ref_s.t_values), True)
    assert _lib.launch_count() - before == 1          # one kernel launch whatever S is
    assert torch.equal(c, c2) and torch.equal(a, a2) and torch.equal(dep, d2)


def test_empty_batch(golden_nerf):
    g, m = golden_nerf
    eng = engine.get_engine(m, torch.device(DEV))
    z3 = torch.zeros((0, 3), device=DEV)
    z1 = torch.zeros((0,), device=DEV)
    c, a, d, _ = eng.net.render_rays(z3, z3, z1, z1, torch.linspace(0, 1, 64).to(DEV), None, False, 0, 0, 64, True)
    assert c.shape == (0, 3) and a.shape == (0,) and d.shape == (0,)
    assert m(z3, z3).shape == (0, 4)


def test_blend_weights_and_composite_kernels():
    g = load("blend.npz")
    w = ffn.calculate_blend_weights(cuda(g["t"]), cuda(g["sigma"])).cpu().numpy()
    np.testing.assert_allclose(w, g["w"], rtol=0, atol=3e-7)
    k = load("ray_data_kat.npz")          # docs/ray_data.tsv known answer
    w = ffn.calculate_blend_weights(cuda(k["t"][None]), cuda(k["opacity"][None])).cpu().numpy()[0]
    ref = oracle.calculate_blend_weights(k["t"][None], k["opacity"][None])[0]
    np.testing.assert_allclose(w, ref, rtol=0, atol=3e-7)
    # stand-alone compositor on reference raw outputs
    n = load("nerf_render.npz")
    c, a, d, w = _lib.composite(cuda(n["raw"].reshape(192, 64, 4)), cuda(n["t_values"]), True, True)
    np.testing.assert_allclose(c.cpu().numpy(), n["color"], rtol=0, atol=2e-6)
    np.testing.assert_allclose(a.cpu().numpy(), n["alpha"], rtol=0, atol=2e-6)
    assert (d.cpu().numpy() != n["depth"]).mean() <= 0.011


def test_weight_update_triggers_repack(golden_nerf):
    g, m = golden_nerf
    m2 = ffn.NeRF(8, 256, 9, 10, 3, 4, [4], True)
    m2.load_state_dict(m.state_dict())
    m2 = m2.to(DEV).eval()
    pos, view = cuda(g["positions"][:8]).reshape(-1, 3), cuda(g["view_directions"][:8]).reshape(-1, 3)
    with torch.no_grad():
        a = m2(pos, view).clone()
        m2.color_out.bias.add_(0.5)                       # in-place, like an optimiser step
        b = m2(pos, view)
    np.testing.assert_allclose((b - a)[:, :3].cpu().numpy(), 0.5, atol=1e-5)
    assert torch.equal(a[:, 3], b[:, 3])


def test_nan_assertion_is_preserved(golden_nerf):
    g, m = golden_nerf
    m2 = ffn.NeRF(8, 256, 9, 10, 3, 4, [4], True)
    m2.load_state_dict(m.state_dict())
    m2 = m2.to(DEV).eval()
    with torch.no_grad():
        m2.opacity_out.bias.fill_(float("nan"))
    rc = ffn.Raycaster(m2)
    samples = ffn.RaySamples(cuda(g["positions"][:4]), cuda(g["view_directions"][:4]), cuda(g["t_values"][:4]), None)
    with torch.no_grad():
        rc.render(samples, True)
    with pytest.raises(AssertionError):
        rc.check_nan()
    rc.check_nan()                                        # flag is cleared after raising


def test_bf16_operand_mode(golden_nerf):
    g, m = golden_nerf
    eng = engine.get_engine(m, torch.device(DEV), "bf16")
    c, a, d = eng.net.render_samples(cuda(g["positions"]), cuda(g["view_directions"]), cuda(g["t_values"]), True)
    assert np.abs(c.cpu().numpy() - g["color"]).max() <= 2e-2
    engine.get_engine(m, torch.device(DEV), "fp16")


def test_default_init_net_tolerance():
    """Un-sharpened (default nn.Linear init) NeRF: the bench workload's weights."""
    torch.manual_seed(20080524)
    m = ffn.NeRF(8, 256, 9, 10, 3, 4, [4], True)
    params = {k: v.numpy() for k, v in m.state_dict().items()}
    m = m.to(DEV).eval()
    rng = np.random.default_rng(3)
    R, S = 512, 64
    o = np.tile(np.array([[0, 0, -4.0]], np.float32), (R, 1))
    d = rng.normal(size=(R, 3)).astype(np.float32) * 0.15 + np.array([0, 0, 1], np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    near, far = np.full(R, 3.0, np.float32), np.full(R, 5.0, np.float32)
    u = rng.random((R, S), dtype=np.float32)
    ref = oracle.render_rays(lambda p, v: oracle.nerf_forward(params, p, v), oracle.sample_rays(o, d, near, far, S, u=u))
    b = ffn.RayBundle(cuda(o), cuda(d), cuda(near), cuda(far), None, S, True, cuda(u))
    with torch.no_grad():
        out = ffn.Raycaster(m).render(b, True).numpy()
    assert np.abs(out.color - ref.color).max() <= 5e-4
    assert np.abs(out.alpha - ref.alpha).max() <= 5e-4


def test_full_size_properties(golden_nerf):
    """BASELINE size (2^20 rays x 64): properties that need no oracle run -- determinism,
    ray independence (any sub-batch / permutation gives bit-identical pixels), ranges."""
    g, m = golden_nerf
    R, S = 1 << 20, 64
    gen = torch.Generator(device=DEV).manual_seed(5)
    o = torch.tensor([0.0, 0.3, -4.0], device=DEV).repeat(R, 1)
    d = torch.nn.functional.normalize(torch.randn((R, 3), device=DEV, generator=gen) * 0.15
                                      + torch.tensor([0, 0, 1.0], device=DEV), dim=-1)
    near = 2.8 + 0.5 * torch.rand(R, device=DEV, generator=gen)
    far = near + 1 + torch.rand(R, device=DEV, generator=gen)
    eng = engine.get_engine(m, torch.device(DEV))
    lin = torch.linspace(0, 1, S).to(DEV)

    def run(sel=None, off=0):
        a = (o, d, near, far) if sel is None else (o[sel], d[sel], near[sel], far[sel])
        return eng.net.render_rays(*a, lin, None, False, 0, off, S, True)[:3]

    c, a, dep = run()
    c2, a2, dep2 = run()
    assert torch.equal(c, c2) and torch.equal(a, a2) and torch.equal(dep, dep2)
    assert c.min() >= 0 and c.max() <= 1 + 1e-5 and a.min() >= 0 and a.max() <= 1 + 1e-5
    assert (dep >= near - 1e-5).all() and (dep <= far + 1e-5).all()
    perm = torch.randperm(R, device=DEV, generator=gen)[:300000]
    cp, ap, dp = run(perm)
    assert torch.equal(cp, c[perm]) and torch.equal(ap, a[perm]) and torch.equal(dp, dep[perm])
    sl = slice(12345, 12345 + 777)
    cs, as_, ds = run(sl)
    assert torch.equal(cs, c[sl]) and torch.equal(as_, a[sl])
    assert eng.net.nan_flag() == 0


def test_engine_follows_changed_encoding_buffers():
    """a GaussianFourierMLP that has already rendered on CUDA and then receives ``load_state_dict``
    of a checkpoint with ANOTHER B matrix must render with the new B (the encoding buffers are baked into the C
    handle: the engine signature covers them and the handle is rebuilt)."""
    torch.manual_seed(1)
    m = ffn.GaussianFourierMLP(3, 4, 3.0).to(DEV).eval()
    x = torch.rand((4096, 3), device=DEV) * 2 - 1
    with torch.no_grad():
        first = m(x)
        torch.manual_seed(2)
        other = ffn.GaussianFourierMLP(3, 4, 5.0)
        m.load_state_dict(other.state_dict())
        got = m(x)
        want = m.forward_torch(x)
    assert not torch.allclose(first, got, atol=1e-2)
    assert (got - want).abs().max() <= 4e-3 * want.abs().max().clamp_min(1.0)
    # NeRF: other frequencies through load_state_dict
    n1 = ffn.NeRF(8, 256, 9, 10, 3, 4, [4], True).to(DEV).eval()
    v = torch.nn.functional.normalize(torch.randn((4096, 3), device=DEV), dim=-1)
    with torch.no_grad():
        n1(x, v)
        sd = ffn.NeRF(8, 256, 6, 10, 2, 4, [4], True).state_dict()
        n1.load_state_dict(sd)
        got, want = n1(x, v), n1.forward_torch(x, v)
    assert (got - want).abs().max() <= 4e-3 * want.abs().max().clamp_min(1.0)


def test_batched_render_draws_independent_jitter_per_batch():
    """in-kernel stratified jitter is Philox(seed, ray_offset + i, sample); the sub-batches of
    ``batched_render`` must not repeat the jitter of batch 0 (the reference draws torch.rand per ray)."""
    torch.manual_seed(3)
    m = ffn.NeRF(8, 256, 9, 10, 3, 4, [4], True)
    with torch.no_grad():
        m.opacity_out.bias += 3.0            # visible density, so that jitter moves the pixels
    m = m.to(DEV).eval()
    n = 512
    o = torch.tensor([0.1, 0.2, -4.0], device=DEV).repeat(2 * n, 1)
    d = torch.nn.functional.normalize(torch.tensor([0.02, -0.03, 1.0], device=DEV), dim=0).repeat(2 * n, 1)
    near, far = torch.full((2 * n,), 3.0, device=DEV), torch.full((2 * n,), 5.0, device=DEV)
    b = ffn.RayBundle(o, d, near, far, torch.arange(2 * n, device=DEV), 64, True, None, seed=11)
    rc = ffn.Raycaster(m)
    whole = rc.batched_render(b, 2 * n, False).color
    halves = rc.batched_render(b, n, False).color
    np.testing.assert_array_equal(whole, halves)               # batching does not change the draws ...
    assert np.abs(halves[:n] - halves[n:]).max() > 1e-4         # ... and identical rays in different batches differ
    assert np.abs(halves[0] - halves[1]).max() > 1e-5


def test_render_stream_pipeline_on_the_gpu(golden_nerf):
    """Raycaster.render_stream with host ray tables + pinned staging: same pixels as the step-by-step calls, results
    stay valid after later batches have reused the pinned buffers."""
    g, m = golden_nerf
    from bench import BOUNDS, make_cameras
    s = ffn.RaySampler(BOUNDS, make_cameras(ffn, 2, 64), 32, stratified=False)
    s.enable_pinned_staging(3000)
    valid = torch.nonzero(s.valid_mask).flatten()
    batches = [valid[i * 1000:(i + 1) * 1000 + 37 * i] for i in range(5)]
    rc = ffn.Raycaster(m)
    with torch.no_grad():
        rc.render(s.sample(batches[0], None).to(DEV), True)      # (packs the weight image)
    before = _lib.launch_count()
    got = list(rc.render_stream(s, batches, True))
    assert _lib.launch_count() - before == len(batches)
    with torch.no_grad():
        for r, b in zip(got, batches):
            want = rc.render(s.sample(b, None).to(DEV), True).numpy()
            np.testing.assert_array_equal(r.color, want.color)
            np.testing.assert_array_equal(r.depth, want.depth)


def test_fp16x3_precise_mode_matches_fp32(golden_nerf):
    """operand="fp16x3": hi + residual split of every tensor-core operand (three UMMAs per product).  On the sharpened
    golden net (sigma up to 50, colour logits x6) the fast fp16 mode is within 2.5e-3 of the reference's pixels; the
    precise mode must be within 5e-5 (fp32 SGEMM reassociation level), raw outputs within 2e-5 relative, depth equal."""
    g, m = golden_nerf
    import copy
    mp = copy.deepcopy(m)
    mp.ffn_operand = "fp16x3"
    s = ffn.RaySamples(*[cuda(g[k]) for k in ("positions", "view_directions", "t_values")], None)
    with torch.no_grad():
        before = _lib.launch_count()
        out = ffn.Raycaster(mp).render(s, True)
        raw = mp(cuda(g["positions"]).reshape(-1, 3), cuda(g["view_directions"]).reshape(-1, 3))
        fast = ffn.Raycaster(m).render(s, True)
    assert _lib.launch_count() > before
    err = np.abs(out.color.cpu().numpy() - g["color"]).max()
    err_a = np.abs(out.alpha.cpu().numpy() - g["alpha"]).max()
    err_fast = np.abs(fast.color.cpu().numpy() - g["color"]).max()
    print("fp16x3 colour %.2e alpha %.2e (fast mode %.2e)" % (err, err_a, err_fast))
    assert err <= 5e-5 and err_a <= 5e-5, (err, err_a)
    assert err < 0.2 * err_fast
    assert np.abs(raw.cpu().numpy() - g["raw"]).max() <= 2e-5 * np.abs(g["raw"]).max()
    assert (out.depth.cpu().numpy() != g["depth"]).mean() == 0
    # rays mode, a size with several tiles per cluster, samples-per-ray that straddles tiles
    R, S = 3000, 192
    rng = np.random.default_rng(5)
    o = np.tile(np.array([[0.1, 0.2, -4.0]], np.float32), (R, 1))
    d = rng.normal(size=(R, 3)).astype(np.float32) * 0.12 + np.array([0, 0, 1], np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    near, far = rng.uniform(2.8, 3.3, R).astype(np.float32), rng.uniform(4.5, 5.2, R).astype(np.float32)
    u = rng.random((R, S), dtype=np.float32)
    ref_s = oracle.sample_rays(o, d, near, far, S, u=u)
    ref = oracle.render_rays(lambda p, v: oracle.nerf_forward(weights(g), p, v), ref_s, True)
    eng = engine.get_engine(mp, torch.device(DEV))
    c, a, dep, _ = eng.net.render_rays(cuda(o), cuda(d), cuda(near), cuda(far), torch.linspace(0, 1, S).to(DEV), cuda(u),
                                       True, 0, 0, S, True)
    assert np.abs(c.cpu().numpy() - ref.color).max() <= 5e-5
    assert np.abs(a.cpu().numpy() - ref.alpha).max() <= 5e-5
    assert (dep.cpu().numpy() != ref.depth).mean() <= 1e-3


_FOLD_PROBE = r"""
import sys, numpy as np, torch
sys.path.insert(0, sys.argv[1])
import fourier_feature_nets_b200 as ffn
g = np.load(sys.argv[2])
m = ffn.NeRF(8, 256, 9, 10, 3, 4, [4], True)
m.load_state_dict({k[2:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("w.")})
m = m.to("cuda:0").eval()
with torch.no_grad():
    p = torch.from_numpy(g["positions"]).to("cuda:0").reshape(-1, 3)
    v = torch.from_numpy(g["view_directions"]).to("cuda:0").reshape(-1, 3)
    raw = m(p, v)
    # the weights change: the folded layer must follow (it is rebuilt lazily by the next render)
    m.bottleneck.weight.mul_(0.5)          # (in-place on the parameter: bumps its version counter -> re-pack)
    m.hidden_view.bias.add_(0.25)
    raw2 = m(p, v)
np.savez(sys.argv[3], raw=raw.cpu().numpy(), raw2=raw2.cpu().numpy())
"""


def test_folded_inference_program_equals_the_reference_layer_structure(tmp_path):
    """The inference kernel runs bottleneck . hidden_view as ONE folded layer (nerf_model.py:119-122 has no activation
    in between).  FFN_FOLD=0 runs the reference's two layers through the same kernel: raw outputs of the
    two programs agree to operand rounding, before and after a weight update (the folded image is rebuilt)."""
    import subprocess
    import sys
    from conftest import ROOT
    script = tmp_path / "probe.py"
    script.write_text(_FOLD_PROBE)
    outs = {}
    for fold in ("1", "0"):
        out = tmp_path / ("fold%s.npz" % fold)
        env = dict(os.environ, FFN_FOLD=fold)
        res = subprocess.run([sys.executable, str(script), ROOT, os.path.join(GOLDEN, "nerf_render.npz"), str(out)],
                             env=env, capture_output=True, text=True, timeout=600)
        assert res.returncode == 0, res.stderr[-2000:]
        outs[fold] = np.load(out)
    g = load("nerf_render.npz")
    scale = float(np.abs(g["raw"]).max())
    for key in ("raw", "raw2"):
        a, b = outs["1"][key], outs["0"][key]
        assert np.isfinite(a).all() and np.abs(a - b).max() <= 4e-3 * scale, (key, np.abs(a - b).max(), scale)
    # both follow the reference's fp32 outputs, and the update really changed them
    assert np.abs(outs["1"]["raw"].reshape(g["raw"].shape) - g["raw"]).max() <= 4e-3 * scale
    assert np.abs(outs["1"]["raw2"] - outs["1"]["raw"]).max() > 1e-2
