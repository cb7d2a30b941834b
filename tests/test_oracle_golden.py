"""Pins the numpy oracle against vectors produced by the REAL reference
(tests/golden/make_golden.py) and the reference's only numeric fixture
(docs/ray_data.tsv)."""
import os

import numpy as np
import pytest

import oracle
from conftest import GOLDEN


def load(name):
    return np.load(os.path.join(GOLDEN, name))


def weights(npz):
    return {k[2:]: npz[k] for k in npz.files if k.startswith("w.")}


def test_torch_linspace_bit_exact():
    g = load("sampler.npz")
    for n in (64, 63, 7):
        assert np.array_equal(oracle.torch_linspace(0, 1, n), g["linspace_%d" % n])


def test_raycast_and_near_far():
    g = load("sampler.npz")
    o, d = oracle.raycast(g["intrinsics"][1], g["extrinsics"][1], g["points"])
    np.testing.assert_allclose(o, g["cam1_starts"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(d, g["cam1_dirs"], rtol=0, atol=2e-6)
    nf, valid = oracle.near_far(g["bounds"], g["starts"], g["directions"])
    assert np.array_equal(np.nonzero(~valid)[0], g["invalid"])
    assert np.array_equal(nf[:, valid], g["near_far"][:, valid])


def test_sample_stratified_bit_exact():
    g = load("sampler.npz")
    idx = g["idx"]
    near, far = g["near_far"][:, idx]
    s = oracle.sample_rays(g["starts"][idx], g["directions"][idx], near, far, 64, u=g["u"])
    assert np.array_equal(s.t_values, g["t_values"])
    assert np.array_equal(s.positions, g["positions"])
    assert np.array_equal(s.view_directions, g["view_directions"])
    s = oracle.sample_rays(g["starts"][idx], g["directions"][idx], near, far, 64, u=None)
    assert np.array_equal(s.t_values, g["t_uniform"])


def test_sample_anneal_bit_exact():
    g = load("sampler.npz")
    idx = g["idx"]
    near, far = g["near_far"][:, idx]
    s = oracle.sample_rays(g["starts"][idx], g["directions"][idx], near, far, 64,
                           u=g["u_anneal"], step=500, num_anneal_steps=2000, anneal_start=0.2)
    assert np.array_equal(s.t_values, g["t_anneal"])
    assert np.array_equal(s.positions, g["pos_anneal"])


def test_blend_weights():
    g = load("blend.npz")
    w = oracle.calculate_blend_weights(g["t"], g["sigma"])
    # exp() differs by 1 ulp between numpy and ATen; 1-alpha amplifies it where alpha~1
    np.testing.assert_allclose(w, g["w"], rtol=2e-6, atol=1.2e-7)


def test_ray_data_tsv_known_answer():
    """docs/ray_data.tsv: T column == inclusive transmittance."""
    g = load("ray_data_kat.npz")
    t, sig, T = g["t"][None], g["opacity"][None], g["T"]
    deltas = np.concatenate([t[:, 1:] - t[:, :-1], np.full((1, 1), 1e10, np.float32)], -1)
    alpha = 1 - np.exp(-(sig * deltas))
    w = oracle.calculate_blend_weights(t, sig)[0]
    with np.errstate(divide="ignore", invalid="ignore"):
        trans_excl = np.where(alpha[0] > 1e-20, w / alpha[0], np.nan)
    ok = ~np.isnan(trans_excl[1:])
    # exclusive transmittance at i+1 == inclusive (TSV) at i; the TSV holds 3-8 printed digits
    np.testing.assert_allclose(trans_excl[1:][ok], T[:-1][ok], rtol=2e-3, atol=2e-5)


def test_nerf_forward_and_render():
    g = load("nerf_render.npz")
    p = weights(g)
    raw = oracle.nerf_forward(p, g["positions"].reshape(-1, 3), g["view_directions"].reshape(-1, 3))
    # fp32 GEMM summation order differs between MKL (reference) and numpy's BLAS
    err = np.abs(raw - g["raw"]).max()
    scale = np.abs(g["raw"]).max()
    assert err <= 2e-5 * max(1.0, scale), (err, scale)
    out = oracle.render(raw.reshape(192, 64, 4), g["t_values"], True)
    np.testing.assert_allclose(out.color, g["color"], rtol=0, atol=2e-5)
    np.testing.assert_allclose(out.alpha, g["alpha"], rtol=0, atol=2e-5)
    assert (out.depth != g["depth"]).mean() <= 0.01
    # with the reference's own raw outputs the compositing is near bit-exact
    out = oracle.render(g["raw"].reshape(192, 64, 4), g["t_values"], True)
    np.testing.assert_allclose(out.color, g["color"], rtol=0, atol=5e-7)
    np.testing.assert_allclose(out.alpha, g["alpha"], rtol=0, atol=5e-7)
    assert np.array_equal(out.depth, g["depth"])


def test_nerf_forward_fp64_arbiter():
    g = load("nerf_render.npz")
    p = {k: v.astype(np.float64) for k, v in weights(g).items()}
    raw = oracle.nerf_forward(p, g["positions"].reshape(-1, 3).astype(np.float64),
                              g["view_directions"].reshape(-1, 3).astype(np.float64))
    np.testing.assert_allclose(raw, g["raw64"], rtol=0, atol=1e-9)


@pytest.mark.parametrize("name", ["mlp", "basic", "positional", "gaussian"])
def test_ffmlp_presets(name):
    g = load("ffmlp_%s.npz" % name)
    p = weights(g)
    a = p.pop("a_values", None)
    b = p.pop("b_values", None)
    raw = oracle.ffmlp_forward(p, g["positions"].reshape(-1, 3), a, b)
    err = np.abs(raw - g["raw"]).max()
    assert err <= 3e-5 * max(1.0, np.abs(g["raw"]).max()), err
    out = oracle.render(raw.reshape(64, 64, 4), g["t_values"], True)
    np.testing.assert_allclose(out.color, g["color"], rtol=0, atol=2e-5)
    np.testing.assert_allclose(out.alpha, g["alpha"], rtol=0, atol=2e-5)


def test_positional_b_values_matches_reference_buffer():
    g = load("ffmlp_positional.npz")
    # non-integer exponents: powf differs by <=1 ulp between libm and ATen; the
    # product always reads b_values from the model's own buffer, never recomputes it
    np.testing.assert_allclose(oracle.positional_b_values(5.5, 256, 3), g["w.b_values"], rtol=1.2e-7)
    n = load("nerf_render.npz")
    assert np.array_equal(oracle.nerf_encoding_matrix(9, 10), n["w.pos_encoding"])
    assert np.array_equal(oracle.nerf_encoding_matrix(3, 4), n["w.view_encoding"])


def test_focus_sampling():
    g = load("focus.npz")
    near, far = g["near_far"]
    s = oracle.sample_rays(g["starts"], g["directions"], near, far, 32, u=g["u_uniform"],
                           cdf=g["cdfs"], u_focus=g["u_focus"], focus_stratified=True)
    np.testing.assert_allclose(s.t_values, g["t_values"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(s.positions, g["positions"], rtol=0, atol=2e-6)
    s = oracle.sample_rays(g["starts"], g["directions"], near, far, 32, u=None,
                           cdf=g["cdfs"], focus_stratified=False)
    np.testing.assert_allclose(s.t_values, g["t_det"], rtol=0, atol=1e-6)


def test_determine_cdf_via_reference_cdfs():
    """CDF rows from the reference are monotone, start at 0, end at 1; the oracle
    reproduces them from the oracle's own coarse sigma pass."""
    g = load("focus.npz")
    n = load("nerf_render.npz")
    p = weights(n)
    near, far = g["near_far"]
    t = oracle.linspace(near, far, 16)
    pos = g["starts"][:, None, :] + t[..., None] * g["directions"][:, None, :]
    dirs = np.repeat(g["directions"][:, None, :], 16, 1)
    raw = oracle.nerf_forward(p, pos.reshape(-1, 3).astype(np.float32), dirs.reshape(-1, 3))
    sigma = oracle.softplus(raw[:, 3]).reshape(-1, 16)
    cdf = oracle.determine_cdf(t, sigma)
    np.testing.assert_allclose(cdf, g["cdfs"], rtol=0, atol=5e-5)


def test_torch_oracle_matches_reference_outputs():
    """oracle/ffn_oracle_torch.py (what bench.py times as the reference's CPU path) runs the reference's own ATen op
    sequence: raw outputs within fp32 GEMM blocking differences (the golden run used one 12,288-row batch), pixels
    within 2e-5, and the compositing of the reference's raw outputs bit-identical."""
    import torch
    from oracle import ffn_oracle_torch as ot
    g = load("nerf_render.npz")
    p = {k: torch.from_numpy(v) for k, v in weights(g).items()}
    assert torch.equal(ot.encoding_matrix(9.0, 10), p["pos_encoding"])
    assert torch.equal(ot.encoding_matrix(3.0, 4), p["view_encoding"])
    pos, view = torch.from_numpy(g["positions"]), torch.from_numpy(g["view_directions"])
    with torch.no_grad():
        raw = ot.nerf_forward(p, pos.reshape(-1, 3), view.reshape(-1, 3), pos_enc=p["pos_encoding"],
                              view_enc=p["view_encoding"])
    scale = np.abs(g["raw"]).max()
    assert np.abs(raw.numpy() - g["raw"]).max() <= 2e-5 * max(1.0, scale)
    out = ot.render(raw.reshape(192, 64, 4), torch.from_numpy(g["t_values"]), True)
    np.testing.assert_allclose(out.color.numpy(), g["color"], rtol=0, atol=2e-5)
    np.testing.assert_allclose(out.alpha.numpy(), g["alpha"], rtol=0, atol=2e-5)
    out = ot.render(torch.from_numpy(g["raw"]).reshape(192, 64, 4), torch.from_numpy(g["t_values"]), True)
    assert np.array_equal(out.color.numpy(), g["color"]) and np.array_equal(out.alpha.numpy(), g["alpha"])
    assert np.array_equal(out.depth.numpy(), g["depth"])


def test_torch_oracle_sampling_and_agreement_with_the_numpy_oracle():
    import torch
    from oracle import ffn_oracle_torch as ot
    g = load("sampler.npz")
    rng = np.random.default_rng(3)
    R, S = 64, 64
    o = np.tile(np.array([[0.2, -0.1, -4.0]], np.float32), (R, 1))
    d = rng.normal(size=(R, 3)).astype(np.float32) * 0.15 + np.array([0, 0, 1], np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    near, far = np.full(R, 3.0, np.float32) + rng.random(R, dtype=np.float32), np.full(R, 5.5, np.float32)
    u = rng.random((R, S), dtype=np.float32)
    ref = oracle.sample_rays(o, d, near, far, S, u=u)
    pos, dirs, t = ot.sample_rays(*[torch.from_numpy(a) for a in (o, d, near, far)], S, torch.from_numpy(u))
    assert np.array_equal(t.numpy(), ref.t_values) and np.array_equal(pos.numpy(), ref.positions)
    params = oracle.init_nerf_params(seed=1, gain=2.0)
    want = oracle.render_rays(lambda p, v: oracle.nerf_forward(params, p, v), ref)
    tp = {k: torch.from_numpy(v) for k, v in params.items()}
    got = ot.render_rays(tp, *[torch.from_numpy(a) for a in (o, d, near, far)], S, torch.from_numpy(u), True, batch=48)
    np.testing.assert_allclose(got.color.numpy(), want.color, rtol=0, atol=3e-5)
    np.testing.assert_allclose(got.alpha.numpy(), want.alpha, rtol=0, atol=3e-5)
    assert g is not None


def test_bottleneck_hidden_view_fold_identity():
    """The identity the inference kernel's folded program rests on (nerf_model.py:119-122 has no activation
    between ``bottleneck`` and ``hidden_view``):  hidden_view([bottleneck(h) | enc_v])
    = (W_hv[:, :256] W_b) h + W_hv[:, 256:] enc_v + (W_hv[:, :256] b_b + b_hv).  Checked in fp64 on the reference's own
    golden network: the raw outputs of the folded network equal the reference's recorded fp64 outputs."""
    g = np.load(os.path.join(GOLDEN, "nerf_render.npz"))
    p = {k[2:]: g[k].astype(np.float64) for k in g.files if k.startswith("w.")}
    pos = g["positions"].reshape(-1, 3).astype(np.float64)
    view = g["view_directions"].reshape(-1, 3).astype(np.float64)
    ref = oracle.nerf_forward(p, pos, view)
    assert np.abs(ref - g["raw64"]).max() <= 1e-9 * max(1.0, np.abs(g["raw64"]).max())
    w_hv, w_b = p["hidden_view.weight"], p["bottleneck.weight"]
    folded = dict(p)
    # a "bottleneck" that is the identity and a hidden_view that carries the product: same graph, folded weights
    folded["bottleneck.weight"] = np.eye(256)
    folded["bottleneck.bias"] = np.zeros(256)
    folded["hidden_view.weight"] = np.concatenate([w_hv[:, :256] @ w_b, w_hv[:, 256:]], axis=1)
    folded["hidden_view.bias"] = w_hv[:, :256] @ p["bottleneck.bias"] + p["hidden_view.bias"]
    out = oracle.nerf_forward(folded, pos, view)
    assert np.abs(out - ref).max() <= 1e-10 * max(1.0, np.abs(ref).max())
