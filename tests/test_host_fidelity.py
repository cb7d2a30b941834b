"""Host-side fidelity around the hot path: orbit camera
construction and the EvaluationVisualizer PNG grid against fixtures produced by the REAL reference
(tests/golden/make_golden_r2.py), plus -- where the reference itself is importable (build container: /root/reference;
GPU box: oracle/_ref) -- live comparisons."""
import os
import sys

import cv2
import numpy as np
import pytest
import torch

from conftest import GOLDEN, ROOT

import fourier_feature_nets_b200 as ffn
from oracle import reference as refmod

sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from make_golden_r2 import fake_prediction  # noqa: E402


def _orbit_args(a):
    up, fwd = a[:3].astype(np.float32), a[3:6].astype(np.float32)
    return up, fwd, int(a[6]), int(a[7]), float(a[8])


@pytest.mark.parametrize("tag", ["a", "b"])
def test_orbit_matches_the_reference_fixture(tag):
    """utils.orbit (utils.py:244-300): extrinsics / intrinsics of every frame as the reference computes them."""
    g = np.load(os.path.join(GOLDEN, "orbit.npz"))
    up, fwd, n, res, dist = _orbit_args(g["args_" + tag])
    cams = ffn.orbit(up, fwd, n, 40, ffn.Resolution(res, res), dist)
    assert len(cams) == n
    ext = np.stack([c.extrinsics for c in cams])
    np.testing.assert_allclose(ext, g["ext_" + tag], atol=2e-6)
    np.testing.assert_array_equal(np.stack([c.intrinsics for c in cams]), g["int_" + tag])


@pytest.mark.parametrize("tag", ["a", "b"])
def test_orbit_fixture_against_an_independent_closed_form(tag):
    """The fixture itself (reference code + the scenepic stand-in) checked without either: frame i sits at
    R_up(azi_i) R_right(alt_i) (-forward * distance), looks at the origin along +z of an OpenCV camera (columns of the
    camera-to-world rotation: right, down, forward), its x axis is perpendicular to the orbit's up direction, and the
    rotation is proper.  azi = linspace(0, 4 pi, n, endpoint=False); alt ramps pi/12 -> pi/4 -> pi/12."""
    g = np.load(os.path.join(GOLDEN, "orbit.npz"))
    up, fwd, n, res, dist = _orbit_args(g["args_" + tag])
    up, fwd = up.astype(np.float64), fwd.astype(np.float64)
    right = np.cross(up, fwd)

    def rot(axis, ang):          # active rotation about a unit axis (Rodrigues, written out with cross products)
        def f(v):
            return v * np.cos(ang) + np.cross(axis, v) * np.sin(ang) + axis * (axis @ v) * (1 - np.cos(ang))
        return f

    azi = np.arange(n) * 4 * np.pi / n
    half = n // 2
    alt = np.concatenate([np.pi / 12 + (np.pi / 4 - np.pi / 12) * np.arange(half) / half,
                          np.pi / 4 + (np.pi / 12 - np.pi / 4) * np.arange(n - half) / (n - half)])
    for i in range(n):
        E = g["ext_" + tag][i].astype(np.float64)
        pos = rot(up, azi[i])(rot(right, alt[i])(-fwd * dist))
        np.testing.assert_allclose(E[:3, 3], pos, atol=5e-6)
        np.testing.assert_allclose(E[:3, 2], -pos / np.linalg.norm(pos), atol=5e-6)     # looks at the origin
        assert abs(E[:3, 0] @ up) < 5e-6                                               # horizon stays level
        assert (E[:3, 1] @ up) < 0                                                     # +y is down
        np.testing.assert_allclose(E[:3, :3].T @ E[:3, :3], np.eye(3), atol=5e-6)
        assert np.linalg.det(E[:3, :3]) > 0.999
        # altitude: elevation of the camera above the plane perpendicular to `up`, sign as the reference's rotation
        assert abs(abs(np.arcsin(pos @ up / dist)) - alt[i]) < 1e-6
    focal = .5 * res / np.tan(.5 * 40 * np.pi / 180)
    np.testing.assert_allclose(g["int_" + tag][0], [[focal, 0, res / 2], [0, focal, res / 2], [0, 0, 1]], rtol=1e-6)


def _dataset(pkg):
    g = np.load(os.path.join(GOLDEN, "dataset.npz"))
    cams = [pkg.CameraInfo.create("d%d" % i, pkg.Resolution(20, 20), g["intrinsics"][i], g["extrinsics"][i])
            for i in range(len(g["intrinsics"]))]
    return pkg.ImageDataset("val", g["images"], g["bounds"], cams, 16, True, False, None, 4096, "RGB", 6, 0.2, 0)


def _run_visualizer(pkg, out_dir, as_tensor=False):
    vis = pkg.EvaluationVisualizer(out_dir, _dataset(pkg), 5, max_depth=10)
    for step in (0, 3, 5):          # step 3 is skipped (interval 5)
        def render(samples, include_depth):
            assert include_depth
            c, a, d = fake_prediction(len(samples.rays), 100 + step)
            return pkg.utils.RenderResult(c, a, d)
        vis.visualize(step, render, None)
    names = sorted(os.listdir(os.path.join(out_dir, "val")))
    return names, [cv2.imread(os.path.join(out_dir, "val", n)) for n in names]


def test_evaluation_visualizer_png_equals_the_reference_fixture(tmp_path):
    """visualizers.py:55-102: prediction | depth / ground truth x alpha | normalised error, uint8 truncation, file
    names ``s{step:07}_c{camera:03}.png`` -- pixel for pixel what the reference wrote for the same inputs."""
    g = np.load(os.path.join(GOLDEN, "eval_visualizer.npz"))
    names, grids = _run_visualizer(ffn, str(tmp_path))
    assert names == list(g["names"]) == ["s0000000_c000.png", "s0000005_c001.png"]
    for i, grid in enumerate(grids):
        assert grid.shape == (40, 40, 3)
        np.testing.assert_array_equal(grid, g["grid%d" % i])


@pytest.mark.skipif(not refmod.available(), reason="needs the reference (oracle/_ref or /root/reference)")
def test_evaluation_visualizer_and_orbit_live_against_the_reference(tmp_path):
    ref = refmod.import_reference()
    n_ref, g_ref = _run_visualizer(ref, str(tmp_path / "ref"))
    n_our, g_our = _run_visualizer(ffn, str(tmp_path / "ours"))
    assert n_ref == n_our
    for a, b in zip(g_ref, g_our):
        np.testing.assert_array_equal(a, b)
    for up, fwd, n in (((0, 1, 0), (0, 0, -1), 9), ((1, 0, 0), (0, 1, 0), 4)):
        up, fwd = np.array(up, np.float32), np.array(fwd, np.float32)
        a = ref.orbit(up, fwd, n, 33, ref.Resolution(10, 8), 3.5, np.pi / 10, np.pi / 3)
        b = ffn.orbit(up, fwd, n, 33, ffn.Resolution(10, 8), 3.5, np.pi / 10, np.pi / 3)
        for ca, cb in zip(a, b):
            np.testing.assert_allclose(ca.extrinsics, cb.extrinsics, atol=2e-6)
            np.testing.assert_array_equal(ca.intrinsics, cb.intrinsics)


def test_unsupported_shapes_are_reported_not_silent():
    """engine.supported() checks the shape as well as the kind; note_unfused warns once (FFN_STRICT=1: raises)."""
    from fourier_feature_nets_b200 import _lib, engine
    assert engine.supported(ffn.NeRF(8, 256, 9, 10, 3, 4, [4], True))
    assert engine.supported(ffn.PositionalFourierMLP(3, 4, 5.5))
    for bad in (ffn.NeRF(8, 128, 9, 10, 3, 4, [4], True), ffn.NeRF(8, 256, 9, 12, 3, 4, [4], True),
                ffn.NeRF(4, 256, 9, 10, 3, 4, [0], True), ffn.MLP(3, 4, num_channels=128),
                ffn.GaussianFourierMLP(3, 4, 2.0, embedding_size=320), ffn.MLP(2, 3)):
        assert not engine.supported(bad) and engine.unsupported_reason(bad)
    m = ffn.MLP(3, 4)
    m.keep_activations = True
    assert not engine.supported(m)
    assert not engine.supported(ffn.Voxels(8, 2.0))
    bad = ffn.NeRF(8, 128, 9, 10, 3, 4, [4], True)
    with pytest.warns(UserWarning, match="num_channels = 128"):
        engine.note_unfused(bad)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        engine.note_unfused(bad)                 # once per model
        engine.note_unfused(ffn.Voxels(8, 2.0))    # other model families: nothing to report
    os.environ["FFN_STRICT"] = "1"
    try:
        with pytest.raises(_lib.FFNError):
            engine.note_unfused(ffn.NeRF(8, 128, 9, 10, 3, 4, [4], True))
    finally:
        del os.environ["FFN_STRICT"]


def test_ray_offset_follows_subsets():
    """RayBundle.subset keeps the position of its first ray inside the sampled bundle (in-kernel jitter is keyed on
    seed, ray_offset + i): contiguous batches of batched_render draw different jitter."""
    n = 10
    z = torch.zeros(n)
    b = ffn.RayBundle(torch.zeros(n, 3), torch.zeros(n, 3), z, z + 1, torch.arange(n), 4, True, None, seed=9)
    assert b.ray_offset == 0
    s = b.subset(range(4, 8))
    assert s.ray_offset == 4 and s.seed == 9 and s.num_rays == 4
    assert s.subset(range(2, 4)).ray_offset == 6
    assert b.subset(list(range(3, 6))).ray_offset == 3
    assert b.to("cpu").ray_offset == 0 and s.to("cpu").ray_offset == 4
    # materialised jitter of two batches differs; the same batch twice is reproducible
    t0, t1 = b.subset(range(0, 4)).t_values, b.subset(range(4, 8)).t_values
    assert not torch.equal(t0, t1)
    assert torch.equal(b.subset(range(4, 8)).t_values, t1)


def test_render_stream_equals_batchwise_render():
    """Raycaster.render_stream (pipelined host-in / host-out API) yields, in order, exactly what
    ``render(sampler.sample(idx).to(device)).numpy()`` returns per batch (here on the CPU definition; the CUDA
    pipeline is covered by tests/test_gpu_parity.py)."""
    res = 24
    f = .5 * res / np.tan(.5 * 40 * np.pi / 180)
    K = np.array([[f, 0, res / 2], [0, f, res / 2], [0, 0, 1]], np.float32)
    E = ffn.utils.look_at_extrinsics(np.array([0.5, 1.0, -3.8]), np.array([0, 1.0, 0])).astype(np.float32)
    cam = ffn.CameraInfo.create("c", ffn.Resolution(res, res), K, E)
    s = ffn.RaySampler(np.diag([2, 2, 2, 1]).astype(np.float32), [cam], 8)
    torch.manual_seed(0)
    rc = ffn.Raycaster(ffn.NeRF(2, 32, 3, 4, 2, 2, [1], True))
    valid = torch.nonzero(s.valid_mask).flatten()
    batches = [valid[:100], valid[100:250], valid[250:257]]
    got = list(rc.render_stream(s, batches, True))
    assert len(got) == 3
    with torch.no_grad():
        for g, b in zip(got, batches):
            want = rc.render(s.sample(b, None), True).numpy()
            np.testing.assert_array_equal(g.color, want.color)
            np.testing.assert_array_equal(g.alpha, want.alpha)
            np.testing.assert_array_equal(g.depth, want.depth)
    assert [len(g.alpha) for g in got] == [100, 150, 7]
    assert list(rc.render_stream(s, [], True)) == []
