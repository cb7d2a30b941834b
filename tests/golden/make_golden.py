"""Generate golden vectors by importing and running the REAL reference.

Run in the build container only (``/root/reference`` does not exist on the GPU
box):  ``python tests/golden/make_golden.py``.  Outputs small ``.npz`` fixtures
next to this file; they are committed and are what pins the oracle
(``tests/test_oracle_golden.py``) and the CUDA path (``tests/test_gpu_parity.py``).

The reference imports four packages that are absent here and do no arithmetic on
the hot path (scenepic, matplotlib.pyplot, progress.bar, trimesh -- SURVEY.md
section 8c); they are replaced by empty module shims before the import.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("FFN_REFERENCE", "/root/reference")


def _shim_modules():
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    class _Any:
        def __init__(self, *a, **k):
            pass

        def __getattr__(self, name):
            return _Any()

        def __call__(self, *a, **k):
            return _Any()

    class Bar:
        def __init__(self, *a, **k):
            self.suffix = ""

        def next(self, *a, **k):
            pass

        def finish(self):
            pass

        def writeln(self, line):
            pass

        @property
        def elapsed(self):
            return 0

    mod("scenepic", Camera=_Any, Transforms=_Any(), Scene=_Any, Mesh=_Any, Colors=_Any())
    mpl = mod("matplotlib")
    mpl.pyplot = mod("matplotlib.pyplot", get_cmap=lambda *a, **k: None, Axes=_Any, Figure=_Any)
    prog = mod("progress")
    prog.bar = mod("progress.bar", Bar=Bar, ChargingBar=Bar)
    mod("trimesh")


def import_reference():
    _shim_modules()
    sys.path.insert(0, REF)
    import fourier_feature_nets as ffn  # noqa: E402
    assert os.path.realpath(ffn.__file__).startswith(os.path.realpath(REF))
    return ffn


def look_at_camera(ffn, name, position, resolution, fov_y_deg=40.0):
    """camera-to-world extrinsics looking at the origin (OpenCV convention:
    +z forward, +y down), intrinsics as utils.py:279-286."""
    position = np.asarray(position, np.float32)
    forward = -position / np.linalg.norm(position)
    up = np.array([0, 1, 0], np.float32)
    if abs(forward @ up) > 0.99:
        up = np.array([0, 0, 1], np.float32)
    right = np.cross(forward, up)
    right /= np.linalg.norm(right)
    down = np.cross(forward, right)
    ext = np.eye(4, dtype=np.float32)
    ext[:3, 0], ext[:3, 1], ext[:3, 2], ext[:3, 3] = right, down, forward, position
    focal = 0.5 * resolution / np.tan(np.radians(fov_y_deg) / 2)
    intr = np.array([[focal, 0, resolution / 2], [0, focal, resolution / 2], [0, 0, 1]], np.float32)
    return ffn.CameraInfo.create(name, ffn.Resolution(resolution, resolution), intr, ext)


def sharpen(model, gain, seed):
    """Scale weights so that sigma spans a useful range (random default-init
    nets give near-constant sigma ~ softplus(0)); keeps biases."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if p.requires_grad and name.endswith("weight"):
                p.mul_(gain)
            elif p.requires_grad:
                p.add_(torch.randn(p.shape, generator=g) * 0.1)


def state_np(model):
    return {k: v.detach().cpu().numpy() for k, v in model.state_dict().items()}


def main():
    ffn = import_reference()
    torch.manual_seed(20080524)  # train_nerf.py:48
    torch.set_num_threads(8)

    # ---- cameras / sampler -------------------------------------------------
    res = 24
    cams = [look_at_camera(ffn, "c%d" % i, p, res) for i, p in enumerate(
        [(0, 0, -4), (4, 0.5, 0.3), (-2.5, 1.5, 2.8)])]
    bounds = np.diag([2, 2, 2, 1]).astype(np.float32)  # orbit_video.py:64
    S = 64
    sampler = ffn.RaySampler(bounds, cams, S, stratified=True)
    valid = sampler.to_valid(list(range(len(sampler))))
    rng = np.random.default_rng(7)
    idx = sorted(rng.choice(valid, size=192, replace=False).tolist())

    # capture the jitter torch.rand draws inside sample()
    torch.manual_seed(1234)
    u = torch.rand((len(idx), S), dtype=torch.float32)
    torch.manual_seed(1234)
    samples = sampler.sample(idx, None)
    torch.manual_seed(99)
    u_anneal = torch.rand((len(idx), S), dtype=torch.float32)
    sampler.num_anneal_steps, sampler.anneal_start = 2000, 0.2
    torch.manual_seed(99)
    samples_anneal = sampler.sample(idx, 500)
    sampler.num_anneal_steps = 0
    sampler.stratified = False
    samples_uniform = sampler.sample(idx, None)
    sampler.stratified = True

    cam_pts = sampler.points
    o0, d0 = cams[1].raycast(cam_pts)

    np.savez_compressed(
        os.path.join(HERE, "sampler.npz"),
        bounds=bounds, points=cam_pts.astype(np.int64),
        intrinsics=np.stack([c.intrinsics for c in cams]),
        extrinsics=np.stack([c.extrinsics for c in cams]),
        cam1_starts=o0, cam1_dirs=d0,
        starts=sampler.starts.numpy(), directions=sampler.directions.numpy(),
        near_far=sampler.near_far.numpy(),
        invalid=np.array(sorted(sampler.invalid_rays), np.int64),
        idx=np.array(idx, np.int64), u=u.numpy(),
        positions=samples.positions.numpy(), t_values=samples.t_values.numpy(),
        view_directions=samples.view_directions.numpy(),
        u_anneal=u_anneal.numpy(), t_anneal=samples_anneal.t_values.numpy(),
        pos_anneal=samples_anneal.positions.numpy(),
        t_uniform=samples_uniform.t_values.numpy(),
        linspace_64=torch.linspace(0, 1, 64).numpy(),
        linspace_63=torch.linspace(0, 1, 63).numpy(),
        linspace_7=torch.linspace(0, 1, 7).numpy(),
    )

    # ---- NeRF render -------------------------------------------------------
    torch.manual_seed(20080524)
    nerf = ffn.NeRF(8, 256, 9, 10, 3, 4, [4], True)  # train_nerf.py:80-83
    sharpen(nerf, 2.2, 5)
    with torch.no_grad():  # trained-like dynamic range: sigma from ~0 to >100, saturated colours
        nerf.opacity_out.weight.mul_(40.0)
        nerf.opacity_out.bias.sub_(8.0)
        nerf.color_out.weight.mul_(6.0)
    nerf.eval()
    rc = ffn.Raycaster(nerf)
    with torch.no_grad():
        raw = nerf(samples.positions.reshape(-1, 3), samples.view_directions.reshape(-1, 3))
        out = rc.render(samples, True)
        nerf64 = nerf.double()
        s64 = ffn.RaySamples(samples.positions.double(), samples.view_directions.double(),
                             samples.t_values.double(), samples.rays)
        raw64 = nerf64(s64.positions.reshape(-1, 3), s64.view_directions.reshape(-1, 3))
        out64 = ffn.Raycaster(nerf64).render(s64, True)
        nerf.float()
    sd = state_np(nerf)
    np.savez_compressed(
        os.path.join(HERE, "nerf_render.npz"),
        **{"w." + k: v.astype(np.float32) for k, v in sd.items()},
        idx=np.array(idx, np.int64),
        positions=samples.positions.numpy(), view_directions=samples.view_directions.numpy(),
        t_values=samples.t_values.numpy(),
        raw=raw.numpy(), color=out.color.numpy(), alpha=out.alpha.numpy(), depth=out.depth.numpy(),
        raw64=raw64.numpy(), color64=out64.color.numpy(), alpha64=out64.alpha.numpy(),
        depth64=out64.depth.numpy(),
    )
    print("nerf: sigma range", torch.nn.functional.softplus(raw[:, 3]).min().item(),
          torch.nn.functional.softplus(raw[:, 3]).max().item(),
          "alpha range", out.alpha.min().item(), out.alpha.max().item())

    # ---- FourierFeatureMLP presets (train_tiny_nerf.py:75-88) --------------
    sub = ffn.RaySamples(samples.positions[:64], samples.view_directions[:64],
                         samples.t_values[:64], samples.rays[:64])
    presets = {
        "mlp": lambda: ffn.MLP(3, 4),
        "basic": lambda: ffn.BasicFourierMLP(3, 4),
        "positional": lambda: ffn.PositionalFourierMLP(3, 4, 5.5),
        "gaussian": lambda: ffn.GaussianFourierMLP(3, 4, 3.14),
    }
    for name, ctor in presets.items():
        torch.manual_seed(20080524)
        model = ctor()
        sharpen(model, 1.6, 11)
        model.eval()
        with torch.no_grad():
            raw = model(sub.positions.reshape(-1, 3))
            out = ffn.Raycaster(model).render(sub, True)
        sd = state_np(model)
        np.savez_compressed(
            os.path.join(HERE, "ffmlp_%s.npz" % name),
            **{"w." + k: v.astype(np.float32) for k, v in sd.items()},
            positions=sub.positions.numpy(), t_values=sub.t_values.numpy(),
            raw=raw.numpy(), color=out.color.numpy(), alpha=out.alpha.numpy(),
            depth=out.depth.numpy())
        print(name, "alpha range", out.alpha.min().item(), out.alpha.max().item())

    # ---- focus sampling (ray_sampler.py:59-67,234-269,301-357) -------------
    torch.manual_seed(20080524)
    cams2 = cams[:1]
    fs = ffn.RaySampler(bounds, cams2, 32, stratified=True, opacity_model=nerf, batch_size=256)
    valid2 = fs.to_valid(list(range(len(fs))))
    idx2 = valid2[:: max(1, len(valid2) // 48)][:48]
    torch.manual_seed(4321)
    u_a = torch.rand((len(idx2), 16), dtype=torch.float32)
    u_b = torch.rand((len(idx2), 16), dtype=torch.float32)
    torch.manual_seed(4321)
    fsamp = fs.sample(idx2, None)
    fs.stratified = False
    fsamp_det = fs.sample(idx2, None)
    np.savez_compressed(
        os.path.join(HERE, "focus.npz"),
        idx=np.array(idx2, np.int64), cdfs=fs.cdfs.numpy()[idx2],
        starts=fs.starts.numpy()[idx2], directions=fs.directions.numpy()[idx2],
        near_far=fs.near_far.numpy()[:, idx2],
        u_uniform=u_a.numpy(), u_focus=u_b.numpy(),
        t_values=fsamp.t_values.numpy(), positions=fsamp.positions.numpy(),
        t_det=fsamp_det.t_values.numpy())

    # ---- ImageDataset: index tables, GT lookup, loss (image_dataset.py) ----------------
    rng = np.random.default_rng(11)
    res2 = 20
    dcams = [look_at_camera(ffn, "d%d" % i, p, res2) for i, p in enumerate(
        [(0, 0, -4), (4, 0.5, 0.3), (-2.5, 1.5, 2.8), (0.3, 3.9, 0.2), (2, -2, 2.5)])]
    imgs = rng.integers(0, 256, size=(5, res2, res2, 4), dtype=np.uint8)
    yy, xx = np.mgrid[:res2, :res2]
    for i in range(5):      # alpha: a disc, zero outside
        imgs[i, ..., 3] = np.where((xx - 10 - i) ** 2 + (yy - 9) ** 2 < 30, imgs[i, ..., 3] | 1, 0)
    ds = ffn.ImageDataset("train", imgs, bounds, dcams, 16, True, True, None, 4096, "RGB", 6, 0.2, 0)
    out = {"images": imgs, "bounds": bounds,
           "intrinsics": np.stack([c.intrinsics for c in dcams]),
           "extrinsics": np.stack([c.extrinsics for c in dcams]),
           "crop_index": ds.crop_index.numpy(), "sparse_index": ds.sparse_index.numpy(),
           "dilate_index": ds.dilate_index.numpy(), "dilate_ranges": np.array(ds.dilate_ranges),
           "colors": ds.colors.numpy(), "alphas": ds.alphas.numpy()}
    batch = rng.integers(0, 5 * res2 * res2, size=200).tolist()
    for mode in ("Full", "Center", "Sparse", "Dilate"):
        ds.mode = getattr(ffn.RayDataset.Mode, mode)
        b = [i % len(ds) for i in batch]
        torch.manual_seed(5)
        rays = ds.get_rays(b, 3)
        fake = ffn.utils.RenderResult(torch.rand(len(rays.rays), 3, generator=torch.Generator().manual_seed(1)),
                                      torch.rand(len(rays.rays), generator=torch.Generator().manual_seed(2)), None)
        out["len_" + mode] = np.array(len(ds))
        out["batch_" + mode] = np.array(b)
        out["rays_" + mode] = rays.rays.numpy()
        out["loss_" + mode] = ds.loss(3, rays, fake).numpy()
        out["fake_color_" + mode] = fake.color.numpy()
        out["fake_alpha_" + mode] = fake.alpha.numpy()
        out["index_cam2_" + mode] = np.array(ds.index_for_camera(2))
        out["rays_cam2_" + mode] = ds.rays_for_camera(2).rays.numpy()
    ds.mode = ffn.RayDataset.Mode.Full
    sub = ds.sample_cameras(3, 16, False)
    out["sample_cameras_names"] = np.array([c.name for c in sub.cameras])
    np.savez_compressed(os.path.join(HERE, "dataset.npz"), **out)

    # ---- known-answer: docs/ray_data.tsv -----------------------------------
    tsv = np.loadtxt(os.path.join(REF, "docs", "ray_data.tsv"), skiprows=1, dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "ray_data_kat.npz"),
                        t=tsv[:, 0].astype(np.float32), opacity=tsv[:, 1].astype(np.float32),
                        T=tsv[:, 2].astype(np.float32))

    # ---- blend weights on adversarial inputs -------------------------------
    g = torch.Generator().manual_seed(3)
    t = torch.sort(torch.rand((32, 48), generator=g) * 3 + 2, -1)[0]
    sig = torch.exp(torch.randn((32, 48), generator=g) * 3)
    sig[0] = 0
    sig[1] = 1e4
    w = ffn.calculate_blend_weights(t, sig)
    np.savez_compressed(os.path.join(HERE, "blend.npz"), t=t.numpy(), sigma=sig.numpy(), w=w.numpy())
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
