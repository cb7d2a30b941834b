"""CPU suite: C-ABI surface, host-side mirror of the reference API, multi-process plumbing."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

import fourier_feature_nets_b200 as ffn
from fourier_feature_nets_b200 import _lib
from conftest import GOLDEN, ROOT


def load(name):
    return np.load(os.path.join(GOLDEN, name))


def golden_cameras(g):
    return [ffn.CameraInfo.create("c%d" % i, ffn.Resolution(24, 24), g["intrinsics"][i], g["extrinsics"][i])
            for i in range(len(g["intrinsics"]))]


def test_library_builds_and_exports_every_declared_symbol():
    path = _lib.build()
    L = ctypes.CDLL(path)
    header = open(os.path.join(ROOT, "include", "ffn_b200.h")).read()
    declared = set(re.findall(r"\b(ffn_[a-z_0-9]+)\s*\(", header))
    assert declared, "no declarations found in include/ffn_b200.h"
    assert declared == set(_lib.EXPORTED_SYMBOLS)
    for sym in declared:
        assert hasattr(L, sym), sym
    L.ffn_version.restype = ctypes.c_int32
    assert L.ffn_version() == 100


def test_library_has_blackwell_sass():
    """tcgen05 / TMEM / bulk-copy instructions are in the cubin (B200_PROFILING.md table)."""
    out = subprocess.run(["cuobjdump", "-sass", _lib.build()], capture_output=True, text=True).stdout
    # tcgen05.mma (cta_group::2 in the render kernels), tcgen05.ld, 2-SM tensor-map loads of the weight ring, TMA stores
    for mnemonic in ("UTCHMMA.2CTA", "LDTM", "UTMALDG.2D.2CTA", "UTMASTG", "UTCBAR.2CTA.MULTICAST"):
        assert mnemonic in out, mnemonic
    assert "HMMA." not in out.replace("UTCHMMA", "")   # no legacy mma.sync path


def test_no_cpu_fallback_for_cuda_entry_points():
    net_less = torch.zeros(4, 3)
    with pytest.raises(_lib.FFNError):
        _lib.blend_weights(torch.zeros(2, 4), torch.zeros(2, 4))     # CPU tensors are refused
    with pytest.raises(_lib.FFNError):
        _lib._f32c(net_less, "x")


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "fourier_feature_nets_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f


def test_sampler_tables_match_reference():
    g = load("sampler.npz")
    s = ffn.RaySampler(g["bounds"], golden_cameras(g), 64, stratified=True)
    assert np.array_equal(s.starts.numpy(), g["starts"])
    assert np.array_equal(s.directions.numpy(), g["directions"])
    v = s.valid_mask.numpy()
    assert np.array_equal(np.nonzero(~v)[0], g["invalid"])
    assert s.invalid_rays == set(g["invalid"].tolist())
    assert np.array_equal(s.near_far.numpy()[:, v], g["near_far"][:, v])
    assert len(s) == 3 * 24 * 24 and s.num_cameras == 3 and s.rays_per_camera == 576


def test_sample_is_a_raysamples_and_matches_reference_bitwise():
    g = load("sampler.npz")
    s = ffn.RaySampler(g["bounds"], golden_cameras(g), 64, stratified=True)
    torch.manual_seed(1234)
    b = s.sample(g["idx"].tolist(), None)
    assert isinstance(b, ffn.RaySamples) and isinstance(b, ffn.RayBundle)
    assert np.array_equal(b.t_values.numpy(), g["t_values"])
    assert np.array_equal(b.positions.numpy(), g["positions"])
    assert np.array_equal(b.view_directions.numpy(), g["view_directions"])
    assert np.array_equal(b.rays.numpy(), g["idx"])
    pos, views, t, rays = b                       # tuple protocol
    assert pos.shape == (192, 64, 3) and rays.shape == (192,)
    sub = b.subset([3, 4, 5])
    assert np.array_equal(sub.positions.numpy(), g["positions"][3:6])
    sub = b.subset([7, 2])
    assert np.array_equal(sub.t_values.numpy(), g["t_values"][[7, 2]])
    n = b.numpy()
    assert isinstance(n.positions, np.ndarray)
    # annealing (ray_sampler.py:373-378)
    s.num_anneal_steps, s.anneal_start = 2000, 0.2
    torch.manual_seed(99)
    a = s.sample(g["idx"].tolist(), 500)
    assert np.array_equal(a.t_values.numpy(), g["t_anneal"])
    assert np.array_equal(a.positions.numpy(), g["pos_anneal"])
    s.num_anneal_steps, s.stratified = 0, False
    assert np.array_equal(s.sample(g["idx"].tolist(), None).t_values.numpy(), g["t_uniform"])


def test_to_valid_and_to_image():
    g = load("sampler.npz")
    s = ffn.RaySampler(g["bounds"], golden_cameras(g), 8)
    inv = set(g["invalid"].tolist())
    idx = list(range(500, 700))
    assert s.to_valid(idx) == [i for i in idx if i not in inv]
    cam = 1
    valid = s._valid_for_camera(cam)
    colors = np.random.default_rng(0).random((len(valid), 3)).astype(np.float32)
    img = s.to_image(cam, colors, "RGB")
    assert img.shape == (24, 24, 3) and img.dtype == np.uint8
    flat = img.reshape(-1, 3)
    local = (valid - cam * 576).numpy()
    assert np.array_equal(flat[local], (colors * 255).astype(np.uint8))
    mask = np.ones(576, bool)
    mask[local] = False
    assert (flat[mask] == 0).all()


def test_cpu_render_definition_matches_reference():
    """The differentiable torch definition (used under autograd / on CPU) equals the
    reference's output on the golden vectors."""
    g = load("nerf_render.npz")
    m = ffn.NeRF(8, 256, 9, 10, 3, 4, [4], True)
    m.load_state_dict({k[2:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("w.")})
    samples = ffn.RaySamples(*[torch.from_numpy(g[k]) for k in ("positions", "view_directions", "t_values")],
                             torch.from_numpy(g["idx"]))
    with torch.no_grad():
        out = ffn.Raycaster(m).render(samples, True)
    np.testing.assert_allclose(out.color.numpy(), g["color"], atol=1e-6)
    np.testing.assert_allclose(out.alpha.numpy(), g["alpha"], atol=1e-6)
    assert np.array_equal(out.depth.numpy(), g["depth"])


@pytest.mark.parametrize("name", ["mlp", "basic", "positional", "gaussian"])
def test_ffmlp_state_dicts_interchange(name):
    g = load("ffmlp_%s.npz" % name)
    ctor = {"mlp": lambda: ffn.MLP(3, 4), "basic": lambda: ffn.BasicFourierMLP(3, 4),
            "positional": lambda: ffn.PositionalFourierMLP(3, 4, 5.5),
            "gaussian": lambda: ffn.GaussianFourierMLP(3, 4, 3.14)}[name]
    m = ctor()
    m.load_state_dict({k[2:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("w.")})
    with torch.no_grad():
        raw = m(torch.from_numpy(g["positions"]).reshape(-1, 3))
    np.testing.assert_allclose(raw.numpy(), g["raw"], atol=2e-6 * max(1, np.abs(g["raw"]).max()))


def test_model_save_load_roundtrip(tmp_path):
    m = ffn.NeRF(8, 256, 9, 10, 3, 4, [4], True)
    path = str(tmp_path / "nerf.pt")
    m.save(path)
    raw = torch.load(path, weights_only=False)
    assert raw["type"] == "nerf" and raw["params"]["skips"] == [4]
    m2 = ffn.load_model(path)
    for (k1, v1), (k2, v2) in zip(m.state_dict().items(), m2.state_dict().items()):
        assert k1 == k2 and torch.equal(v1, v2)
    f = ffn.PositionalFourierMLP(3, 4, 5.5)
    path = str(tmp_path / "f.pt")
    f.save(path)
    f2 = ffn.load_model(path)
    assert torch.equal(f.b_values, f2.b_values)


def test_blend_weights_definition_and_lr_decay():
    g = load("blend.npz")
    w = ffn.calculate_blend_weights(torch.from_numpy(g["t"]), torch.from_numpy(g["sigma"]))
    np.testing.assert_allclose(w.numpy(), g["w"], atol=1e-7)
    opt = torch.optim.Adam([torch.nn.Parameter(torch.zeros(1))], 5e-4)
    ffn.exponential_lr_decay(opt, 5e-4, 250, 0.1, 500)
    assert abs(opt.param_groups[0]["lr"] - 5e-4 * 0.1 ** 0.5) < 1e-12


def test_shard_range_partitions():
    from fourier_feature_nets_b200.parallel import shard_range
    for n in (0, 1, 7, 1024, 1025):
        for ws in (1, 2, 3, 8):
            spans = [shard_range(n, r, ws) for r in range(ws)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


_WORKER = r"""
import os, sys
sys.path.insert(0, sys.argv[1])
import torch, torch.distributed as dist
import fourier_feature_nets_b200 as ffn
from fourier_feature_nets_b200 import parallel
dist.init_process_group("gloo")
rank, ws = dist.get_rank(), dist.get_world_size()
torch.manual_seed(100 + rank)                      # different init per rank on purpose
m = ffn.MLP(3, 4)
parallel.broadcast_parameters(m)
ref = ffn.MLP(3, 4); torch.manual_seed(100); ref = ffn.MLP(3, 4)
torch.manual_seed(7)
x = torch.randn(64, 3); y = torch.randn(64, 4)
lo, hi = parallel.shard_range(64)
loss = ((m(x[lo:hi]) - y[lo:hi]) ** 2).sum() / 64      # sum over shards == full-batch mean
loss.backward()
parallel.allreduce_gradients(m, average=False)
# single-process reference on the full batch with rank 0's weights
torch.manual_seed(100); full = ffn.MLP(3, 4)
((full(x) - y) ** 2).sum().div(64).backward()
err = max((a.grad - b.grad).abs().max().item() for a, b in zip(m.parameters(), full.parameters()))
assert err < 1e-5, err
# flat-buffer path (what the training kernels produce): .grad are views of one buffer registered on the model
from fourier_feature_nets_b200.autograd import _flat_grads
params = list(m.parameters())
flat, views = _flat_grads(params, "cpu")
for p, v in zip(params, views):
    v.copy_(torch.full_like(v, float(rank + 1)))
    p.grad = v
m.__dict__["_ffn_flat_grad"] = flat
parallel.allreduce_gradients(m, average=True)
assert all(p.grad.untyped_storage().data_ptr() == flat.untyped_storage().data_ptr() for p in params)
assert all(torch.equal(p.grad, torch.full_like(p.grad, (1 + ws) / 2)) for p in params)
print("rank", rank, "ok", err)
dist.destroy_process_group()
"""


def test_gradient_allreduce_two_ranks_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="2")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29631", str(script), ROOT]
    res = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=240)
    assert res.returncode == 0, res.stdout + res.stderr
    assert res.stdout.count("ok") == 2


def _golden_dataset():
    from fourier_feature_nets_b200.image_dataset import ImageDataset
    g = load("dataset.npz")
    cams = [ffn.CameraInfo.create("d%d" % i, ffn.Resolution(20, 20), g["intrinsics"][i], g["extrinsics"][i])
            for i in range(5)]
    ds = ImageDataset("train", g["images"], g["bounds"], cams, 16, True, True, None, 4096, "RGB", 6, 0.2, 0)
    return g, ds


def test_image_dataset_tables_match_reference():
    g, ds = _golden_dataset()
    assert np.array_equal(ds.crop_index.numpy(), g["crop_index"])
    assert np.array_equal(ds.sparse_index.numpy(), g["sparse_index"])
    assert np.array_equal(ds.dilate_index.numpy(), g["dilate_index"])
    assert np.array_equal(np.array(ds.dilate_ranges), g["dilate_ranges"])
    assert np.array_equal(ds.colors.numpy(), g["colors"]) and np.array_equal(ds.alphas.numpy(), g["alphas"])


@pytest.mark.parametrize("mode", ["Full", "Center", "Sparse", "Dilate"])
def test_image_dataset_get_rays_and_loss_match_reference(mode):
    g, ds = _golden_dataset()
    ds.mode = getattr(ffn.Mode, mode)
    assert len(ds) == int(g["len_" + mode])
    rays = ds.get_rays(g["batch_" + mode].tolist(), 3)
    assert np.array_equal(rays.rays.numpy(), g["rays_" + mode])
    fake = ffn.RenderResult(torch.from_numpy(g["fake_color_" + mode]), torch.from_numpy(g["fake_alpha_" + mode]), None)
    assert abs(float(ds.loss(3, rays, fake)) - float(g["loss_" + mode])) < 1e-7
    assert np.array_equal(np.array(ds.index_for_camera(2)), g["index_cam2_" + mode])
    assert np.array_equal(ds.rays_for_camera(2).rays.numpy(), g["rays_cam2_" + mode])


def test_image_dataset_sample_cameras_and_npz_load(tmp_path):
    g, ds = _golden_dataset()
    sub = ds.sample_cameras(3, 16, False)
    assert sorted(c.name for c in sub.cameras) == sorted(g["sample_cameras_names"].tolist())
    path = str(tmp_path / "toy.npz")
    np.savez(path, images=g["images"], bounds=g["bounds"], intrinsics=g["intrinsics"],
             extrinsics=g["extrinsics"], split_counts=np.array([3, 1, 1]))
    from fourier_feature_nets_b200.image_dataset import ImageDataset
    tr = ImageDataset.load(path, "train", 8, True, True)
    va = ImageDataset.load(path, "val", 8, True, False)
    assert tr.num_cameras == 3 and va.num_cameras == 1 and tr.cameras[2].name == "train002"
    img = va.to_image(0, np.ones((len(va.index_for_camera(0)), 3), np.float32))
    assert img.shape == (20, 20, 3) and img.max() == 255


# ---- checkpoint fidelity: .pt files written by the real reference ----------------
def test_checkpoints_written_by_the_reference_load_and_reproduce():
    g = load("checkpoints.npz")
    pos, view = torch.from_numpy(g["pos"]), torch.from_numpy(g["view"])
    nerf = ffn.load_model(os.path.join(GOLDEN, "ref_nerf_small.pt"))
    four = ffn.load_model(os.path.join(GOLDEN, "ref_fourier_small.pt"))
    assert isinstance(nerf, ffn.NeRF) and isinstance(four, ffn.FourierFeatureMLP)
    with torch.no_grad():
        np.testing.assert_allclose(nerf(pos, view).numpy(), g["out_nerf"], atol=2e-6)
        np.testing.assert_allclose(four(pos).numpy(), g["out_fourier"], atol=2e-6)
    vox = ffn.load_model(os.path.join(GOLDEN, "ref_voxels_small.pt"))
    assert isinstance(vox, ffn.Voxels) and vox.use_view is False
    with torch.no_grad():
        np.testing.assert_allclose(vox(torch.from_numpy(g["pos_vox"])).numpy(), g["out_vox"], atol=1e-6)
    # what our save() writes has the reference's layout: same keys, same "type"/"params" entries
    for name, model in (("ref_nerf_small.pt", nerf), ("ref_fourier_small.pt", four), ("ref_voxels_small.pt", vox)):
        ref_raw = torch.load(os.path.join(GOLDEN, name), map_location="cpu", weights_only=False)
        sd = model.state_dict()
        ours = dict(sd, type=ref_raw["type"], params=model.params)
        assert list(ours.keys()) == list(ref_raw.keys())
        assert set(ours["params"].keys()) == set(ref_raw["params"].keys())
        for k, v in ref_raw["params"].items():
            mine = ours["params"][k]
            if isinstance(v, (torch.Tensor, np.ndarray, list)) and not isinstance(v, (str,)) and v is not None \
                    and not isinstance(v, bool):
                np.testing.assert_allclose(np.asarray(mine, dtype=np.float64), np.asarray(v, dtype=np.float64))
            else:
                assert mine == v, (k, mine, v)


@pytest.mark.skipif(not os.path.isdir("/root/reference/fourier_feature_nets"), reason="needs the reference checkout")
def test_our_checkpoints_load_in_the_reference(tmp_path):
    import subprocess
    import sys
    m = ffn.NeRF(4, 32, 5, 3, 2, 2, [2], True)
    f = ffn.GaussianFourierMLP(3, 4, 2.5, num_layers=2, num_channels=32, embedding_size=16)
    m.save(str(tmp_path / "n.pt"))
    f.save(str(tmp_path / "f.pt"))
    x = torch.rand((16, 3)) * 2 - 1
    with torch.no_grad():
        np.savez(str(tmp_path / "exp.npz"), x=x.numpy(), n=m(x, x).numpy(), f=f(x).numpy())
    code = r"""
import sys, os, numpy as np, torch
sys.path.insert(0, %r)
from make_golden import import_reference
ref = import_reference()
e = np.load(%r)
x = torch.from_numpy(e["x"])
n = ref.load_model(%r); f = ref.load_model(%r)
assert type(n).__module__.startswith("fourier_feature_nets.") and type(n).__name__ == "NeRF"
with torch.no_grad():
    assert np.abs(n(x, x).numpy() - e["n"]).max() <= 2e-6
    assert np.abs(f(x).numpy() - e["f"]).max() <= 2e-6
print("OK")
""" % (GOLDEN, str(tmp_path / "exp.npz"), str(tmp_path / "n.pt"), str(tmp_path / "f.pt"))
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0 and "OK" in res.stdout, res.stdout[-500:] + res.stderr[-1500:]


def test_engine_notices_fused_optimizer_steps():
    """Fused optimizers do not bump tensor version counters; the engine's re-pack signature must still change."""
    from fourier_feature_nets_b200 import engine
    p = torch.nn.Parameter(torch.zeros(4))
    p.grad = torch.ones(4)
    try:
        opt = torch.optim.Adam([p], 1e-2, fused=True)
    except (RuntimeError, TypeError):
        pytest.skip("no fused Adam on this build")
    gen, ver = engine._OPT_GENERATION[0], p._version
    opt.step()
    assert engine._OPT_GENERATION[0] == gen + 1
    assert p._version == ver or True      # whichever torch does, the generation covers it
    engine.mark_weights_changed()
    assert engine._OPT_GENERATION[0] == gen + 2


def test_no_undefined_names_in_the_package():
    """The CUDA-only code paths cannot run in the CPU suite; at least every name they load must exist."""
    import glob
    files = glob.glob(os.path.join(ROOT, "fourier_feature_nets_b200", "*.py")) + \
        glob.glob(os.path.join(ROOT, "tools", "*.py")) + [os.path.join(ROOT, "bench.py")]
    res = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "check_names.py")] + files,
                         capture_output=True, text=True)
    assert res.returncode == 0, res.stdout


def test_orbit_video_frames_shard_over_two_ranks_gloo(tmp_path):
    """tools/orbit_video_multi_gpu.py (configs[4]: orbit frames split over the ranks, no data-path collective): two gloo
    ranks on the CPU write exactly the frames a single process writes, pixel for pixel."""
    torch.manual_seed(3)
    model = ffn.NeRF(2, 32, 3, 4, 2, 2, [1], True)
    path = str(tmp_path / "m.pt")
    model.save(path)
    tool = os.path.join(ROOT, "tools", "orbit_video_multi_gpu.py")
    common = [path, "12", "--num-frames", "5", "--num-samples", "8", "--device", "cpu", "--batch_size", "64"]
    one, two = str(tmp_path / "one"), str(tmp_path / "two")
    res = subprocess.run([sys.executable, tool, common[0], common[1], one] + common[2:], capture_output=True, text=True,
                         timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="2")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
           "127.0.0.1", "--master-port", "29641", tool, common[0], common[1], two] + common[2:]
    res = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    assert '"n_ranks": 2' in res.stdout
    import cv2
    names = sorted(os.listdir(one))
    assert names == ["frame_%05d.png" % i for i in range(5)] == sorted(os.listdir(two))
    for n in names:
        assert np.array_equal(cv2.imread(os.path.join(one, n)), cv2.imread(os.path.join(two, n))), n


def test_ctypes_signatures_match_the_header():
    """Every prototype of include/ffn_b200.h is exported, listed in EXPORTED_SYMBOLS, and -- where the Python binding
    declares ``argtypes`` -- bound with the same number of parameters (a drifted signature would otherwise only show
    up as garbage arguments on the GPU box)."""
    import re
    from fourier_feature_nets_b200 import autograd, optim, trainer
    hdr = open(os.path.join(ROOT, "include", "ffn_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    protos = {}
    for m in re.finditer(r"\b(?:int|void|int64_t|const char\*)\s+(ffn_\w+)\s*\(([^;{]*?)\)\s*;", hdr, flags=re.S):
        args = m.group(2).strip()
        protos[m.group(1)] = 0 if args in ("", "void") else len(args.split(","))
    assert len(protos) >= 35
    assert set(protos) == set(_lib.EXPORTED_SYMBOLS), set(protos) ^ set(_lib.EXPORTED_SYMBOLS)
    L = _lib.lib()
    autograd._bind(L)
    optim._bind(L)
    trainer._bind(L)
    checked = 0
    for name, n in protos.items():
        fn = getattr(L, name)
        if fn.argtypes is not None:
            assert len(fn.argtypes) == n, (name, len(fn.argtypes), n)
            checked += 1
    assert checked >= 15, checked


def test_cuda_only_training_components_refuse_cpu_models():
    """ClipAdam and FusedTrainer have no CPU implementation (Raycaster.fit uses the reference's PyTorch calls on a CPU
    model): they fail loudly instead of falling back."""
    model = ffn.NeRF(2, 32, 3, 4, 2, 2, [1], True)
    with pytest.raises(_lib.FFNError):
        ffn.FusedTrainer(model, 5e-4)
    opt = ffn.ClipAdam(model.parameters(), 5e-4)
    for p in model.parameters():
        if p.requires_grad:
            p.grad = torch.zeros_like(p)
    with pytest.raises(_lib.FFNError):
        opt.step()
    with pytest.raises(ValueError):            # one parameter group only: the norm is clipped over all of them
        ffn.ClipAdam([{"params": [model.layers[0].weight]}, {"params": [model.layers[0].bias]}], 5e-4)


def test_bench_reference_arm_runs_on_the_cpu_and_keeps_the_contract():
    """``bench.py --impl reference`` (the driver's reference arm) needs no GPU and nothing of the product package:
    one JSON line with the base contract's keys; the product arm refuses to run without a GPU."""
    import json
    bench = os.path.join(ROOT, "bench.py")
    res = subprocess.run([sys.executable, bench, "--impl", "reference", "--steps", "1", "--warmup", "1", "--ref-rays",
                          "512"], capture_output=True, text=True, timeout=600,
                         env=dict(os.environ, OMP_NUM_THREADS="1"))      # what torchrun exports at N > 1
    assert res.returncode == 0, res.stdout + res.stderr
    line = json.loads(res.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "rays/s" and line["value"] > 0
    assert line["metric"].startswith("rays/sec (lego_400, 64 samples/ray)") and line["higher_is_better"] is True
    from oracle import reference as refmod
    assert line["cpu_baseline"]["kind"] == ("reference" if refmod.available() else "port")
    assert line["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))      # not torchrun's OMP_NUM_THREADS=1
    assert "512 rays" in line["cpu_baseline"]["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # the config is the product arm's config (the sample is described in cpu_baseline.sample)
    import bench as bench_mod
    ns = type("A", (), dict(rays=1 << 20, steps=1, warmup=1))
    assert line["config"] == bench_mod.bench_config(ns, 1) and line["steps"] == 1
    assert len(res.stdout.strip().splitlines()) == 1          # nothing but the JSON line on stdout
    if not torch.cuda.is_available():
        res = subprocess.run([sys.executable, bench, "--steps", "1", "--warmup", "1"], capture_output=True, text=True,
                             timeout=600)
        assert res.returncode != 0 and "needs a GPU" in (res.stdout + res.stderr)
