"""The reference's own scripts, UNCHANGED, against this package (north_star: "train_nerf.py and orbit_video.py
run unchanged").  Needs the read-only reference checkout, so it is skipped where that is absent (GPU box)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT

REF = os.environ.get("FFN_REFERENCE", "/root/reference")
pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(REF, "train_nerf.py")),
                                reason="reference checkout not available")


def run(args, cwd):
    cmd = [sys.executable, os.path.join(ROOT, "tools", "run_reference_script.py")] + args
    res = subprocess.run(cmd, cwd=cwd, capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    return res.stdout


def test_train_nerf_then_orbit_video_unchanged(tmp_path):
    data = str(tmp_path / "toy.npz")
    subprocess.run([sys.executable, os.path.join(ROOT, "tools", "make_synthetic_dataset.py"), data,
                    "--resolution", "24", "--train", "6", "--val", "2", "--test", "1", "--steps", "48"],
                   check=True, capture_output=True, timeout=300)
    out = str(tmp_path / "out")
    log = run([os.path.join(REF, "train_nerf.py"), data, out, "--device", "cpu", "--num-steps", "3",
               "--batch-size", "128", "--num-samples", "8", "--image-interval", "2", "--report-interval", "2",
               "--crop-steps", "0", "--num-layers", "2"], str(tmp_path))
    assert "psnr_train" in log
    assert os.path.exists(os.path.join(out, "nerf.pt")) and os.path.exists(os.path.join(out, "log.txt"))
    assert any(f.endswith(".png") for f in os.listdir(os.path.join(out, "train")))
    frames = str(tmp_path / "orbit")
    run([os.path.join(REF, "orbit_video.py"), os.path.join(out, "nerf.pt"), "12", frames, "--num-frames", "2",
         "--num-samples", "8", "--device", "cpu", "--batch_size", "128"], str(tmp_path))
    assert sorted(os.listdir(frames)) == ["frame_00000.png", "frame_00001.png"]


def _toy(tmp_path):
    data = str(tmp_path / "toy.npz")
    subprocess.run([sys.executable, os.path.join(ROOT, "tools", "make_synthetic_dataset.py"), data,
                    "--resolution", "24", "--train", "6", "--val", "2", "--test", "1", "--steps", "48"],
                   check=True, capture_output=True, timeout=300)
    return data


def test_train_tiny_nerf_unchanged(tmp_path):
    """BASELINE.json configs[1]: train_tiny_nerf.py (FourierFeatureMLP presets through Raycaster.fit)."""
    data = _toy(tmp_path)
    for preset in ("positional", "gaussian"):
        out = str(tmp_path / preset)
        log = run([os.path.join(REF, "train_tiny_nerf.py"), data, preset, out, "--device", "cpu", "--num-steps", "3",
                   "--batch-size", "128", "--num-samples", "8", "--image-interval", "2", "--report-interval", "2",
                   "--crop-steps", "0", "--num-anneal-steps", "2", "--num-channels", "32", "--embedding-size", "16"],
                  str(tmp_path))
        assert "psnr_train" in log
        assert os.path.exists(os.path.join(out, "tiny_nerf.pt"))


def test_train_voxels_unchanged(tmp_path):
    data = _toy(tmp_path)
    out = str(tmp_path / "vox")
    log = run([os.path.join(REF, "train_voxels.py"), data, "16", out, "--device", "cpu", "--num-steps", "3",
               "--batch-size", "128", "--num-samples", "8", "--image-interval", "2", "--report-interval", "2",
               "--num-cameras", "4"], str(tmp_path))
    assert "psnr_train" in log and os.path.exists(os.path.join(out, "voxels.pt"))


def test_train_image_regression_unchanged(tmp_path):
    """BASELINE.json configs[0]: train_image_regression.py cat.jpg, gaussian FourierFeatureMLP on CPU."""
    out = str(tmp_path / "img")
    log = run([os.path.join(REF, "train_image_regression.py"), os.path.join(REF, "data", "cat.jpg"), "gaussian", out,
               "--device", "cpu", "--image-size", "32", "--num-steps", "4", "--report-interval", "2",
               "--num-channels", "32", "--embedding_size", "16"], str(tmp_path))
    assert "step 4 val:" in log
    for name in ("val00000.png", "val00004.png", "superres.png", "model.pt"):
        assert os.path.exists(os.path.join(out, name)), name
    model = __import__("fourier_feature_nets_b200").load_model(os.path.join(out, "model.pt"))
    assert model.layers[0].in_features == 32
