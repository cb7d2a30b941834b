"""Config 0 (train_image_regression.py, CPU): ``PixelDataset`` and three full-batch training steps of the gaussian
preset against vectors recorded from the real reference (tests/golden/make_pixel_golden.py).  Data tensors are
bit-exact; the training steps agree to fp32/fp64 rounding of the same op sequence (1e-6 relative)."""
import os

import cv2
import numpy as np
import torch

import fourier_feature_nets_b200 as ffn

from conftest import ROOT

G = np.load(os.path.join(ROOT, "tests", "golden", "pixel.npz"))


def _image(tmp_path):
    path = str(tmp_path / "in.png")
    with open(path, "wb") as f:
        f.write(G["png"].tobytes())
    assert cv2.imread(path).shape == (46, 70, 3)
    return path


def test_create_matches_reference_bitwise(tmp_path):
    path = _image(tmp_path)
    for cs in ("RGB", "YCrCb"):
        ds = ffn.PixelDataset.create(path, cs, 32)
        assert ds.train_color.dtype == torch.float64 and ds.train_uv.dtype == torch.float32
        for name in ("train_uv", "train_color", "val_uv", "val_color"):
            assert np.array_equal(getattr(ds, name).numpy(), G[cs + "_" + name]), (cs, name)
        assert np.array_equal(ds.image, G[cs + "_image"])
        pred = torch.from_numpy(G[cs + "_pred"])
        assert ds.psnr(pred.reshape(32, 32, 3)) == float(G[cs + "_psnr"])
        assert np.array_equal(ds.to_image(pred), G[cs + "_to_image"])
        moved = ds.to("cpu")
        assert np.array_equal(moved.val_color.numpy(), G[cs + "_val_color"]) and moved.size == 32
    assert np.array_equal(ffn.PixelDataset.generate_uvs(8, "cpu").numpy(), G["uvs_8"])
    assert ffn.PixelDataset.create(str(tmp_path / "missing.png"), "RGB", 32) is None


def test_three_training_steps_match_reference(tmp_path):
    ds = ffn.PixelDataset.create(_image(tmp_path), "RGB", 32)
    torch.manual_seed(11)
    model = ffn.GaussianFourierMLP(2, 3, sigma=10, num_channels=32, embedding_size=16)
    for k, v in model.state_dict().items():      # same construction order -> same initial weights
        assert np.array_equal(v.numpy(), G["w0_" + k]), k
    optim = torch.optim.Adam(model.parameters(), 1e-3)
    losses = []
    for step in range(3):
        ffn.exponential_lr_decay(optim, 1e-3, step, 0.1, 2500)
        optim.zero_grad()
        output = torch.sigmoid(model(ds.train_uv))
        loss = 0.5 * torch.square(output - ds.train_color).mean()
        loss.backward()
        optim.step()
        losses.append(loss.item())
    assert np.allclose(losses, G["losses"], rtol=1e-6, atol=0)
    for k, v in model.state_dict().items():
        assert np.allclose(v.numpy(), G["w3_" + k], rtol=1e-5, atol=1e-7), k


def test_activation_mosaic_shape(tmp_path):
    ds = ffn.PixelDataset.create(_image(tmp_path), "RGB", 32)
    torch.manual_seed(0)
    model = ffn.BasicFourierMLP(2, 3, num_channels=64)
    img = ds.to_act_image(model, 64)
    assert img.shape == (64, 64, 3) and img.dtype == np.uint8 and not model.keep_activations


def test_signal_dataset_create():
    """signal_dataset.py:39-67: x = linspace(0, 2, n * rate, endpoint=False) float32, every rate-th point trains."""
    ds = ffn.SignalDataset.create(lambda x: np.sin(3 * np.pi * x), 16, 8)
    assert ds.val_x.shape == (128, 1) and ds.train_x.shape == (16, 1) and ds.val_x.dtype == torch.float32
    assert np.array_equal(ds.train_x.numpy(), ds.val_x.numpy()[::8]) and np.array_equal(ds.train_y.numpy(), ds.val_y.numpy()[::8])
    assert ds.val_x[0, 0].item() == 0.0 and abs(ds.val_x[-1, 0].item() - (2 - 2 / 128)) < 1e-6
    lo, hi = ds.x_lim
    assert lo < 0 < 2 - 2 / 128 < hi and abs((hi - lo) - 1.1 * (2 - 2 / 128)) < 1e-5
