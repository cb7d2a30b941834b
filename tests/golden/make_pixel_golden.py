"""Golden vectors for ``PixelDataset`` (config 0, train_image_regression.py) from the REAL reference.
Run in the build container only:  python tests/golden/make_pixel_golden.py  ->  tests/golden/pixel.npz
The input image is synthetic (seeded noise + gradients, non-square so the centre crop is exercised) and is stored
PNG-encoded inside the fixture."""
import os
import sys
import tempfile

import cv2
import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import import_reference  # noqa: E402


def main():
    ffn = import_reference()
    rng = np.random.default_rng(7)
    h, w = 46, 70
    yy, xx = np.mgrid[0:h, 0:w]
    img = np.stack([xx * 3 % 256, yy * 5 % 256, (xx + yy) * 2 % 256], -1).astype(np.float32)
    img = np.clip(img + rng.normal(0, 20, img.shape), 0, 255).astype(np.uint8)
    ok, png = cv2.imencode(".png", img)
    assert ok
    out = {"png": png}
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "in.png")
        with open(path, "wb") as f:
            f.write(png.tobytes())
        for cs in ("RGB", "YCrCb"):
            ds = ffn.PixelDataset.create(path, cs, 32)
            out[cs + "_train_uv"] = ds.train_uv.numpy()
            out[cs + "_train_color"] = ds.train_color.numpy()
            out[cs + "_val_uv"] = ds.val_uv.numpy()
            out[cs + "_val_color"] = ds.val_color.numpy()
            out[cs + "_image"] = ds.image
            g = torch.Generator().manual_seed(3)
            pred = torch.rand((32 * 32, 3), generator=g, dtype=torch.float64)
            out[cs + "_pred"] = pred.numpy()
            out[cs + "_psnr"] = np.float64(ds.psnr(pred.reshape(32, 32, 3)))
            out[cs + "_to_image"] = ds.to_image(pred)
        out["uvs_8"] = ffn.PixelDataset.generate_uvs(8, "cpu").numpy()
        # one full-batch training step of the gaussian preset on the RGB data, as train_image_regression.py:180-185
        torch.manual_seed(11)
        model = ffn.GaussianFourierMLP(2, 3, sigma=10, num_channels=32, embedding_size=16)
        ds = ffn.PixelDataset.create(path, "RGB", 32)
        out["model_b"] = model.b_values.numpy() if hasattr(model, "b_values") else model.params["b_values"]
        for k, v in model.state_dict().items():
            out["w0_" + k] = v.detach().clone().numpy()
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
        out["losses"] = np.asarray(losses, np.float64)
        for k, v in model.state_dict().items():
            out["w3_" + k] = v.numpy()
    np.savez_compressed(os.path.join(HERE, "pixel.npz"), **out)
    print("wrote pixel.npz", {k: getattr(v, "shape", None) for k, v in out.items()})


if __name__ == "__main__":
    main()
