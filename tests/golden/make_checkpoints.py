"""Checkpoint fixtures written by the REAL reference (run in the build container, /root/reference present):
    python tests/golden/make_checkpoints.py
Small models saved with the reference's own ``.save`` (nerf_model.py:126-135, fourier_feature_models.py:80-89)
plus their outputs on fixed inputs -> ref_nerf_small.pt, ref_fourier_small.pt, ref_voxels_small.pt, checkpoints.npz."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import import_reference  # noqa: E402


def main():
    ffn = import_reference()
    torch.manual_seed(7)
    nerf = ffn.NeRF(4, 32, 5, 3, 2, 2, [2], True)
    four = ffn.PositionalFourierMLP(3, 4, 4.0, num_layers=2, num_channels=32, embedding_size=24)
    g = torch.Generator().manual_seed(11)
    pos = torch.rand((64, 3), generator=g) * 2 - 1
    view = torch.nn.functional.normalize(torch.randn((64, 3), generator=g), dim=-1)
    with torch.no_grad():
        out_nerf = nerf(pos, view)
        out_four = four(pos)
    vox = ffn.Voxels(6, 1.2)
    with torch.no_grad():
        vox.voxels.copy_(torch.randn(vox.voxels.shape, generator=g))
        vox.bias.add_(torch.randn(vox.bias.shape, generator=g) * 0.1)
        pos_v = torch.rand((256, 3), generator=g) * 3.6 - 1.8      # inside, on the border cells and outside
        pos_v[:8] = torch.tensor([[1.2, 1.2, 1.2], [-1.2, -1.2, -1.2], [0, 0, 0], [1.0, -1.0, 1.0],
                                  [0.999, 0.2, -0.999], [5.0, 0, 0], [0, -7.0, 0], [0.1, 0.1, 9.0]])
        out_vox = vox(pos_v)
    vox.save(os.path.join(HERE, "ref_voxels_small.pt"))
    nerf.save(os.path.join(HERE, "ref_nerf_small.pt"))
    four.save(os.path.join(HERE, "ref_fourier_small.pt"))
    np.savez_compressed(os.path.join(HERE, "checkpoints.npz"), pos=pos.numpy(), view=view.numpy(),
                        out_nerf=out_nerf.numpy(), out_fourier=out_four.numpy(), pos_vox=pos_v.numpy(),
                        out_vox=out_vox.numpy())
    print("checkpoint fixtures written")


if __name__ == "__main__":
    main()
