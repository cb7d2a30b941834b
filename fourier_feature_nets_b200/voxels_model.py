"""``Voxels``: the reference's voxel-grid radiance field (fourier_feature_nets/voxels_model.py:9-57), the
README's coarse opacity model for hierarchical sampling.  Same parameters (``voxels (1,4,s,s,s)``,
``bias (1,4)``), ``params`` / ``save`` format and ``forward`` semantics.  On a CUDA device without autograd the
interpolation runs in ``ffn_voxels_forward``; with gradients (train_voxels.py) it is ``F.grid_sample``."""
import torch
import torch.nn as nn
import torch.nn.functional as F


class Voxels(nn.Module):
    """A voxel based radiance field model."""

    def __init__(self, side: int, scale: float):
        nn.Module.__init__(self)
        # plain Python numbers (train_voxels.py:99-101 passes a numpy scalar for ``scale``): the checkpoint then loads
        # under torch's safe unpickler
        side, scale = int(side), float(scale)
        self.params = {"side": side, "scale": scale}
        self.voxels = nn.Parameter(torch.zeros((1, 4, side, side, side), dtype=torch.float32))
        bias = torch.zeros(4, dtype=torch.float32)
        bias[:3] = torch.logit(torch.FloatTensor([1e-5, 1e-5, 1e-5]))
        bias[3] = -2
        self.bias = nn.Parameter(bias.unsqueeze(0))
        self.scale = scale
        self.use_view = False
        self._packed = None          # (version, channels-last copy of the grid, bias floats)

    def forward_torch(self, positions: torch.Tensor) -> torch.Tensor:
        grid = (positions.reshape(1, -1, 1, 1, 3) / self.scale)
        output = F.grid_sample(self.voxels, grid, padding_mode="border", align_corners=False)
        output = output.transpose(1, 2).reshape(-1, 4) + self.bias
        assert not output.isnan().any()
        return output

    def forward(self, positions: torch.Tensor) -> torch.Tensor:
        """Interpolates the positions within the voxel volume -> (N, 4) raw [rgb | sigma]."""
        needs_grad = torch.is_grad_enabled() and (self.voxels.requires_grad or self.bias.requires_grad
                                                  or positions.requires_grad)
        if not positions.is_cuda or not self.voxels.is_cuda or needs_grad:
            return self.forward_torch(positions)
        from . import _lib
        from .engine import _OPT_GENERATION
        # fused optimisers (ClipAdam, torch's fused Adam) update parameters without bumping the version counters:
        # any optimiser step since the last copy invalidates it (same rule as engine.Engine.sync_weights)
        key = (_OPT_GENERATION[0], self.voxels._version, self.bias._version, self.voxels.data_ptr())
        if self._packed is None or self._packed[0] != key:
            self._packed = (key, self.voxels.detach()[0].permute(1, 2, 3, 0).contiguous(),
                            self.bias.detach().reshape(4).tolist())
        return _lib.voxels_forward(self._packed[1], self._packed[2], self.scale, positions)

    def save(self, path: str):
        state_dict = self.state_dict()
        state_dict["type"] = "voxels"
        state_dict["params"] = self.params
        torch.save(state_dict, path)
