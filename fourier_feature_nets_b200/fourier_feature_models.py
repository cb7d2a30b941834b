"""``FourierFeatureMLP`` and its four presets with the reference's constructors,
parameter names and ``.pt`` format (reference:
fourier_feature_nets/fourier_feature_models.py:10-191)."""
import math
from typing import List

import torch
import torch.nn as nn

from . import engine as _engine
from .nerf_model import _needs_grad, positional_frequencies


class FourierFeatureMLP(nn.Module):
    """``[a cos(pi x B) | a sin(pi x B)]`` followed by a ReLU MLP."""

    _ffn_kind = "fourier"

    def __init__(self, num_inputs: int, num_outputs: int,
                 a_values: torch.Tensor, b_values: torch.Tensor,
                 layer_channels: List[int]):
        super().__init__()
        encoded = b_values is not None
        # the constructor arguments, JSON-friendly: ``save`` / ``load_model`` rebuild the model from them
        self.params = dict(num_inputs=num_inputs, num_outputs=num_outputs,
                           a_values=a_values.tolist() if a_values is not None else None,
                           b_values=b_values.tolist() if encoded else None,
                           layer_channels=layer_channels)
        self.num_inputs = num_inputs
        self.use_view = False
        self.keep_activations = False
        self.activations = []
        if encoded:
            if b_values.shape[0] != num_inputs or a_values.shape[0] != b_values.shape[1]:
                raise AssertionError("b_values must be (num_inputs, E) and a_values (E,)")
            # frozen: they travel with the state dict but are never optimised
            self.a_values = nn.Parameter(a_values, requires_grad=False)
            self.b_values = nn.Parameter(b_values, requires_grad=False)
        else:
            self.a_values = self.b_values = None
        widths = [2 * b_values.shape[1] if encoded else num_inputs] + list(layer_channels) + [num_outputs]
        self.layers = nn.ModuleList(nn.Linear(w_in, w_out) for w_in, w_out in zip(widths, widths[1:]))

    def forward_torch(self, inputs: torch.Tensor) -> torch.Tensor:
        if self.b_values is None:
            h = inputs
        else:
            # pi (not 2 pi): inputs already span [-1,1] / [0,2]  (fourier_feature_models.py:62-66)
            e = (math.pi * inputs) @ self.b_values
            h = torch.cat([self.a_values * e.cos(), self.a_values * e.sin()], dim=-1)
        self.activations.clear()
        for layer in self.layers[:-1]:
            h = torch.relu(layer(h))
        if self.keep_activations:
            self.activations.append(h.detach().cpu().numpy())
        return self.layers[-1](h)

    def _engine_ok(self) -> bool:
        return (self.num_inputs == 3 and len(self.layers) >= 2
                and self.layers[-1].out_features == 4
                and all(l.out_features == 256 for l in list(self.layers)[:-1])
                and (self.b_values is None or self.b_values.shape[1] <= 256))

    def forward(self, inputs: torch.Tensor) -> torch.Tensor:
        if (inputs.is_cuda and not self.keep_activations and inputs.shape[-1] == 3
                and self._engine_ok() and not _needs_grad(self, inputs)):
            eng = _engine.get_engine(self, inputs.device)
            return eng.net.mlp_forward(inputs.reshape(-1, 3), None).reshape(*inputs.shape[:-1], 4)
        return self.forward_torch(inputs)

    def __getstate__(self):
        """``copy.deepcopy`` / pickling: the C handles bound to this instance (engine, trainer, flat gradient buffer)
        stay behind; the copy builds its own on first use."""
        state = self.__dict__.copy()
        for key in [k for k in state if k.startswith("_ffn_")]:
            del state[key]
        return state

    def save(self, path: str):
        state_dict = self.state_dict()
        state_dict["type"] = "fourier"
        state_dict["params"] = self.params
        torch.save(state_dict, path)


def _hidden(num_layers: int, num_channels: int) -> List[int]:
    return [num_channels] * num_layers


class MLP(FourierFeatureMLP):
    """Un-encoded MLP."""

    def __init__(self, num_inputs: int, num_outputs: int, num_layers=3, num_channels=256):
        super().__init__(num_inputs, num_outputs, None, None, _hidden(num_layers, num_channels))


class BasicFourierMLP(FourierFeatureMLP):
    """Inputs projected onto the unit circle (B = I, a = 1)."""

    def __init__(self, num_inputs: int, num_outputs: int, num_layers=3, num_channels=256):
        super().__init__(num_inputs, num_outputs, torch.ones(num_inputs), torch.eye(num_inputs),
                         _hidden(num_layers, num_channels))


class PositionalFourierMLP(FourierFeatureMLP):
    """Axis-aligned log-spaced frequencies."""

    def __init__(self, num_inputs: int, num_outputs: int, max_log_scale: float,
                 num_layers=3, num_channels=256, embedding_size=256):
        freqs = self._encoding(max_log_scale, embedding_size, num_inputs)
        super().__init__(num_inputs, num_outputs, torch.ones(freqs.shape[1]), freqs, _hidden(num_layers, num_channels))

    @staticmethod
    def _encoding(max_log_scale: float, embedding_size: int, num_inputs: int):
        return positional_frequencies(max_log_scale, embedding_size // num_inputs, num_inputs)


class GaussianFourierMLP(FourierFeatureMLP):
    """Dense Gaussian frequency matrix."""

    def __init__(self, num_inputs: int, num_outputs: int, sigma: float,
                 num_layers=3, num_channels=256, embedding_size=256):
        freqs = torch.normal(0, sigma, size=(num_inputs, embedding_size))
        super().__init__(num_inputs, num_outputs, torch.ones(embedding_size), freqs, _hidden(num_layers, num_channels))
