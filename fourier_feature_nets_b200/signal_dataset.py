"""``SignalDataset``: the 1-D regression data of the lecture demo (``train_signal_regression.py``), mirroring
fourier_feature_nets/signal_dataset.py:25-127.  Outside the CUDA hot path; kept so that the package surface of the
reference is complete.  ``plot`` needs matplotlib (not installed in the build image) and imports it lazily."""
from typing import Callable, NamedTuple

import numpy as np
import torch

SignalData = NamedTuple("SignalData", [("x", torch.Tensor), ("y", torch.Tensor)])


def _stretched_range(values, stretch=1.1):
    """(min, max) widened about the midpoint (signal_dataset.py:17-22)."""
    lo, hi = values.min().item(), values.max().item()
    mid = 0.5 * (lo + hi)
    return mid + stretch * (lo - mid), mid + stretch * (hi - mid)


class SignalDataset:
    """x in [0, 2) sampled densely for validation, every ``sample_rate``-th point for training."""

    def __init__(self, train_data: SignalData, val_data: SignalData):
        self.train_x, self.train_y = train_data
        self.val_x, self.val_y = val_data
        self.x_lim = _stretched_range(self.val_x)
        self.y_lim = _stretched_range(self.val_y)

    @staticmethod
    def create(signal: Callable[[np.ndarray], np.ndarray], num_samples: int, sample_rate: int) -> "SignalDataset":
        x = np.linspace(0, 2, num_samples * sample_rate, endpoint=False).astype(np.float32)
        y = signal(x)
        x, y = x.reshape(-1, 1), y.reshape(-1, 1)
        pick = slice(None, None, sample_rate)
        return SignalDataset(SignalData(torch.from_numpy(x[pick]), torch.from_numpy(y[pick])),
                             SignalData(torch.from_numpy(x), torch.from_numpy(y)))

    def plot(self, space_ax, hidden_ax, model, num_points: int, colors: np.ndarray, max_hidden: int):
        """Reconstruction on ``space_ax``; on ``hidden_ax`` the ``max_hidden`` last-layer activations with the largest
        range, scaled by the output weights and shifted by the output bias (signal_dataset.py:69-127)."""
        import matplotlib.pyplot as plt
        x_vals = torch.linspace(self.val_x[0, 0], self.val_x[-1, 0], num_points)
        model.eval()
        model.keep_activations = True
        with torch.no_grad():
            y_vals = model(x_vals.reshape(-1, 1)).reshape(-1).cpu().numpy()
        model.keep_activations = False
        model.train()
        out = model.layers[-1]
        act = np.asarray(model.activations[-1])
        scaled = act * out.weight.detach().cpu().numpy().reshape(1, -1) + out.bias.item()
        order = np.argsort(scaled.max(0) - scaled.min(0))[::-1][:max_hidden]
        cmap = plt.get_cmap("jet")
        for rank, unit in enumerate(order):
            live = act[:, unit] > 0
            hidden_ax.plot(x_vals, scaled[:, unit], color=cmap(rank / max_hidden)[:3], zorder=1,
                           label="h{:02d}".format(unit))
            hidden_ax.scatter(x_vals[live], scaled[live, unit], color=colors[live], marker=".", zorder=2)
        hidden_ax.set_ylim(*_stretched_range(scaled[act > 0]))
        hidden_ax.legend(loc="upper right", ncol=2)
        space_ax.set_xlim(*self.x_lim)
        space_ax.set_ylim(*self.y_lim)
        space_ax.plot(self.val_x.numpy(), self.val_y.numpy(), "r-", label="val", zorder=1)
        space_ax.plot(self.train_x.numpy(), self.train_y.numpy(), "go", label="train", zorder=2)
        space_ax.scatter(x_vals.numpy(), y_vals, color=colors, marker="P", label="pred", zorder=3)
        space_ax.legend()
