"""B200-native volume-rendering hot path with the API of ``fourier_feature_nets``.

Public names mirror the reference package (fourier_feature_nets/__init__.py:3-35) for the
render/training path: ``Raycaster``, ``RaySampler``, ``RaySamples``, ``NeRF``,
``FourierFeatureMLP`` (+ presets), ``CameraInfo``, ``Resolution``, ``RenderResult``,
``calculate_blend_weights``, ``exponential_lr_decay``, ``load_model``, ``orbit``.
"""
from .camera_info import CameraInfo, Resolution
from .fourier_feature_models import (BasicFourierMLP, FourierFeatureMLP, GaussianFourierMLP, MLP,
                                     PositionalFourierMLP)
from .image_dataset import ImageDataset, RayDataset
from .nerf_model import NeRF
from .optim import ClipAdam
from .pixel_dataset import PixelData, PixelDataset
from .signal_dataset import SignalData, SignalDataset
from .ray_caster import Raycaster
from .trainer import FusedTrainer
from .ray_dataset_modes import Mode
from .ray_sampler import FocusBundle, RayBundle, RaySampler, RaySamples
from .utils import (ETABar, RenderResult, calculate_blend_weights, exponential_lr_decay, linspace,
                    load_model, orbit)
from .voxels_model import Voxels
from .visualizers import (ActivationVisualizer, ComparisonVisualizer, EvaluationVisualizer,
                          OrbitVideoVisualizer)

RayCaster = Raycaster   # BASELINE.json spells it this way; the reference class is ``Raycaster``

__version__ = "0.1.0"

__all__ = ["CameraInfo", "Resolution", "MLP", "NeRF", "BasicFourierMLP", "FourierFeatureMLP",
           "PositionalFourierMLP", "GaussianFourierMLP", "Raycaster", "RayCaster", "RaySampler",
           "RaySamples", "RayBundle", "FocusBundle", "RenderResult", "Mode", "ImageDataset", "RayDataset", "calculate_blend_weights",
           "exponential_lr_decay", "linspace", "load_model", "orbit", "ETABar", "EvaluationVisualizer",
           "OrbitVideoVisualizer", "ActivationVisualizer", "ComparisonVisualizer", "Voxels", "ClipAdam", "FusedTrainer", "PixelDataset", "PixelData", "SignalDataset", "SignalData", "__version__"]
