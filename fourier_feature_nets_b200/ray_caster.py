"""``Raycaster`` with the reference's API (fourier_feature_nets/ray_caster.py:36-376).

``render`` on CUDA without autograd is ONE launch of the fused sm_100a kernel
(sampling -> encoding -> MLP on tcgen05 -> compositing); under autograd it is the
plain differentiable definition of the same maths (training kernels: DESIGN.md).
"""
import copy
import time
from typing import List, NamedTuple, Optional, OrderedDict

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib
from . import engine as _engine
from .nerf_model import _needs_grad
from .ray_sampler import FocusBundle, RayBundle, RaySampler, RaySamples
from .utils import RenderResult, blend_weights_torch, exponential_lr_decay

LogEntry = NamedTuple("LogEntry", [("step", int), ("timestamp", float),
                                   ("state", OrderedDict[str, torch.Tensor]),
                                   ("train_psnr", float), ("val_psnr", float)])


class Raycaster(nn.Module):
    """Differentiable volumetric raycaster around a radiance-field model."""

    def __init__(self, model: nn.Module):
        super().__init__()
        self.model = model
        self.check_nan_every_call = False   # True: reference behaviour (a host sync per render)
        self.train_kernels = True           # False: differentiate the plain PyTorch definition instead
        self.fused_trainer = True           # False: fit() steps through autograd even where the C trainer applies
        self._lin_cache = {}

    # ---- the hot path -----------------------------------------------------------------
    def _lin(self, num_samples: int, device) -> torch.Tensor:
        key = (num_samples, str(device))
        if key not in self._lin_cache:
            self._lin_cache[key] = torch.linspace(0, 1, num_samples).to(device)
        return self._lin_cache[key]

    def render(self, ray_samples: RaySamples, include_depth=False) -> RenderResult:
        """Render the ray samples -> per-ray colour (R,3), alpha (R), depth (R)."""
        device = next(self.model.parameters()).device
        needs_grad = _needs_grad(self.model)
        eng_ok = device.type == "cuda" and _engine.supported(self.model)      # kind AND shape
        fused = eng_ok and not needs_grad
        if device.type == "cuda" and not eng_ok:
            _engine.note_unfused(self.model)       # warns once (FFN_STRICT=1: raises); Voxels etc. are silent
        if needs_grad and eng_ok and self.train_kernels:
            from .autograd import render_nerf_train
            if isinstance(ray_samples, FocusBundle):     # t values come from the (frozen) coarse model
                with torch.no_grad():
                    ray_samples = ray_samples.materialize()
            color, alpha, depth = render_nerf_train(self.model, ray_samples, include_depth,
                                                    lambda S: self._lin(S, device))
            return RenderResult(color, alpha, depth)
        if not fused:
            if device.type == "cuda" and not needs_grad and not eng_ok:
                # a model family without a fused engine (Voxels, voxels_model.py:35-45, or any nn.Module the caller
                # brings): its own forward (Voxels: ffn_voxels_forward) + the compositing kernel
                return self._render_composite_kernel(ray_samples, include_depth)
            return self._render_torch(ray_samples, include_depth)

        eng = _engine.get_engine(self.model, device)
        if isinstance(ray_samples, FocusBundle):
            b = ray_samples
            t = b.focus_t()                  # coarse sigma pass + CDF + inverse transform + sort, on the GPU
            color, alpha, depth = eng.net.render_rays_t(b.starts, b.directions, t, include_depth)
        elif isinstance(ray_samples, RayBundle):
            b = ray_samples
            color, alpha, depth, _ = eng.net.render_rays(
                b.starts, b.directions, b.near, b.far, self._lin(b.num_samples, device), b.jitter,
                b.stratified, b.seed, b.ray_offset, b.num_samples, include_depth)
        else:
            views = ray_samples.view_directions if self.model.use_view else None
            color, alpha, depth = eng.net.render_samples(ray_samples.positions, views,
                                                         ray_samples.t_values, include_depth)
        if self.check_nan_every_call:
            self.check_nan()
        return RenderResult(color, alpha, depth)

    def check_nan(self):
        """Raise like the reference's asserts (ray_caster.py:73-74) if any render since the
        last check produced a NaN colour or opacity.  One device->host read."""
        flag = self.__dict__.get("_nan_flag")          # models rendered through forward + ffn_composite
        if flag is not None:
            v = int(flag.item())
            if v:
                flag.zero_()
            assert not (v & 1), "NaN in rendered colour/opacity"
        for eng in list(self.model.__dict__.get("_ffn_engines", {}).values()):
            flag = eng.net.nan_flag()
            assert not (flag & 0x40000000), "libffn_b200: shared memory base is not 1024-byte aligned"
            assert not (flag & 1), "NaN in rendered colour/opacity"

    def _render_composite_kernel(self, ray_samples: RaySamples, include_depth: bool) -> RenderResult:
        """model(positions[, views]) followed by ``ffn_composite`` (ray_caster.py:67-93 in one launch)."""
        if isinstance(ray_samples, RayBundle):
            ray_samples = ray_samples.materialize()
        num_rays, num_samples = ray_samples.positions.shape[:2]
        positions = ray_samples.positions.reshape(-1, 3)
        if getattr(self.model, "use_view", False):
            raw = self.model(positions, ray_samples.view_directions.reshape(-1, 3))
        else:
            raw = self.model(positions)
        eng_flag = self.__dict__.setdefault("_nan_flag", torch.zeros(1, dtype=torch.int32, device=raw.device))
        color, alpha, depth, _ = _lib.composite(raw.reshape(num_rays, num_samples, 4), ray_samples.t_values,
                                                include_depth, False, eng_flag)
        if self.check_nan_every_call:
            assert not (int(eng_flag.item()) & 1), "NaN in rendered colour/opacity"
        return RenderResult(color, alpha, depth)

    def _render_torch(self, ray_samples: RaySamples, include_depth: bool) -> RenderResult:
        """Differentiable definition (ray_caster.py:60-93)."""
        num_rays, num_samples = ray_samples.positions.shape[:2]
        positions = ray_samples.positions.reshape(-1, 3)
        if self.model.use_view:
            raw = self.model(positions, ray_samples.view_directions.reshape(-1, 3))
        else:
            raw = self.model(positions)
        raw = raw.reshape(num_rays, num_samples, 4)
        color = torch.sigmoid(raw[..., :3])
        opacity = F.softplus(raw[..., 3])
        assert not color.isnan().any()
        assert not opacity.isnan().any()
        weights = blend_weights_torch(ray_samples.t_values, opacity)
        out_color = (weights.unsqueeze(-1) * color).sum(-2)
        weights = weights[:, :-1]
        out_alpha = weights.sum(-1)
        out_depth = None
        if include_depth:
            cutoff = weights.argmax(-1)
            cutoff[out_alpha < .1] = -1
            out_depth = ray_samples.t_values[torch.arange(num_rays), cutoff]
        return RenderResult(out_color, out_alpha, out_depth)

    def _loss(self, step: int, dataset, batch: List[int]) -> torch.Tensor:
        device = next(self.model.parameters()).device
        rays = dataset.get_rays(batch, step).to(device)
        return dataset.loss(step, rays, self.render(rays, True))

    # ---- inference --------------------------------------------------------------------
    def batched_render(self, samples: RaySamples, batch_size: int, include_depth: bool) -> RenderResult:
        """Render in ray batches; returns numpy arrays (ray_caster.py:103-138).  Results stay on
        the device until the end: one device->host copy per call instead of one per batch."""
        self.model.eval()
        colors, alphas, depths = [], [], []
        with torch.no_grad():
            device = next(self.model.parameters()).device
            num_rays = samples.num_rays if isinstance(samples, RayBundle) else len(samples.positions)
            for start in range(0, num_rays, batch_size):
                end = min(start + batch_size, num_rays)
                batch = samples.subset(range(start, end)) if isinstance(samples, RayBundle) \
                    else samples.subset(list(range(start, end)))
                pred = self.render(batch.to(device), include_depth)
                colors.append(pred.color)
                alphas.append(pred.alpha)
                if include_depth:
                    depths.append(pred.depth)
            self.check_nan()
        self.model.train()
        return RenderResult(torch.cat(colors).cpu().numpy(), torch.cat(alphas).cpu().numpy(),
                            torch.cat(depths).cpu().numpy() if include_depth else None)

    def render_stream(self, sampler: RaySampler, index_batches, include_depth: bool = False, step=None):
        """Generator over ``RenderResult``s (numpy) of consecutive ray-index batches, host buffers in and out:
        ``sampler.sample(idx, step)`` -> host->device copy -> fused kernel -> device->host copy of the pixels, with the
        stages of neighbouring batches overlapped -- while the GPU renders batch i the host gathers batch i + 1 from the
        ray tables (into pinned staging when the sampler has it: ``RaySampler.enable_pinned_staging``) and unpacks the
        pixels of batch i - 1.  Equivalent to ``[self.render(sampler.sample(b, step).to(device), d).numpy() for b in
        index_batches]`` (what ``batched_render`` does per frame, ray_caster.py:122-131), minus the idle gaps."""
        device = next(self.model.parameters()).device
        if device.type != "cuda":
            for idx in index_batches:
                with torch.no_grad():
                    yield self.render(sampler.sample(idx, step).to(device), include_depth).numpy()
            return
        pending = None      # (event, pinned color, alpha, depth)
        pool = []           # two sets of pinned result buffers, rotating

        def pinned(i, n):
            while len(pool) <= i:
                pool.append(None)
            if pool[i] is None or pool[i][0].shape[0] < n:
                pool[i] = (torch.empty((n, 3), dtype=torch.float32).pin_memory(),
                           torch.empty((n,), dtype=torch.float32).pin_memory(),
                           torch.empty((n,), dtype=torch.float32).pin_memory())
            return pool[i]

        def unpack(p):
            ev, n, bufs = p
            ev.synchronize()
            return RenderResult(bufs[0][:n].numpy().copy(), bufs[1][:n].numpy().copy(),
                                bufs[2][:n].numpy().copy() if include_depth else None)

        with torch.no_grad():
            for i, idx in enumerate(index_batches):
                bundle = sampler.sample(idx, step)                      # host work, overlaps the previous launch
                pred = self.render(bundle.to(device, non_blocking=True), include_depth)
                n = pred.color.shape[0]
                bufs = pinned(i & 1, n)
                bufs[0][:n].copy_(pred.color, non_blocking=True)
                bufs[1][:n].copy_(pred.alpha, non_blocking=True)
                if include_depth:
                    bufs[2][:n].copy_(pred.depth, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record()
                if pending is not None:
                    yield unpack(pending)
                pending = (ev, n, bufs)
            if pending is not None:
                yield unpack(pending)
            self.check_nan()

    def render_image(self, sampler: RaySampler, index: int, batch_size: int,
                     color_space="RGB") -> np.ndarray:
        """Render camera ``index % num_cameras`` to an (H,W,3) uint8 image."""
        camera = index % sampler.num_cameras
        samples = sampler.rays_for_camera(camera)
        pred = self.batched_render(samples, batch_size, False)
        return sampler.to_image(camera, pred.color, color_space)

    # ---- training driver (ray_caster.py:220-376) --------------------------------------------
    def _validate(self, dataset, batch_size: int, step: int) -> float:
        num_rays = len(dataset)
        num_validate = min(num_rays, 1024 * 100)
        if num_validate < num_rays:
            val_index = np.linspace(0, num_rays, num_validate, endpoint=False).astype(np.int32)
            val_index = dataset.to_valid(val_index.tolist())
        else:
            val_index = np.arange(num_rays)
        self.model.eval()
        losses = []
        with torch.no_grad():
            for start in range(0, num_validate, batch_size):
                if start + batch_size > len(val_index):
                    break
                losses.append(self._loss(step, dataset, val_index[start:start + batch_size]))
        self.model.train()
        loss = torch.stack(losses).mean().item() if losses else float("nan")
        return -10. * np.log10(loss)

    def fit(self, train_dataset, val_dataset, batch_size: int, learning_rate: float,
            num_steps: int, crop_steps: int, report_interval: int, decay_rate: float,
            decay_steps: int, weight_decay: float, visualizers: List, disable_aml=False,
            grad_sync=None) -> List[LogEntry]:
        """Adam + value/norm clipping + exponential LR decay, as ray_caster.py:248-376.
        ``grad_sync`` (optional) is called between backward and clipping -- the data-parallel
        gradient all-reduce hook (``parallel.allreduce_gradients``)."""
        from .ray_dataset_modes import Mode
        trainval = train_dataset.sample_cameras(val_dataset.num_cameras, val_dataset.num_samples, False)
        on_cuda = next(self.model.parameters()).is_cuda
        # same update rule as ray_caster.py:283,327-329; on a GPU both clips and Adam are two launches (optim.ClipAdam)
        fused_step = on_cuda and self.train_kernels
        trainer = None
        device = next(self.model.parameters()).device
        if on_cuda:       # ground truth, index tables and ray tables live in HBM for the whole run
            for ds in (train_dataset, val_dataset, trainval):
                if hasattr(ds, "to"):
                    ds.to(device)
        if fused_step:
            from . import trainer as _trainer
            from .optim import ClipAdam
            if (self.fused_trainer and _trainer.supported(self.model) and hasattr(train_dataset, "loss_tables")
                    and getattr(train_dataset, "fused_loss", False)):
                # the whole step as two C calls (trainer.FusedTrainer); it doubles as the "optimizer" of the loop
                optim = trainer = _trainer.FusedTrainer(self.model, learning_rate, weight_decay=weight_decay,
                                                        clip_value=0.1, max_norm=0.1)
            else:
                optim = ClipAdam(self.model.parameters(), learning_rate, weight_decay=weight_decay, clip_value=0.1,
                                 max_norm=0.1)
        else:
            optim = torch.optim.Adam(self.model.parameters(), learning_rate, weight_decay=weight_decay)
        step, epoch, log = 0, 0, []
        start_time = time.time()
        dataset_mode = train_dataset.mode
        for ds in (train_dataset, val_dataset, trainval):
            ds.mode = Mode.Center if crop_steps else dataset_mode

        def render_image(samples, include_depth):
            return self.batched_render(samples, batch_size, include_depth)

        def render_act(sampler, camera):
            raise NotImplementedError("activation rendering is lecture visualisation (out of scope)")

        while step <= num_steps:
            num_rays = len(train_dataset)
            index = np.arange(num_rays)
            np.random.shuffle(index)
            for start in range(0, num_rays, batch_size):
                if step > num_steps:
                    break
                exponential_lr_decay(optim, learning_rate, step, decay_rate, decay_steps)
                batch = index[start:min(start + batch_size, num_rays)]       # numpy slice (the reference makes a list)
                if trainer is not None:
                    rays = train_dataset.get_rays(batch, step).to(device)
                    if len(rays.rays) > 0:      # (an all-background batch is a NaN loss in the reference)
                        colors, alphas, weight = train_dataset.loss_tables()
                        trainer.backward(rays, colors, alphas, weight, self._lin(train_dataset.num_samples, device))
                        if grad_sync is not None:
                            grad_sync(self.model)
                        trainer.update()
                else:
                    optim.zero_grad()
                    loss = self._loss(step, train_dataset, batch)
                    loss.backward()
                    if grad_sync is not None:
                        grad_sync(self.model)
                    if not fused_step:
                        torch.nn.utils.clip_grad_value_(self.model.parameters(), 0.1)
                        torch.nn.utils.clip_grad_norm_(self.model.parameters(), 0.1)
                    optim.step()

                if step < 10 or step % report_interval == 0:
                    epoch += 1
                    train_psnr = self._validate(trainval, batch_size, step)
                    val_psnr = self._validate(val_dataset, batch_size, step)
                    # the reference asserts "no NaN colour / opacity" on every render (ray_caster.py:73-74); the
                    # kernels raise a device flag instead, read here where the loop synchronises anyway
                    self.check_nan()
                    now = time.time()
                    if step >= report_interval:
                        time_per_step = (now - start_time) / step
                        eta = time.strftime("%a, %d %b %Y %H:%M:%S +0000",
                                            time.gmtime(now + (num_steps - step) * time_per_step))
                    else:
                        time_per_step, eta = 0, "N/A"
                    print("{:07}".format(step), "{:2f} s/step".format(time_per_step),
                          "psnr_train: {:2f}".format(train_psnr), "val_psnr: {:2f}".format(val_psnr),
                          "lr: {:.2e}".format(optim.param_groups[0]["lr"]), "eta:", eta)
                    if step % report_interval == 0:
                        log.append(LogEntry(step, now - start_time,
                                            copy.deepcopy(self.model.state_dict()), train_psnr, val_psnr))
                    if train_dataset.mode == Mode.Center and step >= crop_steps:
                        print("Removing center crop...")
                        for ds in (train_dataset, val_dataset, trainval):
                            ds.mode = dataset_mode
                        step += 1
                        break
                for visualizer in visualizers:
                    visualizer.visualize(step, render_image, render_act)
                step += 1
        return log


    def to_scenepic(self, dataset, *args, **kwargs):
        """The reference writes an interactive scenepic HTML here (ray_caster.py:379-488): lecture
        visualisation, out of scope.  Returns an object whose ``save_as_html`` writes a stub page."""
        class _Stub:
            def save_as_html(self, path, *a, **k):
                with open(path, "w") as f:
                    f.write("<html><body>scenepic visualisation is not produced by the B200 build</body></html>")
        return _Stub()
