#!/usr/bin/env python
"""rays/sec of the fused render path (lego_400 shape, NeRF 8x256, 64 samples/ray).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is one pass of the hot path (stratified sampling -> encoding -> MLP ->
compositing) over one batch of ``--rays`` synthetic lego_400-shaped rays per GPU.
``value`` times the pass with the rays already in HBM (CUDA events, max over ranks);
``e2e`` is the same pass through the public API (``Raycaster.render`` on a host
``RayBundle``): pinned host rays -> device -> render -> host pixels.
For N > 1 launch with torchrun (one rank per GPU); rays shard by index, no data-path
collective ("scaling": "weak": per-GPU batch fixed).

``--impl reference`` times the reference's algorithm on the host CPU cores: the numpy
oracle port (the reference is pure Python/PyTorch and cannot travel to the GPU box).
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_SAMPLE = 1186816          # 593,408 MAC per sample
SAMPLES = 64
METRIC = "rays/sec (lego_400, 64 samples/ray)"
WORKLOAD = ("lego_400-shaped synthetic rays (400x400 look-at cameras, radius 4, fov 40deg, AABB [-1,1]^3, "
            "valid rays only), NeRF(8,256,9,10,3,4,[4],True) random init seed 20080524, "
            "stratified 64 samples/ray, render-forward incl. depth")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rays", type=int, default=1 << 20, help="rays per step per GPU")
    ap.add_argument("--cameras", type=int, default=0, help="cameras in the ray pool (0 = enough for all steps)")
    ap.add_argument("--operand", default="fp16", choices=["fp16", "bf16"])
    ap.add_argument("--cpu-rays", type=int, default=8192, help="rays of the parity sample (numpy oracle)")
    ap.add_argument("--ref-rays", type=int, default=32768, help="rays per step of the timed CPU reference sample")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline / parity legs")
    return ap.parse_args()


def lego_camera_matrices(num, res=400):
    """(K (3,3), [E (4,4)]): look-at cameras on the upper hemisphere, radius 4, fov_y 40 deg (orbit_video.py:21-24);
    camera-to-world, +z forward / +y down.  numpy only: shared by both arms."""
    focal = .5 * res / np.tan(.5 * 40 * np.pi / 180)
    K = np.array([[focal, 0, res / 2], [0, focal, res / 2], [0, 0, 1]], np.float32)
    golden = np.pi * (3 - np.sqrt(5))
    exts = []
    for i in range(num):
        z = 0.15 + 0.8 * (i + 0.5) / num
        r = np.sqrt(1 - z * z)
        pos = 4.0 * np.array([r * np.cos(golden * i), z, r * np.sin(golden * i)])
        fwd = -pos / np.linalg.norm(pos)
        right = np.cross(fwd, np.array([0, 1.0, 0]))
        right /= np.linalg.norm(right)
        ext = np.eye(4)
        ext[:3, 0], ext[:3, 1], ext[:3, 2], ext[:3, 3] = right, np.cross(fwd, right), fwd, pos
        exts.append(ext.astype(np.float32))
    return K, exts


def lego_cameras(num, res=400):
    import fourier_feature_nets_b200 as ffn
    K, exts = lego_camera_matrices(num, res)
    return [ffn.CameraInfo.create("cam%d" % i, ffn.Resolution(res, res), K, e) for i, e in enumerate(exts)]


def model_params_numpy(model):
    return {k: v.detach().cpu().numpy() for k, v in model.state_dict().items()}


def oracle_render(params, o, d, near, far, u, chunk=8192):
    import oracle
    samples = oracle.sample_rays(o, d, near, far, SAMPLES, u=u)
    return oracle.render_rays(lambda p, v: oracle.nerf_forward(params, p, v), samples, True, True, chunk)


def torch_oracle_render(params, o, d, near, far, u):
    """The reference's own ATen op sequence on the host cores (oracle/ffn_oracle_torch.py); params: name -> CPU tensor."""
    import torch
    from oracle import ffn_oracle_torch as ot
    t = [torch.from_numpy(np.ascontiguousarray(a)) for a in (o, d, near, far, u)]
    return ot.render_rays(params, t[0], t[1], t[2], t[3], SAMPLES, t[4], True, batch=4096)


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML during the timed region."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()

    def run(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            names = {
                pynvml.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
                pynvml.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                pynvml.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                pynvml.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                pynvml.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake",
            }
            while not self._stop_evt.is_set():
                self.samples.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                r = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
                self._stop_evt.wait(0.05)
        except Exception as e:  # NVML missing: report it, do not fake numbers
            self.reasons.add("nvml_unavailable:%s" % type(e).__name__)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return float(p["bf16_tflops"]), "MEASURED_PEAKS.json bf16_tflops (burst, cuBLAS bf16; fp16 runs at the same rate)"
    return 1590.0, "fallback 1.59 PFLOP/s (B200_PROFILING.md)"


def sustained_peak():
    """cuBLAS bf16 TFLOP/s inside a long step (the bench step is one ~65 ms launch under the power cap)."""
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["bf16_tflops_sustained"])
    except Exception:
        return None


def ncu_traffic():
    """dram bytes per launch of the render kernel from the committed ncu capture, if any."""
    path = os.path.join(ROOT, "profiles", "ncu_render_kernel.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f)
    return None


def run_reference(args):
    """The reference's CPU path on the host cores, same config: only ``oracle/`` (+ numpy / torch) runs here, nothing of
    the product package."""
    import torch
    import oracle
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    params = {k: torch.from_numpy(v) for k, v in oracle.init_nerf_params(seed=20080524).items()}
    K_, exts = lego_camera_matrices(1)
    xs, ys = np.meshgrid(np.arange(400), np.arange(400))
    starts, directions = oracle.raycast(K_, exts[0], np.stack([xs, ys], -1).reshape(-1, 2))
    nf, ok = oracle.near_far(np.diag([2, 2, 2, 1]).astype(np.float32), starts, directions)
    valid = np.nonzero(ok)[0]
    n = min(args.ref_rays, len(valid))
    rng = np.random.default_rng(0)
    times = []
    for step in range(args.warmup + args.steps):
        idx = rng.choice(valid, n, replace=False)
        o, d = starts[idx].astype(np.float32), directions[idx].astype(np.float32)
        near, far = nf[0, idx].astype(np.float32), nf[1, idx].astype(np.float32)
        u = rng.random((n, SAMPLES), dtype=np.float32)
        t0 = time.perf_counter()
        torch_oracle_render(params, o, d, near, far, u)
        dt = time.perf_counter() - t0
        if step >= args.warmup:
            times.append(dt)
    total = sum(times)
    value = n * len(times) / total
    cores = torch.get_num_threads()
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "rays/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "rays_per_step": n, "samples_per_ray": SAMPLES},
        "cpu_baseline": {"value": value, "unit": "rays/s", "cores": cores, "kind": "port",
                         "sample": "%d rays x %d samples per step in batches of 4096, the reference's ATen op sequence "
                                   "on the host (oracle/ffn_oracle_torch.py), torch threads: %d of %d cores"
                                   % (n, SAMPLES, cores, os.cpu_count())},
        "e2e": {"value": value, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist
    import fourier_feature_nets_b200 as ffn
    from fourier_feature_nets_b200 import _lib, engine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU: the render path has no CPU fallback")
    _lib.lib()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    R, K, W = args.rays, args.steps, args.warmup
    torch.manual_seed(20080524)          # train_nerf.py:48
    model = ffn.NeRF(8, 256, 9, 10, 3, 4, [4], True).to(dev).eval()
    model.ffn_operand = args.operand
    rc = ffn.Raycaster(model)

    # ---- ray pool: every rank owns its own shard of cameras (rays shard by index) -------
    need = R * (K + W)
    ncam = args.cameras or int(np.ceil(need / (0.70 * 160000))) + 1
    cams = lego_cameras(ncam * world)[rank::world]
    sampler = ffn.RaySampler(np.diag([2, 2, 2, 1]).astype(np.float32), cams, SAMPLES, stratified=True)
    sampler.device_jitter = True   # in-kernel Philox jitter (perf mode); parity legs feed explicit jitter
    valid = torch.nonzero(sampler.valid_mask).flatten()
    perm = valid[torch.randperm(len(valid), generator=torch.Generator().manual_seed(rank))]
    pool_rays = len(perm)
    step_idx = [perm[(i * R) % max(1, pool_rays - R):][:R] for i in range(K + W)]
    assert all(len(ix) == R for ix in step_idx), "ray pool too small; raise --cameras"

    # HBM-resident inputs (value) and pinned host inputs (e2e)
    host_bundles = [sampler.sample(ix, None) for ix in step_idx]
    dev_bundles = [b.to(dev) for b in host_bundles]
    pinned = [b.pin_memory() for b in host_bundles]
    pool_bytes = sum(t.numel() * 4 for b in dev_bundles for t in (b.starts, b.directions, b.near, b.far))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing ------------------------------------------------------
    with torch.no_grad():
        for i in range(W):
            rc.render(dev_bundles[i], True)
        barrier()
        clocks = ClockSampler(local_rank)
        clocks.start()
        launches0 = _lib.launch_count()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
        for i in range(K):
            evs[i][0].record()
            out = rc.render(dev_bundles[W + i], True)
            evs[i][1].record()
        barrier()
        launches = _lib.launch_count() - launches0
        clock_info = clocks.stop()
    step_ms = [a.elapsed_time(b) for a, b in evs]
    total_ms = evs[0][0].elapsed_time(evs[-1][1])
    rc.check_nan()

    # ---- end to end through the public API with host buffers ----------------------------
    with torch.no_grad():
        for i in range(min(W, 2)):
            rc.render(pinned[i].to(dev, non_blocking=True), True).numpy()
        barrier()
        t0 = time.perf_counter()
        for i in range(K):
            res = rc.render(pinned[W + i].to(dev, non_blocking=True), True).numpy()
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
    h2d = R * 32
    d2h = R * 20

    t = torch.tensor([total_ms, e2e_s * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms = t.tolist()

    if rank == 0:
        value = world * R * K / (total_ms / 1e3)
        e2e_value = world * R * K / (e2e_ms / 1e3)
        peak, peak_src = measured_peak()
        kernel_ms = float(np.mean(step_ms))
        achieved = R * SAMPLES * FLOP_PER_SAMPLE / (kernel_ms / 1e3) / 1e12
        traffic = ncu_traffic()
        line = {
            "metric": METRIC, "value": value, "unit": "rays/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": args.operand + " operands, f32 accumulate/encode/composite",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "rays_per_step_per_gpu": R, "samples_per_ray": SAMPLES,
                       "parallelism": "rays sharded by index across %d rank(s), no data-path collective" % world,
                       "l2": "inputs larger than L2: %.0f MB ray pool per GPU, every step reads a different slice"
                             % (pool_bytes / 1e6),
                       "jitter": "in-kernel Philox4x32-10"},
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                         "frac": achieved / peak, "peak_source": peak_src + " (of measured)",
                         "frac_of_sustained": None if sustained_peak() is None else achieved / sustained_peak(),
                         # ncu (one --set full capture) measured dram bytes for a launch of
                         # `rays_per_launch` rays; traffic scales linearly with the rays of a launch
                         "traffic": None if traffic is None else
                         traffic["dram_bytes_per_launch"] * R / traffic.get("rays_per_launch", R),
                         "kernel": "ffn_render_kernel<fp16>", "kernel_ms": kernel_ms,
                         "flop_per_launch": R * SAMPLES * FLOP_PER_SAMPLE,
                         "hbm_gbs_algorithmic": R * 52 / (kernel_ms / 1e3) / 1e9},
            "e2e": {"value": e2e_value, "unit": "rays/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "api": "Raycaster.render(RayBundle.pin_memory().to(device), include_depth=True).numpy()"},
            "gpu_launches": int(launches),
            "clocks": clock_info,
        }
        if not args.no_cpu and world == 1:
            # parity + CPU baseline on a bounded sample of the same workload (rank 0, N=1 only)
            n = args.cpu_rays
            b = host_bundles[W]
            o, d = b.starts[:n].numpy(), b.directions[:n].numpy()
            near, far = b.near[:n].numpy(), b.far[:n].numpy()
            u = np.random.default_rng(1).random((n, SAMPLES), dtype=np.float32)
            params = model_params_numpy(model)
            ref = oracle_render(params, o, d, near, far, u)          # parity arbiter: the numpy oracle
            # timed CPU baseline: the reference's ATen op sequence on the host cores, a bounded sample
            nb_cpu = min(args.ref_rays, len(host_bundles[W].starts))
            hb = host_bundles[W]
            ucpu = np.random.default_rng(2).random((nb_cpu, SAMPLES), dtype=np.float32)
            cpu_model = {k: v.detach().cpu() for k, v in model.state_dict().items()}
            cargs = (hb.starts[:nb_cpu].numpy(), hb.directions[:nb_cpu].numpy(), hb.near[:nb_cpu].numpy(),
                     hb.far[:nb_cpu].numpy(), ucpu)
            torch_oracle_render(cpu_model, *[a[:4096] for a in cargs])     # warm the thread pool
            t0 = time.perf_counter()
            torch_oracle_render(cpu_model, *cargs)
            cpu_s = time.perf_counter() - t0
            with torch.no_grad():
                bb = ffn.RayBundle(b.starts[:n], b.directions[:n], b.near[:n], b.far[:n], b.rays[:n], SAMPLES,
                                   True, torch.from_numpy(u))
                ours = rc.render(bb.to(dev), True).numpy()
            err = np.abs(ours.color - ref.color)
            mse = float(np.mean((ours.color - ref.color) ** 2))
            cores = torch.get_num_threads()
            line["cpu_baseline"] = {
                "value": nb_cpu / cpu_s, "unit": "rays/s", "cores": cores, "kind": "port",
                "sample": "%d rays x %d samples of the same workload in batches of 4096, the reference's ATen op "
                          "sequence on the host (oracle/ffn_oracle_torch.py), torch threads: %d of %d cores"
                          % (nb_cpu, SAMPLES, cores, os.cpu_count())}
            # the reference's op sequence (RaySamples -> NeRF.forward -> compositing, ray_caster.py:48-93) as plain
            # fp32 PyTorch on the SAME GPU, inference batches of 4096 rays like orbit_video.py:37 -- the "1-GPU
            # PyTorch" denominator of north_star's >= 10x target.  Bounded sample, device-resident inputs.
            with torch.no_grad():
                tb = 4096
                nb = 16
                mats = [dev_bundles[W].subset(range(i * tb, (i + 1) * tb)).materialize() for i in range(nb)]
                model.forward = model.forward_torch      # plain PyTorch layers instead of the CUDA engine
                try:
                    launches_t = _lib.launch_count()
                    ref_t = rc._render_torch(mats[0], True)
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    for m in mats:
                        rc._render_torch(m, True)
                    torch.cuda.synchronize()
                    torch_s = time.perf_counter() - t0
                    assert _lib.launch_count() == launches_t, "torch baseline must not touch libffn_b200"
                finally:
                    del model.forward
                ours_t = rc.render(dev_bundles[W].subset(range(0, tb)).materialize(), True)
            line["torch_gpu_baseline"] = {
                "value": nb * tb / torch_s, "unit": "rays/s",
                "sample": "%d batches of %d rays x %d samples, plain fp32 PyTorch ops of the reference definition on "
                          "the same GPU, samples already materialised in HBM" % (nb, tb, SAMPLES),
                "color_max_abs_vs_ours": float((ours_t.color - ref_t.color).abs().max())}
            # the training step of the same model (row a-14 of SURVEY.md section 8): FusedTrainer.backward + update at
            # train_nerf.py's batch (1024 rays) x 128 samples; explanatory, not the headline metric
            try:
                tm = ffn.NeRF(8, 256, 9, 10, 3, 4, [4], True).to(dev)
                tr = ffn.FusedTrainer(tm, 5e-4)
                tb_r, tb_s = 1024, 128
                tbun = dev_bundles[W].subset(range(0, tb_r))
                tbun = ffn.RayBundle(tbun.starts, tbun.directions, tbun.near, tbun.far,
                                     torch.arange(tb_r, device=dev), tb_s, True, None, seed=1)
                gt_c, gt_a = torch.rand((tb_r, 3), device=dev), torch.rand((tb_r,), device=dev)
                lin = torch.linspace(0, 1, tb_s).to(dev)
                l0 = _lib.launch_count()
                for _ in range(5):
                    tr.backward(tbun, gt_c, gt_a, 0.1, lin)
                    tr.update()
                per_step = (_lib.launch_count() - l0) // 5
                torch.cuda.synchronize()
                ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                ev0.record()
                for _ in range(20):
                    tr.backward(tbun, gt_c, gt_a, 0.1, lin)
                    tr.update()
                ev1.record()
                torch.cuda.synchronize()
                ms = ev0.elapsed_time(ev1) / 20
                line["train_step"] = {"ms_per_step": ms, "rays_per_s": tb_r / ms * 1e3, "rays": tb_r, "samples": tb_s,
                                      "kernel_launches_per_step": int(per_step),
                                      "what": "forward-with-saves + loss + dgrad + wgrad + clip + Adam + re-pack "
                                              "(FusedTrainer, two C calls)"}
            except Exception as e:       # never lose the headline line over the explanatory leg
                line["train_step"] = {"error": "%s: %s" % (type(e).__name__, e)}
            line["parity"] = {"rays": n, "color_max_abs": float(err.max()),
                              "alpha_max_abs": float(np.abs(ours.alpha - ref.alpha).max()),
                              "depth_mismatch_frac": float((ours.depth != ref.depth).mean()),
                              "psnr_vs_ref_db": float(-10 * np.log10(max(mse, 1e-20)))}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
