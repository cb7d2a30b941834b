#!/usr/bin/env python
"""rays/sec of the fused render path (lego_400 shape, NeRF 8x256, 64 samples/ray).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is one pass of the hot path (stratified sampling -> encoding -> MLP -> compositing) over one batch of
``--rays`` synthetic lego_400-shaped rays per GPU.  ``value`` times the pass with the rays already in HBM (CUDA
events, max over ranks); ``e2e`` is the same pass through the public API with HOST buffers, everything a user pays
per step inside the timed region: ``RaySampler.sample`` on host-resident ray tables (gather into pinned staging),
host -> device copy, ``Raycaster.render``, device -> host read of the pixels.  For N > 1 launch with torchrun (one
rank per GPU); rays shard by index, no data-path collective ("scaling": "weak": per-GPU batch fixed).

Legs beside the headline (explanatory, same JSON line): ``train`` (FusedTrainer step, with the NCCL gradient
all-reduce at N > 1, weak and strong scaling), ``frame`` (one 800x800 orbit_video.py frame ray-split over the ranks),
``spp`` (128 / 192 samples per ray), and at N = 1 ``parity``, ``cpu_baseline`` and ``torch_gpu_baseline``.

``--impl reference`` times the REAL reference (``oracle/_ref``: the unmodified ``fourier_feature_nets`` package,
vendored by ``__graft_entry__.build()``) on the host cores through its own public API --
``Raycaster.batched_render(RaySampler.sample(idx, None), 4096, True)`` (ray_caster.py:103-138, ray_sampler.py:359-403)
-- with every host thread, on this arm's config; each step is a bounded sample of the step's rays.  Where
``oracle/_ref`` is absent it falls back to the oracle port of the same op sequence (``kind: "port"``).
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

FLOP_PER_SAMPLE = 1186816          # 593,408 MAC per sample = the reference's algorithm
# what the inference kernel EXECUTES per sample: bottleneck (no activation, nerf_model.py:119) is folded into hidden_view
# when the weights are packed, i.e. one 256x256 layer (65,536 MAC) less
FLOP_PER_SAMPLE_EXECUTED = FLOP_PER_SAMPLE - 2 * 65536
SAMPLES = 64
METRIC = "rays/sec (lego_400, 64 samples/ray)"
WORKLOAD = ("lego_400-shaped synthetic rays (400x400 look-at cameras, radius 4, fov 40deg, AABB [-1,1]^3, "
            "valid rays only), NeRF(8,256,9,10,3,4,[4],True) random init seed 20080524, "
            "stratified 64 samples/ray, render-forward incl. depth")
BOUNDS = np.diag([2, 2, 2, 1]).astype(np.float32)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rays", type=int, default=1 << 20, help="rays per step per GPU")
    ap.add_argument("--operand", default="fp16", choices=["fp16", "bf16"])
    ap.add_argument("--cpu-rays", type=int, default=8192, help="rays of the parity sample")
    ap.add_argument("--ref-rays", type=int, default=16384, help="rays per step of the timed CPU reference sample")
    ap.add_argument("--legs", default="all", help="comma list of extra legs: train,frame,spp,precise,cpu,torchgpu,parity | all | none")
    return ap.parse_args()


def bench_config(args, world):
    """The workload description: identical in both arms (the reference arm runs a bounded sample of it, stated in its
    cpu_baseline.sample)."""
    return {"workload": WORKLOAD, "rays_per_step_per_gpu": args.rays, "samples_per_ray": SAMPLES,
            "parallelism": "rays sharded by index across %d rank(s), no data-path collective" % world,
            "l2": "inputs larger than L2: %.0f MB ray pool per GPU, every step reads a different slice"
                  % ((args.steps + args.warmup) * args.rays * 32 / 1e6),
            "jitter": "stratified, one uniform draw per sample (ours: in-kernel Philox4x32-10; reference: torch.rand)"}


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


def make_cameras(pkg, num, res=400):
    """``CameraInfo`` objects of package ``pkg`` (ours or the reference: same constructor, camera_info.py:39-52)."""
    K, exts = lego_camera_matrices(num, res)
    return [pkg.CameraInfo.create("cam%d" % i, pkg.Resolution(res, res), K, e) for i, e in enumerate(exts)]


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


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


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"tensor": float(p["bf16_tflops"]), "tensor_sustained": float(p.get("bf16_tflops_sustained", 0)) or None,
                "hbm": float(p["hbm_gbs"]), "src": "MEASURED_PEAKS.json (cuBLAS bf16 burst; fp16 runs at the same rate)"}
    return {"tensor": 1590.0, "tensor_sustained": None, "hbm": 6550.0, "src": "fallback of B200_PROFILING.md"}


def ncu_traffic():
    """dram bytes per launch of the render kernel from the committed ncu capture, if any."""
    path = os.path.join(ROOT, "profiles", "ncu_render_kernel.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f)
    return None


# =====================================================================================================
# reference arm: the reference's own CPU implementation on the host cores
# =====================================================================================================
class ReferenceCPU:
    """The real reference (oracle/_ref) driven through its public API; falls back to the oracle port."""

    def __init__(self, num_cameras=2, device="cpu", state_dict=None):
        import torch
        from oracle import reference as refmod
        self.torch = torch
        self.device = device
        self.kind = "reference" if refmod.available() else "port"
        K_, exts = lego_camera_matrices(num_cameras)
        if self.kind == "reference":
            ref = refmod.import_reference()
            self.where = os.path.relpath(refmod.location(), ROOT) if refmod.location().startswith(ROOT) else refmod.location()
            cams = make_cameras(ref, num_cameras)
            self.sampler = ref.RaySampler(BOUNDS, cams, SAMPLES, True)            # stratified, no opacity model
            torch.manual_seed(20080524)                                           # train_nerf.py:48
            model = ref.NeRF(8, 256, 9, 10, 3, 4, [4], True)
            if state_dict is not None:
                model.load_state_dict(state_dict)
            self.model = model.to(device).eval()
            self.rc = ref.Raycaster(self.model)
            invalid = self.sampler.invalid_rays
            self.valid = np.array([i for i in range(len(self.sampler)) if i not in invalid], np.int64)
        else:
            import oracle
            self.where = "oracle/ffn_oracle_torch.py"
            self.params = {k: torch.from_numpy(v) for k, v in oracle.init_nerf_params(seed=20080524).items()}
            if state_dict is not None:
                self.params = {k: v.detach().cpu() for k, v in state_dict.items()}
            xs, ys = np.meshgrid(np.arange(400), np.arange(400))
            pts = np.stack([xs, ys], -1).reshape(-1, 2)
            o, d, nf, ok = [], [], [], []
            for e in exts:
                oo, dd = oracle.raycast(K_, e, pts)
                n_f, k = oracle.near_far(BOUNDS, oo, dd)
                o.append(oo), d.append(dd), nf.append(n_f), ok.append(k)
            self.o, self.d = np.concatenate(o).astype(np.float32), np.concatenate(d).astype(np.float32)
            self.nf = np.concatenate(nf, -1).astype(np.float32)
            self.valid = np.nonzero(np.concatenate(ok))[0]

    def describe(self, n, cores):
        if self.kind == "reference":
            return ("%d rays x %d samples per step: the unmodified reference (%s) on %s, Raycaster.batched_render("
                    "RaySampler.sample(idx, None), 4096, True), torch threads %d of %d cores"
                    % (n, SAMPLES, self.where, self.device, cores, os.cpu_count()))
        return ("%d rays x %d samples per step in batches of 4096: oracle port of the reference's ATen op sequence "
                "(%s; oracle/_ref not vendored), torch threads %d of %d cores"
                % (n, SAMPLES, self.where, cores, os.cpu_count()))

    def step(self, idx):
        """sample + render ``idx`` (numpy int64 ray indices); returns the render result (numpy fields)."""
        torch = self.torch
        if self.kind == "reference":
            samples = self.sampler.sample(torch.from_numpy(idx), None)
            return self.rc.batched_render(samples, 4096, True)
        from oracle import ffn_oracle_torch as ot
        n = len(idx)
        u = torch.rand((n, SAMPLES), dtype=torch.float32)
        t = [torch.from_numpy(np.ascontiguousarray(a)) for a in (self.o[idx], self.d[idx], self.nf[0, idx], self.nf[1, idx])]
        return ot.render_rays(self.params, t[0], t[1], t[2], t[3], SAMPLES, u, True, batch=4096)


def run_reference(args):
    """Rank 0 alone runs and prints; only ``oracle/`` (+ numpy / torch) runs here, nothing of the product package."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    import contextlib
    import torch
    cores = host_cores()
    torch.set_num_threads(cores)          # torchrun exports OMP_NUM_THREADS=1: the reference gets every host core
    times = []
    with contextlib.redirect_stdout(sys.stderr):      # the reference prints while it builds its tables
        arm = ReferenceCPU()
        n = min(args.ref_rays, len(arm.valid))
        rng = np.random.default_rng(0)
        for step in range(args.warmup + args.steps):
            idx = np.sort(rng.choice(arm.valid, n, replace=False))
            t0 = time.perf_counter()
            arm.step(idx)
            dt = time.perf_counter() - t0
            if step >= args.warmup:
                times.append(dt)
    total = sum(times)
    value = n * len(times) / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "rays/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": bench_config(args, args.gpus),
        "cpu_baseline": {"value": value, "unit": "rays/s", "cores": torch.get_num_threads(), "kind": arm.kind,
                         "sample": arm.describe(n, torch.get_num_threads())},
        "e2e": {"value": value, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# =====================================================================================================
# our arm
# =====================================================================================================
def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist
    import fourier_feature_nets_b200 as ffn
    from fourier_feature_nets_b200 import _lib, parallel

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
        import datetime
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=300))
    legs = set("train,frame,spp,precise,cpu,torchgpu,parity".split(",")) if args.legs == "all" else \
        set(x for x in args.legs.split(",") if x and x != "none")

    R, K, W = args.rays, args.steps, args.warmup
    torch.manual_seed(20080524)          # train_nerf.py:48
    model = ffn.NeRF(8, 256, 9, 10, 3, 4, [4], True).to(dev).eval()
    model.ffn_operand = args.operand
    rc = ffn.Raycaster(model)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(*vals):
        t = torch.tensor(vals, dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.tolist()

    # ---- ray pool: every rank owns its own shard of cameras (rays shard by index); the tables are generated on the
    # GPU (ffn_generate_rays) and a host copy feeds the e2e leg like the reference's host-resident tables ----------
    need = R * (K + W)
    ncam = int(np.ceil(need / (0.70 * 160000))) + 1
    cams = make_cameras(ffn, ncam * world)[rank::world]
    sampler = ffn.RaySampler(BOUNDS, cams, SAMPLES, stratified=True, device=dev)
    sampler.device_jitter = True   # in-kernel Philox jitter (perf mode); parity legs feed explicit jitter
    valid = torch.nonzero(sampler.valid_mask).flatten()
    perm = valid[torch.randperm(len(valid), generator=torch.Generator().manual_seed(rank)).to(dev)]
    pool_rays = len(perm)
    step_idx = [perm[(i * R) % max(1, pool_rays - R):][:R] for i in range(K + W)]
    assert all(len(ix) == R for ix in step_idx), "ray pool too small"
    dev_bundles = [sampler.sample(ix, None) for ix in step_idx]       # HBM-resident inputs of `value`

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
            rc.render(dev_bundles[W + i], True)
            evs[i][1].record()
        barrier()
        launches = _lib.launch_count() - launches0
        clock_info = clocks.stop()
    step_ms = [a.elapsed_time(b) for a, b in evs]
    total_ms = evs[0][0].elapsed_time(evs[-1][1])
    rc.check_nan()

    # ---- end to end through the public API with HOST buffers: host ray tables -> RaySampler.sample (gather into
    # pinned staging) -> H2D -> Raycaster.render -> D2H; everything inside the timed region ---------------------------
    # torchrun exports OMP_NUM_THREADS=1: the host-side gather of RaySampler.sample (torch.index_select over the host ray
    # tables) would run on ONE core per rank and bound the e2e number at N > 1; give every rank its share of the cores
    host_threads = max(1, host_cores() // world)
    torch.set_num_threads(host_threads)
    host_sampler = ffn.RaySampler.__new__(ffn.RaySampler)
    host_sampler.__dict__.update(sampler.__dict__)
    host_sampler.to("cpu")
    host_sampler.device_jitter = True
    host_sampler.enable_pinned_staging(R)
    host_idx = [ix.cpu() for ix in step_idx]
    with torch.no_grad():
        for _ in rc.render_stream(host_sampler, host_idx[:min(W, 2)], True):
            pass
        barrier()
        t0 = time.perf_counter()
        n_out = 0
        for res in rc.render_stream(host_sampler, host_idx[W:W + K], True):     # numpy pixels of every step, in order
            n_out += len(res.alpha)
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
        assert n_out == R * K
        # the same step by step, nothing overlapped: sample -> H2D -> render -> D2H, then the next step
        barrier()
        t0 = time.perf_counter()
        for i in range(K):
            res = rc.render(host_sampler.sample(host_idx[W + i], None).to(dev, non_blocking=True), True).numpy()
        torch.cuda.synchronize()
        e2e_serial_s = time.perf_counter() - t0
        # (and with device-resident ray tables, SURVEY 8f-1: only the ray indices cross the bus)
        for i in range(min(W, 2)):
            rc.render(sampler.sample(host_idx[i], None), True).numpy()
        barrier()
        t0 = time.perf_counter()
        for i in range(K):
            rc.render(sampler.sample(host_idx[W + i], None), True).numpy()
        torch.cuda.synchronize()
        e2e_dev_s = time.perf_counter() - t0
    del host_sampler
    total_ms, e2e_ms, e2e_dev_ms, e2e_serial_ms = max_over_ranks(total_ms, e2e_s * 1e3, e2e_dev_s * 1e3, e2e_serial_s * 1e3)

    pk = peaks()
    line = None
    if rank == 0:
        value = world * R * K / (total_ms / 1e3)
        kernel_ms = float(np.mean(step_ms))
        achieved = R * SAMPLES * FLOP_PER_SAMPLE / (kernel_ms / 1e3) / 1e12
        traffic = ncu_traffic()
        line = {
            "metric": METRIC, "value": value, "unit": "rays/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": args.operand + " operands, f32 accumulate/encode/composite",
            "data": "synthetic", "config": bench_config(args, world),
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": pk["tensor"], "unit": "TFLOP/s",
                         "frac": achieved / pk["tensor"], "peak_source": pk["src"],
                         "frac_of_sustained": achieved / pk["tensor_sustained"] if pk["tensor_sustained"] else None,
                         # `achieved` counts the reference's algorithmic FLOP; the kernel executes 11 % fewer (folded
                         # bottleneck layer): the tensor pipe itself runs at `executed_frac` of the burst peak
                         "executed_flop_per_launch": R * SAMPLES * FLOP_PER_SAMPLE_EXECUTED,
                         "executed_frac": achieved * FLOP_PER_SAMPLE_EXECUTED / FLOP_PER_SAMPLE / pk["tensor"],
                         # ncu (one --set full capture): dram bytes of a launch of `rays_per_launch` rays, scaled
                         "traffic": None if traffic is None else
                         traffic["dram_bytes_per_launch"] * R / traffic.get("rays_per_launch", R),
                         "kernel": "ffn_infer_kernel<%s>" % args.operand, "kernel_ms": kernel_ms,
                         "flop_per_launch": R * SAMPLES * FLOP_PER_SAMPLE,
                         "hbm_gbs_algorithmic": R * 52 / (kernel_ms / 1e3) / 1e9},
            "e2e": {"value": world * R * K / (e2e_ms / 1e3), "unit": "rays/s", "h2d_bytes_per_step": R * 40,
                    "d2h_bytes_per_step": R * 20,
                    "api": "Raycaster.render_stream(sampler, index_batches, True): host ray tables -> RaySampler.sample "
                           "(gather into pinned staging) -> H2D -> fused kernel -> D2H -> numpy, every stage inside the "
                           "timed region, host work of step i+1 overlapped with the kernel of step i",
                    "host_threads_per_rank": host_threads,
                    "serial_value": world * R * K / (e2e_serial_ms / 1e3),
                    "device_tables_value": world * R * K / (e2e_dev_ms / 1e3)},
            "gpu_launches": int(launches),
            "clocks": clock_info,
        }

    # ================================================================== explanatory legs (never lose the headline)
    extra = {}

    def leg(name, fn):
        import contextlib
        if name.split("_")[0] not in legs:
            return
        try:
            with contextlib.redirect_stdout(sys.stderr):
                out = fn()
            if out is not None:
                extra[name] = out
        except Exception as e:
            extra[name] = {"error": "%s: %s" % (type(e).__name__, str(e)[:200])}
        barrier()

    # ---- training step: FusedTrainer, 128 samples/ray; at N > 1 with the NCCL
    # all-reduce of the 595,844 gradients between backward and update (mean folded into ffn_clip_adam) ----------------
    def train_leg(rays_per_rank, label, steps=120, warm=12, S=128):
        torch.manual_seed(20080524)
        if label == "tiny":       # configs[1]: train_tiny_nerf.py positional preset (train_tiny_nerf.py:75-88)
            tm = ffn.PositionalFourierMLP(3, 4, 5.5).to(dev)
        else:
            tm = ffn.NeRF(8, 256, 9, 10, 3, 4, [4], True).to(dev)
        parallel.broadcast_parameters(tm)
        tr = ffn.FusedTrainer(tm, 5e-4)
        src = dev_bundles[W]
        n = rays_per_rank
        buns = []
        for j in range(4):      # a few different batches (every rank its own rays)
            sl = slice(j * n, (j + 1) * n)
            buns.append(ffn.RayBundle(src.starts[sl], src.directions[sl], src.near[sl], src.far[sl],
                                      torch.arange(n, device=dev), S, True, None, seed=1 + j + 100 * rank))
        gen = torch.Generator(device=dev).manual_seed(7 + rank)
        gt_c = torch.rand((n, 3), device=dev, generator=gen)
        gt_a = torch.rand((n,), device=dev, generator=gen)
        lin = torch.linspace(0, 1, S).to(dev)
        sync = (lambda: parallel.allreduce_gradients(tm)) if world > 1 else (lambda: None)
        l0 = _lib.launch_count()
        for i in range(warm):
            tr.backward(buns[i % 4], gt_c, gt_a, 0.1, lin)
            sync()
            tr.update()
        per_step = (_lib.launch_count() - l0) // warm
        barrier()
        ar = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for i in range(steps):
            tr.backward(buns[i % 4], gt_c, gt_a, 0.1, lin)
            ar[i][0].record()
            sync()
            ar[i][1].record()
            tr.update()
        ev1.record()
        barrier()
        ms = ev0.elapsed_time(ev1) / steps
        ar_us = 1e3 * float(np.median([a.elapsed_time(b) for a, b in ar]))
        flat = torch.cat([p.detach().reshape(-1) for p in tm.parameters()])
        spread = 0.0
        if world > 1:
            lo, hi = flat.clone(), flat.clone()
            dist.all_reduce(lo, op=dist.ReduceOp.MIN)
            dist.all_reduce(hi, op=dist.ReduceOp.MAX)
            spread = float((hi - lo).abs().max())
        ms, ar_us = max_over_ranks(ms, ar_us)
        rows = n * S
        # algorithmic HBM bytes of a step: every saved tensor written once and read once:
        # 10 x 512 B activations + 10 x 512 B dz + 9 x 32 B sign words + 256 B encodings per sample row, x 2
        hbm = 2.0 * rows * (10 * 512 + 10 * 512 + 9 * 32 + 256)
        out = {"ms": round(ms, 4), "rays_s": round(world * n / ms * 1e3), "rays_per_rank": n, "spp": S,
               "launches": int(per_step)}
        if world > 1:
            out.update({"allreduce_us": round(ar_us, 1), "spread": spread})
        if label == "train":
            out.update({"tflops": round(3 * rows * FLOP_PER_SAMPLE / (ms / 1e3) / 1e12, 1),
                        "frac_tensor": round(3 * rows * FLOP_PER_SAMPLE / (ms / 1e3) / 1e12 / pk["tensor"], 3),
                        "hbm_gbs": round(hbm / (ms / 1e3) / 1e9), "frac_hbm": round(hbm / (ms / 1e3) / 1e9 / pk["hbm"], 3)})
        return out

    leg("train", lambda: train_leg(1024, "train"))                    # weak: train_nerf.py:27 batch per rank
    if world == 1:
        leg("train_tiny", lambda: train_leg(1024, "tiny", S=64))
    if world > 1:
        leg("train_strong", lambda: train_leg(1024 // world, "strong"))   # strong: 1024 rays globally

    # ---- one 800x800 frame of orbit_video.py (128 samples: 64 uniform + 64 focused on the model's own coarse pass,
    # orbit_video.py:69-78), valid rays split over the ranks, pixels gathered on every rank ----------------------------
    def frame_leg(res=800, S=128, batch=4096, reps=3):
        cam = ffn.orbit(np.array([0, 1, 0], np.float32), np.array([0, 0, -1], np.float32), 8, 40,
                        ffn.Resolution(res, res), 4)[:1]
        fs = ffn.RaySampler(BOUNDS, cam, S, False, model, batch, device=dev)
        rays = fs._valid_for_camera(0)
        lo, hi = parallel.shard_range(len(rays))
        counts = [parallel.shard_range(len(rays), r, world)[1] - parallel.shard_range(len(rays), r, world)[0]
                  for r in range(world)]
        times = []
        with torch.no_grad():
            for rep in range(reps + 1):
                barrier()
                t0 = time.perf_counter()
                mine = fs.sample(rays[lo:hi], None)
                cols, alphas = [], []
                for s0 in range(0, hi - lo, batch):
                    out = rc.render(mine.subset(range(s0, min(s0 + batch, hi - lo))), False)
                    cols.append(out.color), alphas.append(out.alpha)
                color, alpha = parallel.gather_render(torch.cat(cols), torch.cat(alphas), counts)
                img = fs.to_image(0, color, "RGB")
                torch.cuda.synchronize()
                if rep:
                    times.append(time.perf_counter() - t0)
        assert img.shape == (res, res, 3)
        (ms,) = max_over_ranks(1e3 * float(np.median(times)))
        return {"ms": round(ms, 2), "rays_s": round(len(rays) / ms * 1e3), "valid_rays": len(rays), "res": res, "spp": S,
                "batch": batch, "evals_per_ray": S + S // 2}

    leg("frame", frame_leg)

    if world == 1:
        # ---- more samples per ray (config 3: 128 = 64+64, README 192) through the same fused kernel -----------------
        def spp_leg():
            out = {}
            n = 1 << 17
            src = dev_bundles[W]
            with torch.no_grad():
                for S in (128, 192):
                    b = ffn.RayBundle(src.starts[:n], src.directions[:n], src.near[:n], src.far[:n], src.rays[:n], S,
                                      True, None, seed=5)
                    l0 = _lib.launch_count()
                    rc.render(b, True)
                    per = _lib.launch_count() - l0
                    torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for _ in range(3):
                        rc.render(b, True)
                    e1.record()
                    torch.cuda.synchronize()
                    ms = e0.elapsed_time(e1) / 3
                    out["s%d" % S] = {"rays_s": round(n / ms * 1e3), "launches": int(per),
                                      "frac_tensor": round(n * S * FLOP_PER_SAMPLE / (ms / 1e3) / 1e12 / pk["tensor"], 3)}
            return out

        leg("spp", spp_leg)

        # ---- the precise operand mode (fp16 hi + residual split, three UMMAs per product): fp32-grade pixels --------
        def precise_leg():
            import copy
            n = 1 << 17
            src = dev_bundles[W]
            pm = copy.deepcopy(model)
            pm.ffn_operand = "fp16x3"
            prc = ffn.Raycaster(pm)
            b = ffn.RayBundle(src.starts[:n], src.directions[:n], src.near[:n], src.far[:n], src.rays[:n], SAMPLES, True,
                              None, seed=5)
            with torch.no_grad():
                prc.render(b, True)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(3):
                    prc.render(b, True)
                e1.record()
                torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 3
            return {"rays_s": round(n / ms * 1e3), "operand": "fp16x3",
                    "frac_tensor_algorithmic": round(n * SAMPLES * FLOP_PER_SAMPLE / (ms / 1e3) / 1e12 / pk["tensor"], 3),
                    "frac_tensor_executed": round(3 * n * SAMPLES * FLOP_PER_SAMPLE / (ms / 1e3) / 1e12 / pk["tensor"], 3)}

        leg("precise", precise_leg)

        # ---- the real reference on the host cores (bounded sample) and on this GPU ---------------------------------
        ref_state = {k: v.detach().cpu() for k, v in model.state_dict().items()}

        def cpu_leg():
            cores = host_cores()
            torch.set_num_threads(cores)
            arm = ReferenceCPU(1, "cpu", ref_state)
            n = min(args.ref_rays, len(arm.valid))
            rng = np.random.default_rng(3)
            arm.step(np.sort(rng.choice(arm.valid, 4096, replace=False)))          # warm the thread pool
            idx = np.sort(rng.choice(arm.valid, n, replace=False))
            t0 = time.perf_counter()
            arm.step(idx)
            cpu_s = time.perf_counter() - t0
            return {"value": n / cpu_s, "unit": "rays/s", "cores": torch.get_num_threads(), "kind": arm.kind,
                    "sample": arm.describe(n, torch.get_num_threads())}

        def cpu_baseline():
            out = cpu_leg()
            line["cpu_baseline"] = out
            return None

        leg("cpu", cpu_baseline)

        def torchgpu_leg():
            """north_star's ">= 10x the reference's 1-GPU PyTorch" denominator: the unmodified reference package with
            device='cuda' (fp32, TF32 off = PyTorch default).  kernel: ``Raycaster.render`` on pre-staged device
            RaySamples; e2e: ``batched_render(sampler.sample(idx), 4096)`` incl. host sampling, H2D, per-batch D2H."""
            arm = ReferenceCPU(1, "cuda", ref_state)
            if arm.kind != "reference":
                return {"unavailable": "oracle/_ref not vendored"}
            tb, nb = 4096, 16
            rng = np.random.default_rng(4)
            idx = np.sort(rng.choice(arm.valid, tb * nb, replace=False))
            l0 = _lib.launch_count()
            with torch.no_grad():
                staged = [arm.sampler.sample(torch.from_numpy(idx[i * tb:(i + 1) * tb]), None).to("cuda") for i in range(nb)]
                arm.rc.render(staged[0], True)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for s in staged:
                    ref_out = arm.rc.render(s, True)
                e1.record()
                torch.cuda.synchronize()
                kern_s = e0.elapsed_time(e1) / 1e3
                arm.step(idx[:tb])
                t0 = time.perf_counter()
                arm.step(idx)
                torch.cuda.synchronize()
                e2e_s_ = time.perf_counter() - t0
                ours = rc.render(ffn.RaySamples(*[None if t is None else t.to(dev) for t in staged[-1]]), True)
            assert _lib.launch_count() - l0 == 1, "the reference arm must not touch libffn_b200"
            return {"kernel_rays_s": round(tb * nb / kern_s), "e2e_rays_s": round(tb * nb / e2e_s_),
                    "what": "unmodified reference on cuda, fp32, %d x %d rays" % (nb, tb),
                    "color_max_abs_vs_ours": float((ours.color - ref_out.color).abs().max())}

        leg("torchgpu", torchgpu_leg)

        # ---- parity of the headline configuration against the oracle -------------------------------------------------
        def parity_leg():
            import oracle
            n = args.cpu_rays
            b = dev_bundles[W]
            o, d = b.starts[:n].cpu().numpy(), b.directions[:n].cpu().numpy()
            near, far = b.near[:n].cpu().numpy(), b.far[:n].cpu().numpy()
            u = np.random.default_rng(1).random((n, SAMPLES), dtype=np.float32)
            params = {k: v.numpy() for k, v in ref_state.items()}
            samples = oracle.sample_rays(o, d, near, far, SAMPLES, u=u)
            ref = oracle.render_rays(lambda p, v: oracle.nerf_forward(params, p, v), samples, True, True, 8192)
            with torch.no_grad():
                bb = ffn.RayBundle(b.starts[:n], b.directions[:n], b.near[:n], b.far[:n], b.rays[:n], SAMPLES, True,
                                   torch.from_numpy(u).to(dev))
                ours = rc.render(bb, True).numpy()
            mse = float(np.mean((ours.color - ref.color) ** 2))
            return {"rays": n, "color_max_abs": float(np.abs(ours.color - ref.color).max()),
                    "alpha_max_abs": float(np.abs(ours.alpha - ref.alpha).max()),
                    "depth_mismatch_frac": float((ours.depth != ref.depth).mean()),
                    "psnr_vs_ref_db": float(-10 * np.log10(max(mse, 1e-20)))}

        leg("parity", parity_leg)

    if rank == 0:
        line.update(extra)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
