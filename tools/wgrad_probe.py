"""Bring-up probe of ffn_wgrad on a B200: tries the MN-major descriptor stride conventions (FFN_WG_LBO / FFN_WG_SBO),
reports max errors vs an fp32 matmul, dumps structured one-hot cases for offline decoding, and times the NeRF job list.
Usage (GPU box): python tools/wgrad_probe.py [out.npz]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fourier_feature_nets_b200 import _lib  # noqa: E402
from fourier_feature_nets_b200.autograd import WgradJob, _bind, _run_wgrad, _wg_tensor  # noqa: E402

DEV = "cuda:0"


def job(a_slot, n_mt, b_tensor, b_slot, n_cols, dst, bias=None):
    return WgradJob(0, a_slot, 0, n_mt, b_tensor, b_slot, 0, n_cols, dst.data_ptr(), dst.shape[1], 0, n_cols, 0,
                    0 if bias is None else bias.data_ptr())


def main():
    L = _lib.lib()
    _bind(L)
    out = {}
    g = torch.Generator(device=DEV).manual_seed(0)
    M = 4096
    dz = (torch.randn((1, M, 256), device=DEV, generator=g) * 0.5).to(torch.bfloat16)
    x = torch.randn((1, M, 256), device=DEV, generator=g).to(torch.bfloat16)
    xh = torch.randn((1, M, 64), device=DEV, generator=g).to(torch.bfloat16)
    ref = dz[0].float().t() @ x[0].float()
    refh = dz[0].float().t() @ xh[0].float()
    for lbo, sbo in ((8192, 1024), (1024, 8192)):
        os.environ["FFN_WG_LBO"], os.environ["FFN_WG_SBO"] = str(lbo), str(sbo)
        dW = torch.zeros((256, 256), device=DEV)
        db = torch.zeros((256,), device=DEV)
        _run_wgrad(L, [_wg_tensor(dz), _wg_tensor(x)], [job(0, 2, 1, 0, 256, dW, db)])
        dWh = torch.zeros((256, 64), device=DEV)
        _run_wgrad(L, [_wg_tensor(dz), _wg_tensor(xh)], [job(0, 2, 1, 0, 64, dWh)])
        torch.cuda.synchronize()
        print("lbo %5d sbo %5d: bf16xbf16 max err %.3e (max ref %.1f)  64-col max err %.3e  bias err %.3e" % (
            lbo, sbo, (dW - ref).abs().max().item(), ref.abs().max().item(), (dWh - refh).abs().max().item(),
            (db - dz[0].float().sum(0)).abs().max().item()), flush=True)
    # structured cases with the default strides: A one-hot row m0 with value (o+1) / B one-hot row m0 with value (n+1)
    os.environ.pop("FFN_WG_LBO"), os.environ.pop("FFN_WG_SBO")
    for m0 in (0, 1, 8, 17, 63):
        a = torch.zeros((1, 64, 256), device=DEV)
        b = torch.zeros((1, 64, 256), device=DEV)
        a[0, m0] = torch.arange(1, 257, device=DEV).float()
        b[0, m0] = 1.0
        d = torch.zeros((256, 256), device=DEV)
        ab, bb = a.to(torch.bfloat16), b.to(torch.bfloat16)      # keep alive: the descriptors hold raw pointers
        _run_wgrad(L, [_wg_tensor(ab), _wg_tensor(bb)], [job(0, 2, 1, 0, 256, d)])
        out["A_m%d" % m0] = d.cpu().numpy()
        a[0, m0] = 1.0
        b[0, m0] = torch.arange(1, 257, device=DEV).float()
        d2 = torch.zeros((256, 256), device=DEV)
        ab2, bb2 = a.to(torch.bfloat16), b.to(torch.bfloat16)
        _run_wgrad(L, [_wg_tensor(ab2), _wg_tensor(bb2)], [job(0, 2, 1, 0, 256, d2)])
        out["B_m%d" % m0] = d2.cpu().numpy()
        ok = bool((d == torch.arange(1, 257, device=DEV).float()[:, None]).all()) and \
            bool((d2 == torch.arange(1, 257, device=DEV).float()[None, :]).all())
        print("one-hot m0=%d exact: %s" % (m0, ok), flush=True)
    if len(sys.argv) > 1:
        np.savez_compressed(sys.argv[1], **out)
    # timing: the NeRF job list at 1024 rays x 128 samples
    M = 131072
    dz = torch.randn((10, M, 256), device=DEV).to(torch.bfloat16)
    sh = torch.randn((10, M, 256), device=DEV).to(torch.bfloat16)
    enc = torch.randn((2, M, 64), device=DEV).to(torch.bfloat16)
    Ws = [torch.zeros((256, 256), device=DEV) for _ in range(12)]
    bs = [torch.zeros((256,), device=DEV) for _ in range(12)]
    jobs = [WgradJob(0, 0, 0, 2, 2, 0, 0, 64, Ws[0].data_ptr(), 256, 0, 64, 0, bs[0].data_ptr())]
    for i in range(1, 8):
        jobs.append(job(i, 2, 1, i - 1, 256, Ws[i], bs[i]))
    jobs.append(WgradJob(0, 4, 0, 2, 2, 0, 0, 64, Ws[10].data_ptr(), 256, 0, 64, 0, 0))
    jobs.append(job(8, 2, 1, 7, 256, Ws[8], bs[8]))
    jobs.append(job(9, 1, 1, 8, 256, Ws[9], bs[9]))
    jobs.append(WgradJob(0, 9, 0, 1, 2, 1, 0, 64, Ws[11].data_ptr(), 256, 0, 64, 0, 0))
    tens = [_wg_tensor(dz), _wg_tensor(sh), _wg_tensor(enc)]
    for _ in range(3):
        _run_wgrad(L, tens, jobs)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    n = 20
    for _ in range(n):
        _run_wgrad(L, tens, jobs)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    gb = M * (7 * 1024 + 2 * 640 + 1024 + 768 + 384) / 1e9
    print("NeRF wgrad job list, M=%d: %.3f ms per launch, %.2f GB streamed -> %.0f GB/s" % (M, ms, gb, gb / ms * 1e3))
    ref8 = dz[8].float().t() @ sh[7].float()
    print("timed-case check: max err %.3e of %.1f" % ((Ws[8] / (n + 3) - ref8).abs().max().item(), ref8.abs().max().item()))
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(n):
        for i in range(1, 8):
            torch.mm(dz[i].t(), sh[i - 1], out_dtype=torch.float32)
        torch.mm(dz[8].t(), sh[7], out_dtype=torch.float32)
        torch.mm(dz[9][:, :128].t(), sh[8], out_dtype=torch.float32)
    t1.record()
    torch.cuda.synchronize()
    print("cuBLAS (torch.mm, 9 hidden GEMMs only, no bias sums): %.3f ms" % (t0.elapsed_time(t1) / n))


if __name__ == "__main__":
    main()
