"""GPU tests of ``ffn_wgrad`` (``-m gpu``): the split-K tcgen05 weight-gradient kernel against fp32 matmuls of the same
16-bit operands (torch fp32 reference of a floating-point kernel).

Tolerance: operands are exactly representable (bf16 / fp16 inputs), products are exact in fp32, so the only
difference is fp32 summation order: |err| <= 2e-5 * sqrt(rows) * max|dW| is generous; bias sums likewise."""
import pytest
import torch

from fourier_feature_nets_b200 import _lib
from fourier_feature_nets_b200.autograd import (WgradJob, WgradTensor, _bind, _run_wgrad, _wg_tensor, enc_colmap,
                                                enc_permutation)

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _data(M, slots_a=2, slots_b=2, cols_b=256, dtype_b=torch.bfloat16, seed=0):
    g = torch.Generator(device=DEV).manual_seed(seed)
    dz = (torch.randn((slots_a, M, 256), device=DEV, generator=g) * 0.5).to(torch.bfloat16)
    x = torch.randn((slots_b, M, cols_b), device=DEV, generator=g).to(dtype_b)
    return dz, x


def _job(a_slot, a_col0, n_mt, b_tensor, b_slot, b_col0, n_cols, dst, dst_col0, dst_cols, colmap=None, bias=None):
    return WgradJob(0, a_slot, a_col0, n_mt, b_tensor, b_slot, b_col0, n_cols, dst.data_ptr(), dst.shape[1], dst_col0,
                    dst_cols, 0 if colmap is None else colmap.data_ptr(), 0 if bias is None else bias.data_ptr())


def _tol(ref, M):
    return 2e-5 * (M ** 0.5) * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("M", [64, 200, 4096, 65536 + 40])
def test_full_layer_and_bias(M):
    L = _lib.lib()
    _bind(L)
    dz, x = _data(M)
    dW = torch.zeros((256, 256), device=DEV)
    db = torch.zeros((256,), device=DEV)
    before = _lib.launch_count()
    _run_wgrad(L, [_wg_tensor(dz), _wg_tensor(x)], [_job(1, 0, 2, 1, 0, 0, 256, dW, 0, 256, None, db)])
    torch.cuda.synchronize()
    assert _lib.launch_count() == before + 1
    ref = dz[1].float().t() @ x[0].float()
    ref_b = dz[1].float().sum(0)
    assert (dW - ref).abs().max().item() <= _tol(ref, M), ((dW - ref).abs().max().item(), ref.abs().max().item())
    assert (db - ref_b).abs().max().item() <= _tol(ref_b, M)


def test_unaligned_destination_half_tile_and_window():
    """hidden_view-like job: 128 outputs, destination row stride 283 (scalar reds), plus a column window of x."""
    L = _lib.lib()
    _bind(L)
    M = 3000
    dz, x = _data(M, cols_b=512, seed=1)
    dW = torch.zeros((128, 283), device=DEV)
    db = torch.zeros((128,), device=DEV)
    dW2 = torch.zeros((256, 510), device=DEV)
    jobs = [_job(0, 0, 1, 1, 1, 0, 256, dW, 0, 256, None, db),
            _job(1, 0, 2, 1, 0, 256, 256, dW2, 256, 254)]
    _run_wgrad(L, [_wg_tensor(dz), _wg_tensor(x)], jobs)
    ref = dz[0, :, :128].float().t() @ x[1, :, :256].float()
    assert (dW[:, :256] - ref).abs().max().item() <= _tol(ref, M)
    assert dW[:, 256:].abs().max().item() == 0.0
    assert (db - dz[0, :, :128].float().sum(0)).abs().max().item() <= _tol(ref, M)
    ref2 = dz[1].float().t() @ x[0, :, 256:510].float()
    assert (dW2[:, 256:] - ref2).abs().max().item() <= _tol(ref2, M)
    assert dW2[:, :256].abs().max().item() == 0.0


def test_encoding_operand_with_colmap():
    """layer-0 / skip-layer job: B is the saved encoding chunk (64 columns), columns un-permuted on the way out."""
    L = _lib.lib()
    _bind(L)
    M = 5000
    dz, enc = _data(M, cols_b=64, seed=2)
    dW = torch.zeros((256, 319), device=DEV)
    cm = enc_colmap(10, True, 256, torch.device(DEV))
    _run_wgrad(L, [_wg_tensor(dz), _wg_tensor(enc)], [_job(0, 0, 2, 1, 1, 0, 64, dW, 0, 64, cm)])
    perm = enc_permutation(10, True, torch.device(DEV))
    ref = (dz[0].float().t() @ enc[1].float())[:, perm]
    assert (dW[:, 256:] - ref).abs().max().item() <= _tol(ref, M)
    assert dW[:, :256].abs().max().item() == 0.0


def test_accumulates_into_destination_and_many_jobs():
    L = _lib.lib()
    _bind(L)
    M = 8192
    dz, x = _data(M, slots_a=10, slots_b=10, seed=3)
    outs = [torch.ones((256, 256), device=DEV) for _ in range(10)]
    jobs = [_job(i, 0, 2, 1, (i + 3) % 10, 0, 256, outs[i], 0, 256) for i in range(10)]
    _run_wgrad(L, [_wg_tensor(dz), _wg_tensor(x)], jobs)
    for i in range(10):
        ref = dz[i].float().t() @ x[(i + 3) % 10].float() + 1.0
        assert (outs[i] - ref).abs().max().item() <= _tol(ref, M), i


def test_bad_arguments_fail_loudly():
    L = _lib.lib()
    _bind(L)
    dz, x = _data(128)
    dW = torch.zeros((256, 256), device=DEV)
    with pytest.raises(_lib.FFNError):
        _run_wgrad(L, [_wg_tensor(dz), _wg_tensor(x)], [_job(0, 0, 2, 1, 0, 0, 100, dW, 0, 100)])
    with pytest.raises(_lib.FFNError):
        _run_wgrad(L, [_wg_tensor(dz), _wg_tensor(x)], [_job(5, 0, 2, 1, 0, 0, 256, dW, 0, 256)])
    bad = WgradTensor(x.data_ptr(), 64, 256, 2, 0)      # rows differ from tensor 0
    ta = (WgradTensor * 2)(_wg_tensor(dz), bad)
    ja = (WgradJob * 1)(_job(0, 0, 2, 1, 0, 0, 256, dW, 0, 256))
    assert L.ffn_wgrad(ta, 2, ja, 1, _lib._stream()) != 0
    assert L.ffn_last_error()


@pytest.mark.parametrize("cols", [256, 64])
def test_fp16_b_operand_is_converted_to_bf16_in_shared_memory(cols):
    """The forward saves activations in its operand dtype (fp16 by default); ffn_wgrad rounds such a B tile to bf16 in
    shared memory (kind::f16 UMMAs cannot mix fp16 with the bf16 dz).  Reference: fp32 matmul of dz with the
    bf16-rounded activations; summation-order tolerance.  Mixed with a bf16 job and a bias in the same launch."""
    L = _lib.lib()
    _bind(L)
    M = 9000
    dz, xh = _data(M, cols_b=cols, dtype_b=torch.float16, seed=4)
    _, xb = _data(M, cols_b=256, seed=5)
    xh[0, :7, :5] = torch.tensor([65504.0, -65504.0, 6e-8, -6e-8, 0.0], device=DEV, dtype=torch.float16)   # range ends
    dW = torch.zeros((256, cols), device=DEV)
    dW2 = torch.zeros((256, 256), device=DEV)
    db = torch.zeros((256,), device=DEV)
    jobs = [WgradJob(0, 1, 0, 2, 1, 0, 0, cols, dW.data_ptr(), cols, 0, cols, 0, db.data_ptr()),
            WgradJob(0, 0, 0, 2, 2, 1, 0, 256, dW2.data_ptr(), 256, 0, 256, 0, 0)]
    _run_wgrad(L, [_wg_tensor(dz), _wg_tensor(xh), _wg_tensor(xb)], jobs)
    ref = dz[1].float().t() @ xh[0].to(torch.bfloat16).float()
    assert (dW - ref).abs().max().item() <= _tol(ref, M), ((dW - ref).abs().max().item(), ref.abs().max().item())
    ref2 = dz[0].float().t() @ xb[1].float()
    assert (dW2 - ref2).abs().max().item() <= _tol(ref2, M)
    assert (db - dz[1].float().sum(0)).abs().max().item() <= _tol(ref, M)
    with pytest.raises(_lib.FFNError):          # dz (the A operand) must be bf16
        _run_wgrad(L, [_wg_tensor(xh), _wg_tensor(dz)], [WgradJob(0, 0, 0, 2, 1, 0, 0, 256, dW2.data_ptr(), 256, 0, 256, 0, 0)])
