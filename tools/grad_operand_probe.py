import sys, os, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import fourier_feature_nets_b200 as ffn
from test_gpu_training import trained_like_model, make_batch, loss_fn, DEV
for operand in ("fp16", "bf16"):
    model = trained_like_model()
    model.ffn_operand = operand
    R, S = 256, 64
    bundle = make_batch(R, S).to(DEV)
    g = torch.Generator(device=DEV).manual_seed(1)
    gt_c = torch.rand((R, 3), device=DEV, generator=g); gt_a = torch.rand((R,), device=DEV, generator=g)
    rc = ffn.Raycaster(model)
    rc.train_kernels = False
    model.zero_grad(); out_ref = rc.render(bundle, True); loss_fn(out_ref, gt_c, gt_a).backward()
    ref = {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None}
    rc.train_kernels = True
    model.zero_grad(); out = rc.render(bundle, True); loss_fn(out, gt_c, gt_a).backward()
    worst_cos, worst_rel = 1.0, 0.0
    for n, p in model.named_parameters():
        if n not in ref: continue
        a, b = p.grad.flatten().double(), ref[n].flatten().double()
        cos = (a @ b / (a.norm() * b.norm() + 1e-30)).item(); rel = ((a - b).norm() / (b.norm() + 1e-30)).item()
        worst_cos = min(worst_cos, cos); worst_rel = max(worst_rel, rel)
    print(operand, "color max abs %.2e" % (out.color - out_ref.color).abs().max().item(), "worst cos %.6f worst rel %.4f" % (worst_cos, worst_rel))
