"""Pure-write / pure-read / copy HBM rates on this GPU (context for the training step's roofline: the forward and dgrad
passes are write streams, ffn_wgrad a read stream)."""
import json
import torch

dev = torch.device("cuda:0")
n = 1 << 30       # 1 GiB
x = torch.empty(n, dtype=torch.uint8, device=dev)
y = torch.empty(n, dtype=torch.uint8, device=dev)
xf = x.view(torch.float32)


def timed(fn, reps=10):
    fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


out = {}
out["memset_write_TBs"] = n / timed(lambda: x.zero_()) / 1e9
out["fill_f32_write_TBs"] = n / timed(lambda: xf.fill_(1.5)) / 1e9
out["sum_read_TBs"] = n / timed(lambda: xf.sum()) / 1e9
out["copy_rw_TBs"] = 2 * n / timed(lambda: y.copy_(x)) / 1e9
print(json.dumps({k: round(v, 3) for k, v in out.items()}))
