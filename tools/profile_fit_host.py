"""Host-side profile (cProfile) of Raycaster.fit steps on the GPU: where the Python time of a step goes."""
import cProfile
import os
import pstats
import subprocess
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fourier_feature_nets_b200 as ffn  # noqa: E402

dev = torch.device("cuda:0")
tmp = tempfile.mkdtemp()
data = os.path.join(tmp, "scene.npz")
subprocess.run([sys.executable, os.path.join(ROOT, "tools", "make_synthetic_dataset.py"), data, "--resolution", "100",
                "--train", "30", "--val", "12", "--test", "2", "--steps", "64"], check=True)
torch.manual_seed(0)
np.random.seed(0)
train = ffn.ImageDataset.load(data, "train", 64, True, True).to(dev)
val = ffn.ImageDataset.load(data, "val", 64, True, False).to(dev)
model = ffn.NeRF(8, 256, 9, 10, 3, 4, [4], True).to(dev)
rc = ffn.Raycaster(model)
rc.fit(train, val, 1024, 5e-4, 20, 0, 1000000, 0.1, 250000, 0.0, [])     # warm-up (includes the step<10 validations)
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 400
orig_validate = rc._validate
rc._validate = lambda *a, **k: 1.0          # keep the profile to the optimisation steps
import time  # noqa: E402
torch.cuda.synchronize()
t0 = time.perf_counter()
orig_validate(val, 1024, 0)
torch.cuda.synchronize()
print("one _validate(val) call: %.1f ms" % ((time.perf_counter() - t0) * 1e3))
t0 = time.perf_counter()
rc.fit(train, val, 1024, 5e-4, steps, 0, 1000000, 0.1, 250000, 0.0, [])
torch.cuda.synchronize()
print("fit without validation, no profiler: %.3f ms/step" % ((time.perf_counter() - t0) / steps * 1e3))
pr = cProfile.Profile()
torch.cuda.synchronize()
pr.enable()
rc.fit(train, val, 1024, 5e-4, steps, 0, 1000000, 0.1, 250000, 0.0, [])
torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(45)
