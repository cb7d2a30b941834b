import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import oracle
import fourier_feature_nets_b200 as ffn
from fourier_feature_nets_b200 import engine
DEV = "cuda:0"
g = np.load("tests/golden/nerf_render.npz")
w = {k[2:]: g[k] for k in g.files if k.startswith("w.")}
m = ffn.NeRF(8, 256, 9, 10, 3, 4, [4], True)
m.load_state_dict({k: torch.from_numpy(v) for k, v in w.items()})
m = m.to(DEV).eval()
cuda = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
for R, S in ((300, 192), (300, 256), (300, 100), (37, 300), (300, 64), (5000, 192)):
    rng = np.random.default_rng(R * 1000 + S)
    o = np.tile(np.array([[0.1, 0.2, -4.0]], np.float32), (R, 1))
    d = rng.normal(size=(R, 3)).astype(np.float32) * 0.12 + np.array([0, 0, 1], np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    near = rng.uniform(2.8, 3.3, R).astype(np.float32)
    far = rng.uniform(4.5, 5.2, R).astype(np.float32)
    u = rng.random((R, S), dtype=np.float32)
    ref_s = oracle.sample_rays(o, d, near, far, S, u=u)
    ref = oracle.render_rays(lambda p, v: oracle.nerf_forward(w, p, v), ref_s, True)
    eng = engine.get_engine(m, torch.device(DEV))
    args = (cuda(o), cuda(d), cuda(near), cuda(far), torch.linspace(0, 1, S).to(DEV), cuda(u), True, 0, 0, S, True)
    outs = [eng.net.render_rays(*args)[:3] for _ in range(3)]
    sargs = (cuda(ref_s.positions), cuda(ref_s.view_directions), cuda(ref_s.t_values), True)
    souts = [eng.net.render_samples(*sargs) for _ in range(3)]
    def diff(x, y):
        return [float((a - b).abs().max()) for a, b in zip(x, y)]
    print(R, S, "rays run-to-run", diff(outs[0], outs[1]), diff(outs[0], outs[2]), "samples run-to-run", diff(souts[0], souts[1]),
          "rays vs samples", diff(outs[0], souts[0]),
          "vs oracle", float(np.abs(outs[0][0].cpu().numpy() - ref.color).max()), float(np.abs(souts[0][0].cpu().numpy() - ref.color).max()),
          "depth mism", float((outs[0][2].cpu().numpy() != ref.depth).mean()))
    bad = (outs[0][0] != souts[0][0]).any(-1).nonzero().flatten().tolist()
    print("   rays differing:", bad[:20], "n", len(bad))
