timeout 600 python tools/train_parity_r2.py --steps 20000 --report 2000 --scene solid --binary-alpha --skip-reference --out gpurun_out/r02_train_parity_ours.json 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read())
for k in d:
    if k.startswith('inference'): print(k, d[k])"
mkdir -p /tmp/w && cd /tmp/w
python $GRAFT_REPO_ROOT/tools/make_synthetic_dataset.py toy.npz --resolution 48 --train 10 --val 2 --test 1 --steps 96 > /dev/null
python $GRAFT_REPO_ROOT/tools/run_reference_script.py train_nerf.py toy.npz nerf_out --device cuda --num-steps 300 --batch-size 1024 --num-samples 32 --image-interval 150 --report-interval 100 --crop-steps 50 --num-anneal-steps 100 | tail -1
for op in fp16 fp16x3; do python $GRAFT_REPO_ROOT/tools/frame_parity_probe.py nerf_out/nerf.pt --operand $op | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$op', [(x['t_max_abs'], x['t_rays_over_1e-3'], x['pix_fused(t_ours)_vs_torch(t_ref)']) for x in d])"; done
