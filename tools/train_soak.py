"""Soak test of the training step: N graph-replayed steps on one fixed batch; the loss must fall and stay finite."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rsgnet_b200 import presets, synth
from rsgnet_b200.models import _params, pose_rsgnet
from rsgnet_b200.train import TrainStep
key = sys.argv[1] if len(sys.argv) > 1 else 'w32_coco'
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
N = int(sys.argv[3]) if len(sys.argv) > 3 else 200
cfg = presets.preset(key)
torch.cuda.set_device(0)
net = pose_rsgnet.get_pose_net(cfg, True)
net.load_state_dict(_params.synth_state_dict(net, seed=3))
net = net.cuda().train()
b = synth.train_batch(B, cfg.MODEL.IMAGE_SIZE, cfg.MODEL.HEATMAP_SIZE, cfg.MODEL.NUM_JOINTS, cfg.MODEL.NUM_LIMBS, seed=1)
args = [torch.from_numpy(b[k]).cuda() for k in ('input', 'target', 'target_weight', 'all_ins_target', 'all_ins_target_weight', 'target_limbs')]
ts = TrainStep(net, lr=1e-3)
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    ts.build_graph(*args)
    t0 = time.perf_counter()
    hist = []
    for i in range(N):
        L, _ = ts.step_graph(*args, sync=(i % 20 == 0 or i == N - 1))
        if isinstance(L, dict):
            hist.append(L['loss'])
            print(i, {k: round(v, 6) for k, v in L.items()}, flush=True)
    torch.cuda.synchronize()
print('steps', N, 'wall s', round(time.perf_counter() - t0, 2), 'finite', all(h == h for h in hist), 'decreasing', hist[-1] < hist[0] * 0.2)
sd = net.state_dict()
print('nan params', sum(int(torch.isnan(v).any()) for v in sd.values() if v.is_floating_point()), 'bn counter', int(sd['bn1.num_batches_tracked']))
