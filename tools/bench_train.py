"""Training-step experiments: eager step time (device and wall), per-kernel-family profile.  python tools/bench_train.py [preset] [batch]"""
import os, sys, time, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from rsgnet_b200 import presets, synth
from rsgnet_b200.models import _params, pose_rsgnet
from rsgnet_b200.train import TrainStep

key = sys.argv[1] if len(sys.argv) > 1 else 'w32_coco'
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
cfg = presets.preset(key)
torch.cuda.set_device(0)
net = pose_rsgnet.get_pose_net(cfg, True)
net.load_state_dict(_params.synth_state_dict(net, seed=3))
net = net.cuda().train()
b = synth.train_batch(B, cfg.MODEL.IMAGE_SIZE, cfg.MODEL.HEATMAP_SIZE, cfg.MODEL.NUM_JOINTS, cfg.MODEL.NUM_LIMBS, seed=1)
batch = {k: torch.from_numpy(v).cuda() for k, v in b.items()}
args = (batch['input'], batch['target'], batch['target_weight'], batch['all_ins_target'], batch['all_ins_target_weight'], batch['target_limbs'])
ts = TrainStep(net)
for i in range(3):
    L, _ = ts(*args)
    print('warm', i, L, flush=True)
print('launches', ts.last_launches, 'GFLOP', ts.last_flops / 1e9, 'mem GB', torch.cuda.max_memory_allocated() / 1e9)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter(); e0.record()
n = 5
for i in range(n):
    ts(*args, sync=False)
e1.record(); torch.cuda.synchronize()
print('eager: wall %.1f ms/step, device %.1f ms/step' % ((time.perf_counter() - t0) / n * 1e3, e0.elapsed_time(e1) / n))
ts.profile, ts.concurrent = True, False
ts(*args, sync=False)
torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0.0, 0, 0.0])
shapes = collections.defaultdict(lambda: [0.0, 0, 0.0])
for name, a, bb, fl in ts.tape.prof:
    r = agg[name.split(' ')[0]]; r[0] += a.elapsed_time(bb); r[1] += 1; r[2] += fl
    if ' ' in name:
        q = shapes[name]; q[0] += a.elapsed_time(bb); q[1] += 1; q[2] += fl
tot = sum(r[0] for r in agg.values())
print('profiled total %.1f ms' % tot)
for name, r in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print('%8.2f ms %5.1f%% n=%5d  %7.1f TF/s  %s' % (r[0], 100 * r[0] / tot, r[1], r[2] / r[0] / 1e9 if r[0] else 0, name))
for name, r in sorted(shapes.items(), key=lambda kv: -kv[1][0])[:28]:
    print('  %7.2f ms n=%3d %6.1f TF/s %s' % (r[0], r[1], r[2] / r[0] / 1e9, name))
sys.exit(0)
# the slowest individual matrix calls
rows = sorted(((a.elapsed_time(bb), name, fl) for name, a, bb, fl in ts.tape.prof if fl), reverse=True)[:12]
for ms, name, fl in rows:
    print('  %.3f ms %s %.1f GFLOP %.1f TF/s' % (ms, name, fl / 1e9, fl / ms / 1e9))
