"""One eager training step between cudaProfilerStart/Stop (run under `ncu --profile-from-start off`)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
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
args = [torch.from_numpy(b[k]).cuda() for k in ('input', 'target', 'target_weight', 'all_ins_target', 'all_ins_target_weight', 'target_limbs')]
ts = TrainStep(net)
ts(*args)
torch.cuda.synchronize()
torch.cuda.profiler.start()
ts(*args, sync=False)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print('launches', ts.last_launches)
