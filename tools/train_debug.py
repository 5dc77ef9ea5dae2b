"""Per-tensor gradient errors of the GPU training step against the CPU oracle in float64 (debug aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import numpy as np, torch
from test_train_step_gpu import setup_case, step_args, grads_of
from oracle import train_oracle
from rsgnet_b200.train import TrainStep

key = sys.argv[1] if len(sys.argv) > 1 else 'tiny_cp'
precise = (sys.argv[2] if len(sys.argv) > 2 else '1') == '1'
g, cfg, net, sd, batch = setup_case(key)
rsg = cfg.MODEL.NAME == 'pose_rsgnet'
ts = TrainStep(net, precise=precise)
losses, outs = ts.forward_backward(*step_args(batch, rsg))
print(losses.read())
names = [str(n) for n in g['names']]
grads = grads_of(net, ts.store, names)
torch.set_num_threads(16)
cb = {k: v.cpu() for k, v in batch.items()}
L64, g64, _, o64 = train_oracle.forward_backward(sd, cfg, cb, dtype=torch.float64)
print(L64)
for a, b in zip(outs, o64):
    if a.shape == b.shape:
        print('output err', float((a.cpu().double() - b).abs().max() / b.abs().max()))
rows = []
for k in names:
    r = g64[k]
    e = float((grads[k] - r).abs().max() / max(float(r.abs().max()), 1e-30))
    rows.append((e, k, float(r.abs().max())))
if len(sys.argv) <= 3:
    rows.sort(reverse=True)
for e, k, m in (rows[:25] if len(sys.argv) <= 3 else rows):
    print(f'{e:.2e}  {k}  (max |g| {m:.2e})')
print('median', np.median([r[0] for r in rows]))
