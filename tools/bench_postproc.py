"""BASELINE.json configs[3]: post-processing only -- get_final_preds decode + oks_nms on 100k synthetic
detections x 17 heat-maps at 64x48.  Reports decode GB/s against the measured HBM peak and NMS dets/s,
with the CPU oracle timed on a bounded slice beside them.   python tools/bench_postproc.py [N]"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import decode_oracle, nms_oracle  # noqa: E402
from rsgnet_b200 import synth  # noqa: E402
from rsgnet_b200.core.inference import decode_device  # noqa: E402
from rsgnet_b200.nms.nms import oks_nms_batched  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
K, H, W = 17, 64, 48
peak = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs'] if os.path.isfile(os.path.join(ROOT, 'MEASURED_PEAKS.json')) else 6650.0
# heat-maps generated on the device (20.9 GB for N=100k): smooth bump + noise, 15 % dead maps
g = torch.Generator(device='cuda').manual_seed(2)
hm = torch.empty((N, K, H, W), device='cuda')
ys = torch.arange(H, device='cuda').view(1, 1, H, 1).float()
xs = torch.arange(W, device='cuda').view(1, 1, 1, W).float()
for lo in range(0, N, 5000):
    n = min(5000, N - lo)
    cx = torch.rand((n, K, 1, 1), device='cuda', generator=g) * (W + 1) - 1
    cy = torch.rand((n, K, 1, 1), device='cuda', generator=g) * (H + 1) - 1
    amp = torch.rand((n, K, 1, 1), device='cuda', generator=g) * 0.9 + 0.1
    t = amp * torch.exp(-((xs - cx) ** 2 + (ys - cy) ** 2) / 8.0) + 0.01 * torch.randn((n, K, H, W), device='cuda', generator=g)
    dead = torch.rand((n, K, 1, 1), device='cuda', generator=g) < 0.15
    hm[lo:lo + n] = torch.where(dead, t - 2.0, t)
c_np, s_np = synth.centers_scales(N, seed=5)
c, s = torch.from_numpy(c_np).cuda(), torch.from_numpy(s_np).cuda()


def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


ms = timeit(lambda: decode_device(hm, c, s, post_process=True))
byt = N * K * H * W * 4 + N * K * 12
res = {'decode': {'ms': ms, 'crops_per_s': N / ms * 1e3, 'algorithmic_GB': byt / 1e9, 'GBps': byt / ms / 1e6,
                  'frac_of_measured_hbm_peak': byt / ms / 1e6 / peak}}
# flip-fused variant on half the data (two input tensors)
h2 = N // 2
perm = np.arange(K, dtype=np.int32); perm[1:] = perm[1:].reshape(-1, 2)[:, ::-1].reshape(-1)
ms2 = timeit(lambda: decode_device(hm[:h2], c[:h2], s[:h2], post_process=True, hm_flipped=hm[h2:2 * h2], flip_perm=perm))
byt2 = 2 * h2 * K * H * W * 4 + h2 * K * 12
res['flip_avg_decode'] = {'ms': ms2, 'crops_per_s': h2 / ms2 * 1e3, 'GBps': byt2 / ms2 / 1e6,
                          'frac_of_measured_hbm_peak': byt2 / ms2 / 1e6 / peak}
# parity on a slice + CPU oracle timing
sl = 512
out = decode_device(hm[:sl], c[:sl], s[:sl], post_process=True)
t0 = time.perf_counter()
o_preds, o_mv = decode_oracle.get_final_preds(True, hm[:sl].cpu().numpy(), c_np[:sl], s_np[:sl])
cpu_dec = time.perf_counter() - t0
res['decode']['bit_exact_vs_oracle'] = bool(np.array_equal(out['preds'].cpu().numpy(), o_preds) and np.array_equal(out['maxvals'].cpu().numpy(), o_mv))
res['decode']['cpu_oracle_crops_per_s'] = sl / cpu_dec
# NMS: N detections, 20 per image
kpts, scores, areas, off = synth.detections(N // 20, 20, K, seed=3)
dk, ds, da = torch.from_numpy(kpts).cuda(), torch.from_numpy(scores).cuda(), torch.from_numpy(areas).cuda()
msn = timeit(lambda: oks_nms_batched(dk, ds, da, off, 0.9), reps=3)
keep, counts = oks_nms_batched(dk, ds, da, off, 0.9)
t0 = time.perf_counter()
ok = True
for i in range(200):
    ref, _ = nms_oracle.oks_nms_arrays(kpts[off[i]:off[i + 1]], scores[off[i]:off[i + 1]], areas[off[i]:off[i + 1]], 0.9)
    ok &= list(keep[off[i]:off[i] + counts[i]]) == ref
cpu_nms = time.perf_counter() - t0
res['oks_nms'] = {'ms_incl_host_copies': msn, 'dets_per_s': len(scores) / msn * 1e3, 'kept_frac': float(counts.sum()) / len(scores),
                  'keep_lists_equal_oracle_200_images': bool(ok), 'cpu_oracle_dets_per_s': 200 * 20 / cpu_nms}
print(json.dumps(res, indent=1))
