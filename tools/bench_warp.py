"""Input side: crop affine warp + normalise for one step's worth of person crops.
256 crops of 256x192 (configs[1]'s step) cut from 64 synthetic 480x640 photos resident on the device, 4 people per
photo.  Reports crops/s and algorithmic GB/s (unique source footprint read once + f32 NCHW output written once)
against the measured HBM peak, checks a slice against the CPU oracle, and times the CPU oracle's counterpart
(cv2.warpAffine + LUT when OpenCV is importable, else the NumPy restatement) on a bounded sample.
python tools/bench_warp.py [n_crops]"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import warp_oracle  # noqa: E402
from rsgnet_b200.utils import transforms  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
SIZE = (192, 256)
pk = os.path.join(ROOT, 'MEASURED_PEAKS.json')
peak = json.load(open(pk))['hbm_gbs'] if os.path.isfile(pk) else 6650.0
rs = np.random.RandomState(0)
n_img = max(1, N // 4)
imgs = [rs.randint(0, 256, (480, 640, 3)).astype(np.uint8) for _ in range(min(n_img, 8))]
dev = [torch.from_numpy(imgs[i % len(imgs)]).cuda().clone() for i in range(n_img)]       # distinct device buffers
idx = np.arange(N) % n_img
centers = np.stack([rs.uniform(100, 540, N), rs.uniform(80, 400, N)], axis=1).astype(np.float32)
sc = rs.uniform(0.6, 1.6, N).astype(np.float32)                                            # boxes 120..320 px wide
scales = np.stack([sc, sc * 1.25], axis=1)
mats = transforms.affine_matrices(centers, scales, 0, SIZE)

# unique source bytes under each crop's footprint (clipped to the image), for the algorithmic traffic
w_src = sc * 200.0
h_src = w_src * SIZE[1] / SIZE[0]
x0, x1 = np.clip(centers[:, 0] - w_src / 2, 0, 640), np.clip(centers[:, 0] + w_src / 2, 0, 640)
y0, y1 = np.clip(centers[:, 1] - h_src / 2, 0, 480), np.clip(centers[:, 1] + h_src / 2, 0, 480)
src_bytes = float(((x1 - x0) * (y1 - y0) * 3).sum())
out_bytes = float(N * 3 * SIZE[0] * SIZE[1] * 4)

s = torch.cuda.Stream()
with torch.cuda.stream(s):
    run = lambda: transforms.warp_crops(dev, mats, SIZE, image_index=idx, color_rgb=True)
    x = run()
    s.synchronize()
    # kernel-only timing through the C ABI with the small tables already resident (what a loader loop would do)
    import ctypes as C
    from rsgnet_b200 import _lib
    meta = np.array([[dev[j].data_ptr(), 480, 640, 640 * 3] for j in idx], np.int64)
    ptrs = torch.from_numpy(meta[:, 0].copy()).cuda()
    dims = torch.from_numpy(meta[:, 1:].astype(np.int32)).cuda()
    md = torch.from_numpy(mats.reshape(-1, 6)).cuda()
    lut = torch.from_numpy(transforms.normalize_lut()).cuda()
    out = torch.empty_like(x)
    call = lambda: _lib.check(_lib.lib().rsg_warp_affine(_lib.stream_ptr(), C.c_void_p(ptrs.data_ptr()), C.c_void_p(dims.data_ptr()),
                                                         C.c_void_p(md.data_ptr()), N, SIZE[1], SIZE[0], 1, None,
                                                         C.c_void_p(out.data_ptr()), C.c_void_p(lut.data_ptr())))
    for _ in range(3):
        call()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s)
    for _ in range(20):
        call()
    e1.record(s)
    s.synchronize()
    ms = e0.elapsed_time(e1) / 20
    t0 = time.perf_counter()
    for _ in range(5):
        run()
    s.synchronize()
    ms_api = (time.perf_counter() - t0) / 5 * 1e3
assert torch.equal(out, x)
# parity on a slice
for i in range(0, N, max(1, N // 6)):
    ref = warp_oracle.to_tensor_normalize(warp_oracle.warp_affine_u8(imgs[idx[i] % len(imgs)], mats[i], SIZE)[:, :, ::-1])
    assert np.array_equal(x[i].cpu().numpy(), ref), i
res = {'n_crops': N, 'kernel_ms': ms, 'crops_per_s': N / ms * 1e3, 'algorithmic_MB': (src_bytes + out_bytes) / 1e6,
       'GBps': (src_bytes + out_bytes) / ms / 1e6, 'frac_of_measured_hbm_peak': (src_bytes + out_bytes) / ms / 1e6 / peak,
       'api_ms_incl_host_matrices_and_table_uploads': ms_api, 'parity': 'bit-exact vs oracle on a slice'}
# CPU counterpart on a bounded sample
try:
    import cv2
    cv2.setNumThreads(0)
    lut_np = warp_oracle.normalize_lut()
    k = min(N, 64)
    t0 = time.perf_counter()
    for i in range(k):
        u = cv2.warpAffine(cv2.cvtColor(imgs[idx[i] % len(imgs)], cv2.COLOR_BGR2RGB), mats[i], SIZE, flags=cv2.INTER_LINEAR)
        _ = np.stack([lut_np[c][u[:, :, c]] for c in range(3)])
    dt = time.perf_counter() - t0
    res['cpu'] = {'kind': 'reference (cv2 %s + table)' % cv2.__version__, 'cores': 1, 'sample': f'{k} crops', 'crops_per_s': k / dt}
except ImportError:
    k = min(N, 8)
    t0 = time.perf_counter()
    for i in range(k):
        warp_oracle.to_tensor_normalize(warp_oracle.warp_affine_u8(imgs[idx[i] % len(imgs)], mats[i], SIZE)[:, :, ::-1])
    dt = time.perf_counter() - t0
    res['cpu'] = {'kind': 'port (NumPy oracle)', 'cores': 1, 'sample': f'{k} crops', 'crops_per_s': k / dt}
print(json.dumps(res))
