"""Stand-alone timing of one conv op through the C ABI (used for ncu captures).
python tools/bench_conv.py CIN COUT K H W N [NRES] [ENGINE] [ITERS]"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rsgnet_b200 import _engine, _lib  # noqa: E402
if os.environ.get('RSG_DBG'):      # the build with the kernels' RSG_* switches: make -C rsgnet_b200/csrc DEBUG_SWITCHES=1
    _lib.use_library(os.path.join(os.path.dirname(_lib.LIB_PATH), 'librsg_b200_dbg.so'))
from rsgnet_b200._engine import PlanBuilder, View  # noqa: E402

cin, cout, k, H, W, N = [int(v) for v in sys.argv[1:7]]
nres = int(sys.argv[7]) if len(sys.argv) > 7 else 0
engine = int(sys.argv[8]) if len(sys.argv) > 8 else 0
iters = int(sys.argv[9]) if len(sys.argv) > 9 else 20
stride = int(sys.argv[10]) if len(sys.argv) > 10 else 1
Ho, Wo = (H - 1) // stride + 1, (W - 1) // stride + 1
rs = np.random.RandomState(0)
w = rs.standard_normal((cout, cin, k, k)) / np.sqrt(cin * k * k)
pb = PlanBuilder(N, reuse=False)
xin = pb.buf('x', H, W, cin)
rb = pb.buf('r', Ho, Wo, cout)
ob = pb.buf('o', Ho, Wo, cout)
REPS = 10    # identical launches per graph replay: the host launch path must not bound the measurement
for _ in range(REPS):
    pb.conv(View(xin), w, np.zeros(cout), stride=stride, relu=True, dst=View(ob), res=[(View(rb), 0)] * nres, engine=engine)
pb.allocate('cuda')
h = C.c_void_p()
_lib.check(_lib.lib().rsg_plan_create(C.byref(h), N))
_engine.emit(pb, h)
pb.tensor_of(xin).normal_()
pb.tensor_of(rb).normal_()
ext = (C.c_void_p * _engine.N_EXT)()
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    run = lambda: _lib.check(_lib.lib().rsg_plan_run(h, _lib.stream_ptr(), ext, _engine.N_EXT, N, N, 0, 1))
    for _ in range(3):
        run()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s)
    for _ in range(iters):
        run()
    e1.record(s)
s.synchronize()
ms = e0.elapsed_time(e1) / iters / REPS
fl = 2.0 * k * k * cin * cout * Ho * Wo * N
by = (cin * H * W + cout * (1 + nres) * Ho * Wo) * N * 2.0
print(f'conv {cin}->{cout} k{k} s{stride} {H}x{W} N={N} nres={nres} engine={engine}: {ms * 1e3:.1f} us  {fl / ms / 1e9:.1f} TFLOP/s  {by / ms / 1e6:.1f} GB/s (algorithmic)')
