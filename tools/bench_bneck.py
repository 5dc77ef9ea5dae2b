"""Stand-alone timing of the fused Bottleneck op under its debug skip modes.
python tools/bench_bneck.py [Cin] [H] [W] [N] [skip,skip,...]   (RSG_BNECK_SKIP bits: 1 no stores, 2 no residual loads,
4 conv2 one tap, 16 no TMA)"""
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

cin = int(sys.argv[1]) if len(sys.argv) > 1 else 256
H = int(sys.argv[2]) if len(sys.argv) > 2 else 64
W = int(sys.argv[3]) if len(sys.argv) > 3 else 48
N = int(sys.argv[4]) if len(sys.argv) > 4 else 512
skips = [int(v) for v in sys.argv[5].split(',')] if len(sys.argv) > 5 else [0]
rs = np.random.RandomState(0)
w1 = rs.standard_normal((64, cin, 1, 1)) / cin ** 0.5
w2 = rs.standard_normal((64, 64, 3, 3)) / 24.0
w3 = rs.standard_normal((256, 64, 1, 1)) / 8.0
REPS = 6
pb = PlanBuilder(N, reuse=False)
# REPS independent (x, res, out) triples so that nothing is L2-resident from the previous launch
f32 = lambda n: pb.const(np.zeros(n, np.float32))
for r in range(REPS):
    x = View(pb.buf(f'x{r}', H, W, cin))
    res = x if cin == 256 else View(pb.buf(f'r{r}', H, W, 256))
    dst = View(pb.buf(f'o{r}', H, W, 256))
    pb.simple('bneck', dict(src=x, dst=dst, res=res, Cin=cin, w1=pb.const(_engine._pack_tc5(w1)), b1=f32(64),
                            w2=pb.const(_engine._pack_tc5(w2)), b2=f32(64), w3=pb.const(_engine._pack_tc5(w3[_engine._quad_perm(256)])), b3=f32(256),
                            name='bneck'), [x.buf, res.buf], [dst.buf])
pb.allocate('cuda')
h = C.c_void_p()
_lib.check(_lib.lib().rsg_plan_create(C.byref(h), N))
_engine.emit(pb, h)
for b in pb.bufs:
    pb.tensor_of(b).normal_()
ext = (C.c_void_p * _engine.N_EXT)()
s = torch.cuda.Stream()
fl = 2.0 * (cin * 64 + 9 * 64 * 64 + 64 * 256) * H * W * N
by = (cin + 256 + (256 if cin != 256 else 0)) * 2.0 * H * W * N
with torch.cuda.stream(s):
    run = lambda: _lib.check(_lib.lib().rsg_plan_run(h, _lib.stream_ptr(), ext, _engine.N_EXT, N, N, 0, 0))
    for sk in skips:
        os.environ['RSG_BNECK_SKIP'] = str(sk)
        for _ in range(2):
            run()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(5):
            run()
        e1.record(s)
        s.synchronize()
        ms = e0.elapsed_time(e1) / 5 / REPS
        print(f'bottleneck Cin={cin} {H}x{W} N={N} skip={sk} S={os.environ.get("RSG_BNECK_S", "-")}: {ms * 1e3:.1f} us  '
              f'{fl / ms / 1e9:.1f} TFLOP/s  {by / ms / 1e6:.0f} GB/s (DRAM-algorithmic)', flush=True)
