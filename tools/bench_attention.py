"""Stand-alone timing of the TRP attention op through the C ABI.  python tools/bench_attention.py [C] [H] [W] [N]"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rsgnet_b200 import _engine, _lib  # noqa: E402
if os.environ.get('RSG_DBG'):      # the build with the kernels' RSG_* switches: make -C rsgnet_b200/csrc DEBUG_SWITCHES=1
    _lib.use_library(os.path.join(os.path.dirname(_lib.LIB_PATH), 'librsg_b200_dbg.so'))
from rsgnet_b200._engine import PlanBuilder, View  # noqa: E402

Cc = int(sys.argv[1]) if len(sys.argv) > 1 else 32
H = int(sys.argv[2]) if len(sys.argv) > 2 else 64
W = int(sys.argv[3]) if len(sys.argv) > 3 else 48
N = int(sys.argv[4]) if len(sys.argv) > 4 else 512
pb = PlanBuilder(N, reuse=False)
xb, gb, yb = pb.buf('x', H, W, Cc), pb.buf('g', H, W, Cc), pb.buf('y', H, W, Cc)
REPS = 4
for _ in range(REPS):
    pb.simple('attention', dict(x=View(xb), g=View(gb), y=View(yb)), [xb, gb], [yb])
pb.allocate('cuda')
h = C.c_void_p()
_lib.check(_lib.lib().rsg_plan_create(C.byref(h), N))
_engine.emit(pb, h)
pb.tensor_of(xb).normal_(0, 0.5)
pb.tensor_of(gb).normal_()
ext = (C.c_void_p * _engine.N_EXT)()
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    run = lambda: _lib.check(_lib.lib().rsg_plan_run(h, _lib.stream_ptr(), ext, _engine.N_EXT, N, N, 0, 1))
    for _ in range(3):
        run()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s)
    for _ in range(5):
        run()
    e1.record(s)
s.synchronize()
ms = e0.elapsed_time(e1) / 5 / REPS
S = H * W
print(f'attention C={Cc} S={S} N={N}: {ms * 1e3:.1f} us  {4.0 * S * S * Cc * N / ms / 1e9:.1f} TFLOP/s  '
      f'{S * S * N / ms / 1e6 / 148 / 1.965:.2f} sigmoids/clk/SM @1.965GHz')
