"""Stand-alone timing of the fused BasicBlock op.  python tools/bench_bb.py [C] [H] [W] [N]"""
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

Cc = int(sys.argv[1]) if len(sys.argv) > 1 else 32
H = int(sys.argv[2]) if len(sys.argv) > 2 else 64
W = int(sys.argv[3]) if len(sys.argv) > 3 else 48
N = int(sys.argv[4]) if len(sys.argv) > 4 else 512
g = torch.Generator().manual_seed(0)
sd = {}
for i in (1, 2):
    sd[f'conv{i}.weight'] = torch.randn(Cc, Cc, 3, 3, generator=g) / (Cc * 9) ** 0.5
    sd[f'bn{i}.weight'] = torch.ones(Cc); sd[f'bn{i}.bias'] = torch.zeros(Cc)
    sd[f'bn{i}.running_mean'] = torch.zeros(Cc); sd[f'bn{i}.running_var'] = torch.ones(Cc)
pb = PlanBuilder(N, reuse=False)
xin = pb.buf('x', H, W, Cc)
REPS = 10
for _ in range(REPS):
    _engine._basic(pb, _engine._Params(sd), View(xin))
pb.allocate('cuda')
h = C.c_void_p()
_lib.check(_lib.lib().rsg_plan_create(C.byref(h), N))
_engine.emit(pb, h)
pb.tensor_of(xin).normal_()
ext = (C.c_void_p * _engine.N_EXT)()
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    run = lambda: _lib.check(_lib.lib().rsg_plan_run(h, _lib.stream_ptr(), ext, _engine.N_EXT, N, N, 0, 1))
    for _ in range(3):
        run()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s)
    for _ in range(10):
        run()
    e1.record(s)
s.synchronize()
ms = e0.elapsed_time(e1) / 10 / REPS
fl = 2.0 * 2 * 9 * Cc * Cc * H * W * N
print(f'basic block C={Cc} {H}x{W} N={N} ({[k for k, _, _ in pb.ops][0]}): {ms * 1e3:.1f} us  {fl / ms / 1e9:.1f} TFLOP/s  '
      f'{2.0 * Cc * H * W * N * 2 / ms / 1e6:.0f} GB/s (algorithmic)')
