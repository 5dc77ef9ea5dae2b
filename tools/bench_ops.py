"""Stand-alone timing of the HBM-bound helper kernels at production shapes (512 forwards of RSGNet-W32 256x192):
python tools/bench_ops.py fuse|head|stem|trptail [N]"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rsgnet_b200 import _engine, _lib  # noqa: E402
if os.environ.get('RSG_DBG'):
    _lib.use_library(os.path.join(os.path.dirname(_lib.LIB_PATH), 'librsg_b200_dbg.so'))
from rsgnet_b200._engine import PlanBuilder, View  # noqa: E402

what = sys.argv[1]
N = int(sys.argv[2]) if len(sys.argv) > 2 else 512
REPS = 6
pb = PlanBuilder(N, reuse=False)
nbytes = 0
ext = (C.c_void_p * _engine.N_EXT)()
keep = []
for r in range(REPS):       # independent buffers per repetition: nothing is L2-resident from the previous launch
    if what == 'fuse':
        H, W, Cc = 64, 48, 32
        terms = [(View(pb.buf(f'x{r}', H, W, Cc)), 0)] + [(View(pb.buf(f't{r}{s}', H >> s, W >> s, Cc)), s) for s in (1, 2, 3)]
        out = View(pb.buf(f'o{r}', H, W, Cc))
        pb.simple('fuse', dict(terms=terms, dst=out, relu=True), [t.buf for t, _ in terms], [out.buf])
        nbytes = N * H * W * Cc * 2 * (2 + 1 / 4 + 1 / 16 + 1 / 64)
    elif what == 'head':
        H, W, Cc, K = 128, 96, 32, 14
        x = pb.buf(f'x{r}', H, W, Cc)
        of = pb.buf(f'o{r}', H, W, K, itemsize=4)
        rs = np.random.RandomState(0)
        pb.conv(View(x), rs.standard_normal((K, Cc, 1, 1)), np.zeros(K), out_f32=of)
        nbytes = N * H * W * (Cc * 2 + K * 4)
    elif what == 'stem':
        H, W = 256, 192
        c1 = pb.buf(f'c{r}', H // 2, W // 2, 64)
        rs = np.random.RandomState(0)
        pb.simple('stem', dict(x=('ext', _engine.EXT_X, 0), H=H, W=W, w=pb.const(rs.standard_normal((27, 64)).astype(np.float32)),
                               bias=pb.const(np.zeros(64, np.float32)), out=c1), [], [c1])
        nbytes = (N // 2) * 3 * H * W * 4 + N * (H // 2) * (W // 2) * 64 * 2
    elif what == 'trptail':
        H, W, Cc = 64, 48, 32
        y32 = pb.buf(f'y{r}', H, W, Cc, itemsize=4)
        o = pb.buf(f'o{r}', H, W, 2 * Cc)
        rs = np.random.RandomState(0)
        f32 = lambda a: pb.const(np.ascontiguousarray(a, np.float32))
        pb.simple('trptail', dict(y32=y32, out=View(o, 0, Cc), w=f32(rs.standard_normal((Cc, Cc))), bias=f32(np.zeros(Cc)),
                                  gamma=f32(np.ones(Cc)), beta=f32(np.zeros(Cc)), groups=8, S=H * W, C=Cc), [y32], [o])
        nbytes = N * H * W * Cc * (4 + 2)
pb.allocate('cuda')
h = C.c_void_p()
_lib.check(_lib.lib().rsg_plan_create(C.byref(h), N))
_engine.emit(pb, h)
for b in pb.bufs:
    t = pb.tensor_of(b)
    t.normal_() if t.dtype != torch.float32 else t.normal_()
if what == 'stem':
    xin = torch.randn(N // 2, 3, 256, 192, device='cuda')
    ext[_engine.EXT_X] = xin.data_ptr()
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    run = lambda: _lib.check(_lib.lib().rsg_plan_run(h, _lib.stream_ptr(), ext, _engine.N_EXT, N, N // 2 if what == 'stem' else N, 0, 1))
    for _ in range(3):
        run()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s)
    iters = 10
    for _ in range(iters):
        run()
    e1.record(s)
s.synchronize()
ms = e0.elapsed_time(e1) / iters / REPS
print(f'{what} N={N}: {ms * 1e3:.1f} us  {nbytes / ms / 1e6:.0f} GB/s (algorithmic)')
