"""Micro-benchmark of the training implicit-GEMM kernels: tcgen05 kind::tf32 (precise=0) vs mma.sync TF32 (precise=2)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from rsgnet_b200.train.tape import Tape, T

torch.cuda.set_device(0)
dev = torch.device('cuda:0')
SHAPES = [  # N, H, W, Ci, Co, k, stride
    (32, 64, 48, 32, 32, 3, 1), (32, 32, 24, 64, 64, 3, 1), (32, 16, 12, 128, 128, 3, 1), (32, 8, 6, 256, 256, 3, 1),
    (32, 64, 48, 64, 256, 1, 1), (32, 64, 48, 256, 64, 1, 1), (32, 64, 48, 64, 64, 3, 1), (32, 64, 48, 600, 32, 3, 1),
    (32, 128, 96, 32, 32, 3, 1), (32, 64, 48, 256, 32, 3, 1), (32, 64, 48, 96, 96, 3, 1),
]
only = [int(a) for a in sys.argv[1:]]
for idx, (N, H, W, Ci, Co, k, s) in enumerate(SHAPES):
    if only and idx not in only:
        continue
    x = torch.randn(N, H, W, Ci, device=dev)
    wp = torch.randn(k * k, Ci, Co, device=dev) / (Ci * k * k) ** 0.5
    wT = wp.permute(0, 2, 1).contiguous()
    row = []
    for pr in (0, 2):
        tape = Tape(dev, pr)
        xn, wn = T(x), T(wp, req=True, g=torch.zeros_like(wp))
        out = tape.conv(xn, wn, wT, k, s, k // 2)
        out.g = torch.randn_like(out.v)
        res = {}
        for name, fn in (('fwd', lambda: tape._gemm(pr, x, wT, out.v, None, out.v.numel() // Co, Co, Ci, Ci, Ci, Co, mode=0 if k == 1 else 1, transB=1,
                                                   geom=None if k == 1 else (H, W, out.shape[1], out.shape[2], k, k, s, k // 2))),
                         ('dgrad', lambda: tape._gemm(pr, out.g, wp, x.clone(), None, N * H * W, Ci, Co, Co, Co, Ci, mode=0 if k == 1 else 2, transB=1,
                                                     geom=None if k == 1 else (out.shape[1], out.shape[2], H, W, k, k, s, k // 2))),
                         ('wgrad', lambda: tape._wgrad(pr, x, out.g, wn.g, out.v.numel() // Co, Ci, Co, mode=0 if k == 1 else 1,
                                                      geom=None if k == 1 else (H, W, out.shape[1], out.shape[2], k, k, s, k // 2)))):
            if name == 'dgrad':
                dx = torch.empty_like(x)
                fn = (lambda f=fn: tape._gemm(pr, out.g, wp, dx, None, N * H * W, Ci, Co, Co, Co, Ci, mode=0 if k == 1 else 2, transB=1,
                                              geom=None if k == 1 else (out.shape[1], out.shape[2], H, W, k, k, s, k // 2)))
            # ten launches replayed as a CUDA graph: eager launches from Python cost ~25 us each, more than the small kernels
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(2):
                    fn()
                torch.cuda.synchronize()
                gph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gph, stream=side):
                    for _ in range(10):
                        fn()
                gph.replay()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(side)
                gph.replay()
                e1.record(side)
                torch.cuda.synchronize()
            res[name] = e0.elapsed_time(e1) / 10
        row.append(res)
    gf = 2.0 * N * (H // s) * (W // s) * Ci * Co * k * k / 1e9
    print(f'{N}x{H}x{W} {Ci}->{Co} k{k} s{s}: {gf:6.1f} GF | tc5 fwd {row[0]["fwd"]*1e3:7.0f} us ({gf/row[0]["fwd"]:6.1f} TF) dgrad {row[0]["dgrad"]*1e3:7.0f} us ({gf/row[0]["dgrad"]:6.1f} TF)'
          f' | mma fwd {row[1]["fwd"]*1e3:7.0f} us ({gf/row[1]["fwd"]:6.1f} TF) dgrad {row[1]["dgrad"]*1e3:7.0f} us | wgrad tc5 {row[0]["wgrad"]*1e3:7.0f} us ({gf/row[0]["wgrad"]:6.1f} TF) mma {row[1]["wgrad"]*1e3:7.0f} us ({gf/row[1]["wgrad"]:6.1f} TF)', flush=True)
