"""Unit parity of single kernels through the C ABI against plain PyTorch fp32 math on the same
bf16-rounded inputs (tolerance: fp32 accumulation-order noise + one bf16 output rounding)."""
import ctypes as C

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from rsgnet_b200 import _engine, _lib
from rsgnet_b200._engine import PlanBuilder, View

pytestmark = pytest.mark.gpu
# No environment switches: engine=0 cases follow the PRODUCTION routing (small maps -> conv_mma / conv_ws), engine=2
# forces the tcgen05 kernel on any map size, engine=1 forces the mma.sync kernel, engine=3 the weight-streaming one.


def _run(pb, n):
    pb.allocate('cuda')
    handle = C.c_void_p()
    _lib.check(_lib.lib().rsg_plan_create(C.byref(handle), pb.chunk))
    _engine.emit(pb, handle)
    return handle


def _exec(handle, n, ext=None):
    arr = (C.c_void_p * _engine.N_EXT)()
    for k, t in (ext or {}).items():
        arr[k] = t.data_ptr()
    _lib.check(_lib.lib().rsg_plan_run(handle, _lib.stream_ptr(), arr, _engine.N_EXT, n, n, 1, 0))
    torch.cuda.synchronize()


CASES = [
    # Cin, Cout, k, stride, H, W, relu, residual
    (32, 32, 3, 1, 16, 12, True, True),
    (64, 64, 3, 1, 8, 6, True, False),
    (256, 32, 3, 1, 16, 12, True, False),
    (32, 64, 3, 2, 16, 12, False, False),
    (64, 256, 1, 1, 16, 12, True, True),
    (128, 32, 1, 1, 4, 3, False, False),
    (96, 96, 3, 1, 16, 12, True, False),
    (48, 48, 3, 1, 12, 10, True, True),
    (24, 32, 3, 1, 16, 12, True, False),
    (16, 16, 3, 1, 24, 16, True, True),
    (256, 256, 3, 1, 8, 6, True, True),
    (144, 48, 3, 1, 10, 8, True, False),
    (192, 192, 3, 1, 24, 18, True, True),      # W48 branch 2: weight-streaming kernel with 96-column slices
    (256, 64, 3, 2, 32, 24, True, False),      # transition1: stride 2, weights streamed with the halo stages
    (256, 96, 3, 2, 26, 20, False, True),      # W48 transition1 (ragged map)
]


@pytest.mark.parametrize('cin,cout,k,stride,H,W,relu,use_res', CASES)
def test_conv_vs_torch(cin, cout, k, stride, H, W, relu, use_res):
    N = 5
    g = torch.Generator().manual_seed(cin * 131 + cout)
    x = torch.randn(N, cin, H, W, generator=g).bfloat16().float()
    w = (torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5).bfloat16().float()
    b = torch.randn(cout, generator=g) * 0.1
    Ho, Wo = (H - 1) // stride + 1, (W - 1) // stride + 1
    r = torch.randn(N, cout, Ho, Wo, generator=g).bfloat16().float()
    pb = PlanBuilder(8, reuse=False)
    xin = pb.buf('x', H, W, cin + 8)           # channel-sliced input (offset 8)
    rb = pb.buf('r', Ho, Wo, cout)
    ob = pb.buf('o', Ho, Wo, cout + 16)        # channel-sliced output (offset 8)
    pb.conv(View(xin, 8, cin), w.double().numpy(), b.double().numpy(), stride=stride, relu=relu,
            dst=View(ob, 8, cout), res=[(View(rb), 0)] if use_res else ())
    h = _run(pb, N)
    pb.tensor_of(xin)[:N, ..., 8:] = x.permute(0, 2, 3, 1).cuda().bfloat16()
    pb.tensor_of(rb)[:N] = r.permute(0, 2, 3, 1).cuda().bfloat16()
    pb.tensor_of(ob).fill_(7.0)
    _exec(h, N)
    ref = F.conv2d(x, w, b, stride, k // 2)
    if use_res:
        ref = ref + r
    if relu:
        ref = F.relu(ref)
    got = pb.tensor_of(ob)[:N, ..., 8:8 + cout].float().permute(0, 3, 1, 2).cpu()
    err = (got - ref).abs().max().item()
    assert err <= 2e-2 * max(ref.abs().max().item(), 1.0), err
    # the untouched channel slices keep their sentinel
    assert torch.all(pb.tensor_of(ob)[:N, ..., :8] == 7.0) and torch.all(pb.tensor_of(ob)[:N, ..., 8 + cout:] == 7.0)
    _lib.lib().rsg_plan_destroy(h)


WS_CASES = [
    # Cin, Cout, k, H, W, N, nres   (engine=3: the weight-streaming tcgen05 kernel, conv_ws.cu)
    (128, 128, 3, 16, 12, 5, 1),      # 2 images / supertile, ragged last supertile
    (256, 256, 3, 8, 6, 5, 1),        # 8 images / supertile (3 of them out of bounds), 2 N-slices
    (256, 256, 3, 8, 6, 21, 0),
    (128, 256, 3, 8, 6, 16, 2),
    (384, 384, 3, 12, 9, 7, 1),       # W48 branch 3
    (128, 128, 3, 16, 12, 700, 1),    # 350 supertiles > 148 CTAs: several per CTA (barrier phases wrap)
    (1152, 128, 1, 8, 6, 9, 0),       # 1x1 through the same path
]


@pytest.mark.parametrize('cin,cout,k,H,W,N,nres', WS_CASES)
def test_conv_ws_vs_torch(cin, cout, k, H, W, N, nres):
    g = torch.Generator().manual_seed(cin * 7 + cout + N)
    x = torch.randn(N, cin, H, W, generator=g).bfloat16().float()
    w = (torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5).bfloat16().float()
    b = torch.randn(cout, generator=g) * 0.1
    rs = [torch.randn(N, cout, H, W, generator=g).bfloat16().float() for _ in range(nres)]
    pb = PlanBuilder(N, reuse=False)
    xin = pb.buf('x', H, W, cin + 8)
    rbs = [pb.buf(f'r{i}', H, W, cout) for i in range(nres)]
    ob = pb.buf('o', H, W, cout + 16)
    pb.conv(View(xin, 8, cin), w.double().numpy(), b.double().numpy(), relu=True, dst=View(ob, 8, cout),
            res=[(View(rb), 0) for rb in rbs], engine=3)
    assert pb.ops[-1][1]['engine'] == 3
    h = _run(pb, N)
    pb.tensor_of(xin)[:N, ..., 8:] = x.permute(0, 2, 3, 1).cuda().bfloat16()
    pb.tensor_of(xin)[:N, ..., :8] = 1e4        # a neighbouring channel slice must not leak in
    for rb, r in zip(rbs, rs):
        pb.tensor_of(rb)[:N] = r.permute(0, 2, 3, 1).cuda().bfloat16()
    pb.tensor_of(ob).fill_(7.0)
    _exec(h, N)
    ref = F.conv2d(x.cuda(), w.cuda(), b.cuda(), 1, k // 2)
    for r in rs:
        ref = ref + r.cuda()
    ref = F.relu(ref).cpu()
    got = pb.tensor_of(ob)[:N, ..., 8:8 + cout].float().permute(0, 3, 1, 2).cpu()
    err = (got - ref).abs().max().item()
    assert err <= 2e-2 * max(ref.abs().max().item(), 1.0), err
    assert torch.all(pb.tensor_of(ob)[:N, ..., :8] == 7.0) and torch.all(pb.tensor_of(ob)[:N, ..., 8 + cout:] == 7.0)
    _lib.lib().rsg_plan_destroy(h)


MMA_CASES = [
    # Cin, Cout, k, stride, Hin, Win, nres, shifts  (engine=1: every shape the production plan of RSGNet-W32 sends to
    # the mma.sync kernel -- the 8x6-output layers of stage 3/4 and transition3)
    (128, 256, 3, 2, 16, 12, 3, (0, 0, 0)),     # stage4 fuse_layers.3.2 + three residual terms
    (128, 256, 3, 2, 16, 12, 0, ()),            # transition3
    (64, 256, 3, 2, 16, 12, 0, ()),             # stage4 fuse_layers.3.1.1
    (32, 256, 3, 2, 16, 12, 0, ()),             # stage4 fuse_layers.3.0.2
    (256, 32, 1, 1, 8, 6, 0, ()),               # stage4 fuse_layers.0.3
    (256, 64, 1, 1, 8, 6, 0, ()),               # stage4 fuse_layers.1.3
    (256, 128, 1, 1, 8, 6, 0, ()),              # stage4 fuse_layers.2.3
    (192, 384, 3, 2, 24, 18, 3, (0, 0, 0)),     # W48 stage4 fuse_layers.3.2
    (384, 48, 1, 1, 12, 9, 0, ()),              # W48 stage4 fuse_layers.0.3
    (24, 32, 3, 1, 16, 12, 0, ()),              # odd channel count (Cin % 16 != 0)
]


@pytest.mark.parametrize('cin,cout,k,stride,H,W,nres,shifts', MMA_CASES)
def test_conv_mma_production_shapes_vs_torch(cin, cout, k, stride, H, W, nres, shifts):
    N = 7
    g = torch.Generator().manual_seed(cin * 5 + cout + k + nres)
    x = torch.randn(N, cin, H, W, generator=g).bfloat16().float()
    w = (torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5).bfloat16().float()
    b = torch.randn(cout, generator=g) * 0.1
    Ho, Wo = (H - 1) // stride + 1, (W - 1) // stride + 1
    pb = PlanBuilder(N, reuse=False)
    xin = pb.buf('x', H, W, cin + 8)
    res, rts = [], []
    for i in range(nres):
        rb = pb.buf(f'r{i}', Ho >> shifts[i], Wo >> shifts[i], cout)
        res.append((View(rb), shifts[i]))
        rts.append((rb, torch.randn(N, cout, Ho >> shifts[i], Wo >> shifts[i], generator=g).bfloat16().float(), shifts[i]))
    ob = pb.buf('o', Ho, Wo, cout + 16)
    pb.conv(View(xin, 8, cin), w.double().numpy(), b.double().numpy(), stride=stride, relu=True,
            dst=View(ob, 8, cout), res=res, engine=1)
    h = _run(pb, N)
    pb.tensor_of(xin)[:N, ..., 8:] = x.permute(0, 2, 3, 1).cuda().bfloat16()
    pb.tensor_of(xin)[:N, ..., :8] = 1e4
    for rb, r, _ in rts:
        pb.tensor_of(rb)[:N] = r.permute(0, 2, 3, 1).cuda().bfloat16()
    pb.tensor_of(ob).fill_(7.0)
    _exec(h, N)
    ref = F.conv2d(x.cuda(), w.cuda(), b.cuda(), stride, k // 2)
    for _, r, sh in rts:
        r = r.cuda()
        ref = ref + (F.interpolate(r, scale_factor=2 ** sh, mode='nearest') if sh else r)
    ref = F.relu(ref).cpu()
    got = pb.tensor_of(ob)[:N, ..., 8:8 + cout].float().permute(0, 3, 1, 2).cpu()
    err = (got - ref).abs().max().item()
    assert err <= 2e-2 * max(ref.abs().max().item(), 1.0), err
    assert torch.all(pb.tensor_of(ob)[:N, ..., :8] == 7.0) and torch.all(pb.tensor_of(ob)[:N, ..., 8 + cout:] == 7.0)
    _lib.lib().rsg_plan_destroy(h)


WS2_CASES = [
    # Cin, Cout, k, H, W, N, nres   (engine=4: the CTA-pair weight-streaming kernel, conv_ws2.cu, tcgen05.mma.cta_group::2)
    (128, 128, 3, 16, 12, 5, 1),      # one image per CTA; the last pair-unit has an out-of-bounds image
    (128, 128, 3, 16, 12, 1, 0),      # a single image: the peer CTA computes nothing but must keep the protocol
    (256, 256, 3, 8, 6, 21, 1),       # four images per CTA, two N slices
    (128, 256, 3, 8, 6, 16, 2),
    (128, 128, 3, 16, 12, 700, 1),    # 350 pair-units > 74 pairs: several per pair, both accumulator buffers, phases wrap
    (256, 256, 3, 8, 6, 1200, 0),
    (1152, 128, 1, 8, 6, 9, 0),       # 1x1 through the same path (72 K chunks)
]


@pytest.mark.parametrize('cin,cout,k,H,W,N,nres', WS2_CASES)
def test_conv_ws2_cta_pair_vs_torch(cin, cout, k, H, W, N, nres):
    g = torch.Generator().manual_seed(cin * 11 + cout + N)
    x = torch.randn(N, cin, H, W, generator=g).bfloat16().float()
    w = (torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5).bfloat16().float()
    b = torch.randn(cout, generator=g) * 0.1
    rs = [torch.randn(N, cout, H, W, generator=g).bfloat16().float() for _ in range(nres)]
    pb = PlanBuilder(N, reuse=False)
    xin = pb.buf('x', H, W, cin + 8)
    rbs = [pb.buf(f'r{i}', H, W, cout) for i in range(nres)]
    ob = pb.buf('o', H, W, cout + 16)
    pb.conv(View(xin, 8, cin), w.double().numpy(), b.double().numpy(), relu=True, dst=View(ob, 8, cout),
            res=[(View(rb), 0) for rb in rbs], engine=4)
    assert pb.ops[-1][1]['engine'] == 4
    h = _run(pb, N)
    pb.tensor_of(xin)[:N, ..., 8:] = x.permute(0, 2, 3, 1).cuda().bfloat16()
    pb.tensor_of(xin)[:N, ..., :8] = 1e4
    for rb, r in zip(rbs, rs):
        pb.tensor_of(rb)[:N] = r.permute(0, 2, 3, 1).cuda().bfloat16()
    pb.tensor_of(ob).fill_(7.0)
    for _ in range(2):                       # twice: barrier state / TMEM are per launch
        _exec(h, N)
    ref = F.conv2d(x.cuda(), w.cuda(), b.cuda(), 1, k // 2)
    for r in rs:
        ref = ref + r.cuda()
    ref = F.relu(ref).cpu()
    got = pb.tensor_of(ob)[:N, ..., 8:8 + cout].float().permute(0, 3, 1, 2).cpu()
    err = (got - ref).abs().max().item()
    assert err <= 2e-2 * max(ref.abs().max().item(), 1.0), err
    assert torch.all(pb.tensor_of(ob)[:N, ..., :8] == 7.0) and torch.all(pb.tensor_of(ob)[:N, ..., 8 + cout:] == 7.0)
    _lib.lib().rsg_plan_destroy(h)


WS_S2_CASES = [
    # Cin, Cout, Hin, Win, N, nres, res_shift   (engine=3, stride 2: the flat stride-2 formulation of conv_ws.cu, four phase-plane sets)
    (128, 256, 16, 12, 7, 3, 0),      # stage4 fuse_layers.3.2 + three residual terms -> 8x6
    (128, 256, 16, 12, 9, 0, 0),      # transition3
    (64, 256, 16, 12, 5, 0, 0),
    (32, 256, 16, 12, 5, 0, 0),
    (64, 128, 32, 24, 5, 2, 1),       # stage3/4 fuse_layers.2.1 + identity + up-sampled term -> 16x12
    (32, 128, 32, 24, 300, 0, 0),     # more supertiles than CTAs
    (64, 64, 32, 24, 3, 0, 0),        # NS = 64
    (64, 128, 13, 11, 3, 1, 0),       # odd input size: Hout = 7, Wout = 6
    (192, 384, 24, 18, 4, 3, 0),      # W48 stage4 fuse_layers.3.2
]


@pytest.mark.parametrize('cin,cout,H,W,N,nres,shift', WS_S2_CASES)
def test_conv_ws_stride2_vs_torch(cin, cout, H, W, N, nres, shift):
    g = torch.Generator().manual_seed(cin + 5 * cout + H + N)
    x = torch.randn(N, cin, H, W, generator=g).bfloat16().float()
    w = (torch.randn(cout, cin, 3, 3, generator=g) / (cin * 9) ** 0.5).bfloat16().float()
    b = torch.randn(cout, generator=g) * 0.1
    Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    pb = PlanBuilder(N, reuse=False)
    xin = pb.buf('x', H, W, cin + 8)
    res, rts = [], []
    for i in range(nres):
        sh = shift if i == nres - 1 else 0
        rh, rw = (Ho >> sh, Wo >> sh) if sh else (Ho, Wo)
        rb = pb.buf(f'r{i}', rh, rw, cout)
        res.append((View(rb), sh))
        rts.append((rb, torch.randn(N, cout, rh, rw, generator=g).bfloat16().float(), sh))
    ob = pb.buf('o', Ho, Wo, cout + 8)
    pb.conv(View(xin, 8, cin), w.double().numpy(), b.double().numpy(), stride=2, relu=True,
            dst=View(ob, 0, cout), res=res, engine=3)
    assert pb.ops[-1][1]['engine'] == 3
    h = _run(pb, N)
    pb.tensor_of(xin)[:N, ..., 8:] = x.permute(0, 2, 3, 1).cuda().bfloat16()
    pb.tensor_of(xin)[:N, ..., :8] = 1e4
    for rb, r, _ in rts:
        pb.tensor_of(rb)[:N] = r.permute(0, 2, 3, 1).cuda().bfloat16()
    pb.tensor_of(ob).fill_(7.0)
    _exec(h, N)
    ref = F.conv2d(x.cuda(), w.cuda(), b.cuda(), 2, 1)
    for _, r, sh in rts:
        r = r.cuda()
        ref = ref + (F.interpolate(r, scale_factor=2 ** sh, mode='nearest') if sh else r)
    ref = F.relu(ref).cpu()
    got = pb.tensor_of(ob)[:N, ..., :cout].float().permute(0, 3, 1, 2).cpu()
    err = (got - ref).abs().max().item()
    assert err <= 2e-2 * max(ref.abs().max().item(), 1.0), err
    assert torch.all(pb.tensor_of(ob)[:N, ..., cout:] == 7.0)
    _lib.lib().rsg_plan_destroy(h)


S2_CASES = [
    # Cin, Cout, Hin, Win, N, nres, res_shift   (engine=2: stride-2 3x3 on the tcgen05 kernel, four TMA phase patches)
    (64, 64, 32, 24, 3, 0, 0),
    (32, 64, 32, 24, 5, 2, 1),        # fuse layer: + identity term + nearest-upsampled term
    (256, 64, 16, 12, 4, 0, 0),       # transition: many input channels, sliced weights
    (64, 128, 13, 11, 3, 1, 0),       # odd input size: Hout = 7, Wout = 6
    (32, 32, 64, 48, 40, 0, 0),       # more tiles than SMs
    (128, 256, 16, 12, 6, 3, 0),
]


@pytest.mark.parametrize('cin,cout,H,W,N,nres,shift', S2_CASES)
def test_conv_stride2_tcgen05_vs_torch(cin, cout, H, W, N, nres, shift):
    g = torch.Generator().manual_seed(cin + 3 * cout + H)
    x = torch.randn(N, cin, H, W, generator=g).bfloat16().float()
    w = (torch.randn(cout, cin, 3, 3, generator=g) / (cin * 9) ** 0.5).bfloat16().float()
    b = torch.randn(cout, generator=g) * 0.1
    Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    pb = PlanBuilder(N, reuse=False)
    xin = pb.buf('x', H, W, cin + 8)
    res, rts = [], []
    for i in range(nres):
        sh = shift if i == nres - 1 else 0
        rh, rw = (Ho >> sh, Wo >> sh) if sh else (Ho, Wo)
        rb = pb.buf(f'r{i}', rh, rw, cout)
        res.append((View(rb), sh))
        rts.append((rb, torch.randn(N, cout, rh, rw, generator=g).bfloat16().float(), sh))
    ob = pb.buf('o', Ho, Wo, cout + 8)
    pb.conv(View(xin, 8, cin), w.double().numpy(), b.double().numpy(), stride=2, relu=True,
            dst=View(ob, 0, cout), res=res, engine=2)
    h = _run(pb, N)
    pb.tensor_of(xin)[:N, ..., 8:] = x.permute(0, 2, 3, 1).cuda().bfloat16()
    pb.tensor_of(xin)[:N, ..., :8] = 1e4
    for rb, r, _ in rts:
        pb.tensor_of(rb)[:N] = r.permute(0, 2, 3, 1).cuda().bfloat16()
    pb.tensor_of(ob).fill_(7.0)
    _exec(h, N)
    ref = F.conv2d(x.cuda(), w.cuda(), b.cuda(), 2, 1)
    for _, r, sh in rts:
        r = r.cuda()
        ref = ref + (F.interpolate(r, scale_factor=2 ** sh, mode='nearest') if sh else r)
    ref = F.relu(ref).cpu()
    got = pb.tensor_of(ob)[:N, ..., :cout].float().permute(0, 3, 1, 2).cpu()
    err = (got - ref).abs().max().item()
    assert err <= 2e-2 * max(ref.abs().max().item(), 1.0), err
    assert torch.all(pb.tensor_of(ob)[:N, ..., cout:] == 7.0)
    _lib.lib().rsg_plan_destroy(h)


@pytest.mark.parametrize('C_,H,W,N', [(32, 16, 12, 3), (32, 64, 48, 5), (16, 8, 6, 4), (48, 12, 9, 2)])
def test_deconv4_fused_pixel_shuffle_vs_torch(C_, H, W, N):
    """ConvTranspose2d(4, 2, 1) + BN + ReLU (pose_rsgnet.py:733-744) as one 3x3 conv with a pixel-shuffle
    epilogue, against F.conv_transpose2d."""
    g = torch.Generator().manual_seed(C_ + H)
    x = torch.randn(N, C_, H, W, generator=g).bfloat16().float()
    wt = (torch.randn(C_, C_, 4, 4, generator=g) / (C_ * 4) ** 0.5).bfloat16().float()
    sd = {'d.0.weight': wt, 'd.1.weight': torch.ones(C_), 'd.1.bias': torch.randn(C_, generator=g) * 0.1,
          'd.1.running_mean': torch.zeros(C_), 'd.1.running_var': torch.ones(C_) - _engine.EPS}
    pb = PlanBuilder(N, reuse=False)
    xin = pb.buf('x', H, W, C_)
    ob = pb.buf('o', 2 * H, 2 * W, C_)
    _engine._deconv4(pb, _engine._Params(sd), 'd', View(xin), View(ob))
    assert len(pb.ops) == 1 and pb.ops[0][1]['psc'] == C_     # the fused form was chosen
    h = _run(pb, N)
    pb.tensor_of(xin)[:N] = x.permute(0, 2, 3, 1).cuda().bfloat16()
    pb.tensor_of(ob).fill_(7.0)
    _exec(h, N)
    ref = F.relu(F.conv_transpose2d(x, wt, None, 2, 1) + sd['d.1.bias'][None, :, None, None])
    got = pb.tensor_of(ob)[:N].float().permute(0, 3, 1, 2).cpu()
    err = (got - ref).abs().max().item()
    assert err <= 2e-2 * max(ref.abs().max().item(), 1.0), err
    _lib.lib().rsg_plan_destroy(h)


@pytest.mark.parametrize('C_,H,W,N', [(32, 16, 8, 2), (32, 64, 48, 7), (48, 24, 18, 3), (32, 21, 13, 3), (32, 32, 24, 200),
                                      (32, 19, 20, 5), (32, 16, 16, 1), (32, 64, 48, 300)])
def test_fused_basic_block_vs_torch(C_, H, W, N):
    """BasicBlock (pose_rsgnet.py:25-54) as one kernel: relu(bn2(conv2(relu(bn1(conv1(x))))) + x); the fused kernel
    rounds the intermediate to bf16 exactly like the two-conv path does."""
    g = torch.Generator().manual_seed(C_ * 3 + H + N)
    x = torch.randn(N, C_, H, W, generator=g).bfloat16().float()
    sd = {}
    for i in (1, 2):
        sd[f'conv{i}.weight'] = (torch.randn(C_, C_, 3, 3, generator=g) / (C_ * 9) ** 0.5)
        sd[f'bn{i}.weight'] = torch.rand(C_, generator=g) + 0.5
        sd[f'bn{i}.bias'] = torch.randn(C_, generator=g) * 0.1
        sd[f'bn{i}.running_mean'] = torch.randn(C_, generator=g) * 0.1
        sd[f'bn{i}.running_var'] = torch.rand(C_, generator=g) + 0.5
    pb = PlanBuilder(N, reuse=False)
    xin = pb.buf('x', H, W, C_)
    out = _engine._basic(pb, _engine._Params(sd), View(xin))
    assert [k for k, _, _ in pb.ops] == ['bblock']          # the fused form was chosen
    h = _run(pb, N)
    pb.tensor_of(xin)[:N] = x.permute(0, 2, 3, 1).cuda().bfloat16()
    pb.tensor_of(out.buf).fill_(7.0)
    _exec(h, N)

    def bn(v, i):
        return F.batch_norm(v, sd[f'bn{i}.running_mean'].cuda(), sd[f'bn{i}.running_var'].cuda(), sd[f'bn{i}.weight'].cuda(),
                            sd[f'bn{i}.bias'].cuda(), False, 0.0, _engine.EPS)
    xc = x.cuda()
    y = F.relu(bn(F.conv2d(xc, sd['conv1.weight'].cuda(), None, 1, 1), 1))
    ref = F.relu(bn(F.conv2d(y, sd['conv2.weight'].cuda(), None, 1, 1), 2) + xc).cpu()
    got = pb.tensor_of(out.buf)[:N].float().permute(0, 3, 1, 2).cpu()
    err = (got - ref).abs().max().item()
    assert err <= 2e-2 * max(ref.abs().max().item(), 1.0), err
    _lib.lib().rsg_plan_destroy(h)


@pytest.mark.parametrize('cin,H,W,N', [(256, 16, 8, 2), (64, 16, 8, 1), (256, 64, 48, 5), (64, 64, 48, 3), (256, 21, 13, 3),
                                       (128, 19, 20, 2), (256, 32, 24, 150), (64, 64, 48, 40)])
def test_fused_bottleneck_vs_torch(cin, H, W, N):
    """Bottleneck (pose_rsgnet.py:57-95) as one kernel: relu(bn3(conv3(relu(bn2(conv2(relu(bn1(conv1(x)))))))) + r) with
    r = x (Cin = 256) or bn(downsample(x)); both intermediates are rounded to bf16 exactly like the three-conv path."""
    g = torch.Generator().manual_seed(cin + H * 7 + N)
    x = torch.randn(N, cin, H, W, generator=g).bfloat16().float()
    sd = {}
    shapes = {'conv1': (64, cin, 1, 1), 'conv2': (64, 64, 3, 3), 'conv3': (256, 64, 1, 1)}
    bns = {'conv1': 'bn1', 'conv2': 'bn2', 'conv3': 'bn3'}
    if cin != 256:
        shapes['downsample.0'] = (256, cin, 1, 1)
        bns['downsample.0'] = 'downsample.1'
    for name, shp in shapes.items():
        sd[f'{name}.weight'] = torch.randn(*shp, generator=g) / (shp[1] * shp[2] * shp[3]) ** 0.5
        bn = bns[name]
        sd[f'{bn}.weight'] = torch.rand(shp[0], generator=g) + 0.5
        sd[f'{bn}.bias'] = torch.randn(shp[0], generator=g) * 0.1
        sd[f'{bn}.running_mean'] = torch.randn(shp[0], generator=g) * 0.1
        sd[f'{bn}.running_var'] = torch.rand(shp[0], generator=g) + 0.5
    pb = PlanBuilder(N, reuse=False)
    xin = pb.buf('x', H, W, cin)
    out = _engine._bottleneck(pb, _engine._Params(sd), View(xin))
    assert [k for k, _, _ in pb.ops][-1] == 'bneck'            # the fused form was chosen
    h = _run(pb, N)
    pb.tensor_of(xin)[:N] = x.permute(0, 2, 3, 1).cuda().bfloat16()
    pb.tensor_of(out.buf).fill_(7.0)
    _exec(h, N)

    def cbn(v, name, pad=0):
        bn = bns[name]
        return F.batch_norm(F.conv2d(v, sd[f'{name}.weight'].cuda(), None, 1, pad), sd[f'{bn}.running_mean'].cuda(),
                            sd[f'{bn}.running_var'].cuda(), sd[f'{bn}.weight'].cuda(), sd[f'{bn}.bias'].cuda(), False, 0.0,
                            _engine.EPS)
    xc = x.cuda()
    y = F.relu(cbn(xc, 'conv1'))
    y = F.relu(cbn(y, 'conv2', 1))
    r = xc if cin == 256 else cbn(xc, 'downsample.0')
    ref = F.relu(cbn(y, 'conv3') + r).cpu()
    got = pb.tensor_of(out.buf)[:N].float().permute(0, 3, 1, 2).cpu()
    err = (got - ref).abs().max().item()
    assert err <= 2e-2 * max(ref.abs().max().item(), 1.0), err
    _lib.lib().rsg_plan_destroy(h)


def test_conv_fp32_nchw_output_and_upsampled_residuals():
    N, cin, K, H, W = 3, 32, 17, 16, 12
    g = torch.Generator().manual_seed(1)
    x = torch.randn(N, cin, H, W, generator=g).bfloat16().float()
    w = (torch.randn(K, cin, 1, 1, generator=g) / cin ** 0.5).bfloat16().float()
    b = torch.randn(K, generator=g)
    pb = PlanBuilder(4, reuse=False)
    xin = pb.buf('x', H, W, cin)
    of = pb.buf('of', H, W, K, itemsize=4)
    pb.conv(View(xin), w.double().numpy(), b.double().numpy(), out_f32=of)
    # second conv: 3x3 s2 with two residual terms, one of them nearest-upsampled by 2
    w2 = (torch.randn(32, cin, 3, 3, generator=g) / (cin * 9) ** 0.5).bfloat16().float()
    r0 = pb.buf('r0', H // 2, W // 2, 32)
    r1 = pb.buf('r1', H // 4, W // 4, 32)
    o2 = pb.buf('o2', H // 2, W // 2, 32)
    pb.conv(View(xin), w2.double().numpy(), np.zeros(32), stride=2, relu=True, dst=View(o2),
            res=[(View(r0), 0), (View(r1), 1)])
    h = _run(pb, N)
    pb.tensor_of(xin)[:N] = x.permute(0, 2, 3, 1).cuda().bfloat16()
    a0 = torch.randn(N, 32, H // 2, W // 2, generator=g).bfloat16().float()
    a1 = torch.randn(N, 32, H // 4, W // 4, generator=g).bfloat16().float()
    pb.tensor_of(r0)[:N] = a0.permute(0, 2, 3, 1).cuda().bfloat16()
    pb.tensor_of(r1)[:N] = a1.permute(0, 2, 3, 1).cuda().bfloat16()
    _exec(h, N)
    ref = F.conv2d(x, w, b)
    got = pb.tensor_of(of)[:N].cpu()
    assert (got - ref).abs().max().item() <= 1e-4 * max(1.0, ref.abs().max().item())
    ref2 = F.relu(F.conv2d(x, w2, None, 2, 1) + a0 + F.interpolate(a1, scale_factor=2, mode='nearest'))
    got2 = pb.tensor_of(o2)[:N].float().permute(0, 3, 1, 2).cpu()
    assert (got2 - ref2).abs().max().item() <= 2e-2 * max(1.0, ref2.abs().max().item())
    _lib.lib().rsg_plan_destroy(h)


@pytest.mark.parametrize('C_,S_hw', [(32, (16, 12)), (16, (12, 8)), (48, (9, 7)), (64, (8, 8)), (32, (64, 48)),
                                     (48, (24, 18)), (64, (20, 13))])
def test_trp_attention_vs_torch(C_, S_hw):
    N = 3
    H, W = S_hw
    S = H * W
    g = torch.Generator().manual_seed(C_)
    x = (torch.randn(N, S, C_, generator=g) * 0.5).bfloat16().float()
    gv = torch.randn(N, S, C_, generator=g).bfloat16().float()
    pb = PlanBuilder(4, reuse=False)
    xb = pb.buf('x', H, W, 2 * C_)
    gb = pb.buf('g', H, W, C_)
    yb = pb.buf('y', H, W, C_)
    rel = pb.buf('rel', 1, 1, S * S, itemsize=4)
    pb.simple('attention', dict(x=View(xb, C_, C_), g=View(gb), y=View(yb)), [xb, gb], [yb])
    pb.simple('relscores', dict(x=View(xb, C_, C_), out=rel), [xb], [rel])
    h = _run(pb, N)
    pb.tensor_of(xb)[:N, ..., C_:] = x.view(N, H, W, C_).cuda().bfloat16()
    pb.tensor_of(gb)[:N] = gv.view(N, H, W, C_).cuda().bfloat16()
    _exec(h, N)
    A = torch.sigmoid(x @ x.transpose(1, 2))
    ref = A @ gv
    got = pb.tensor_of(yb)[:N].float().view(N, S, C_).cpu()
    assert (got - ref).abs().max().item() <= 2e-2 * ref.abs().max().item()
    got_rel = pb.tensor_of(rel)[:N].view(N, S, S).cpu()
    assert (got_rel - A).abs().max().item() <= 1e-5
    _lib.lib().rsg_plan_destroy(h)


def test_groupnorm_maxpool_bilinear_fuse_vs_torch():
    N, C_, H, W = 3, 32, 12, 8
    g = torch.Generator().manual_seed(3)
    x = (torch.randn(N, C_, H, W, generator=g) * 2 + 0.7).bfloat16().float()
    gamma, beta = torch.randn(C_, generator=g), torch.randn(C_, generator=g)
    pb = PlanBuilder(4, reuse=False)
    xb = pb.buf('x', H, W, C_)
    yb = pb.buf('y', H, W, 2 * C_)
    pool = pb.buf('p', H // 2, W // 2, C_)
    f32 = pb.buf('f', H, W, 5, itemsize=4)
    up = pb.buf('u', 2 * H, 2 * W, 5, itemsize=4)
    fz = pb.buf('fz', H, W, C_)
    pb.simple('groupnorm', dict(x=View(xb), y=View(yb, C_, C_), groups=8,
                                gamma=pb.const(gamma.numpy()), beta=pb.const(beta.numpy())), [xb], [yb])
    pb.simple('maxpool', dict(src=View(xb), dst=pool), [xb], [pool])
    pb.simple('bilinear', dict(src=f32, out=up, C=5, H=H, W=W, sigmoid=True), [f32], [up])
    pb.simple('fuse', dict(terms=[(View(xb), 0), (View(pool), 1)], dst=View(fz), relu=True), [xb, pool], [fz])
    h = _run(pb, N)
    pb.tensor_of(xb)[:N] = x.permute(0, 2, 3, 1).cuda().bfloat16()
    src = torch.randn(N, 5, H, W, generator=g)
    pb.tensor_of(f32)[:N] = src.cuda()
    _exec(h, N)
    ref = F.group_norm(x, 8, gamma, beta, 1e-5)
    got = pb.tensor_of(yb)[:N, ..., C_:].float().permute(0, 3, 1, 2).cpu()
    assert (got - ref).abs().max().item() <= 2e-2 * ref.abs().max().item()
    refp = F.max_pool2d(x, 2)
    assert torch.equal(pb.tensor_of(pool)[:N].float().permute(0, 3, 1, 2).cpu(), refp)
    refu = torch.sigmoid(F.interpolate(src, scale_factor=2, mode='bilinear', align_corners=True))
    assert (pb.tensor_of(up)[:N].cpu() - refu).abs().max().item() <= 1e-5
    reff = F.relu(x + F.interpolate(refp, scale_factor=2, mode='nearest')).bfloat16().float()
    assert torch.equal(pb.tensor_of(fz)[:N].float().permute(0, 3, 1, 2).cpu(), reff)
    _lib.lib().rsg_plan_destroy(h)


TC5_CASES = [
    # Cin, Cout, k, H, W, relu, nres
    (32, 32, 3, 16, 8, False, 0),      # exactly one tile
    (32, 32, 3, 64, 48, True, 1),
    (64, 64, 3, 32, 24, True, 1),
    (32, 32, 1, 16, 8, False, 0),
    (64, 32, 3, 64, 48, True, 0),
    (96, 32, 3, 24, 16, True, 0),
    (16, 16, 3, 24, 16, True, 1),
    (32, 32, 3, 20, 12, True, 1),      # partial tiles in both directions
    (64, 256, 1, 16, 12, True, 1),
    (128, 64, 1, 16, 12, True, 0),
    (128, 128, 3, 16, 12, True, 1),
    (256, 256, 3, 16, 8, True, 1),
    (256, 32, 3, 32, 24, True, 0),
    (256, 64, 1, 16, 12, True, 0),
    (96, 96, 3, 16, 16, True, 0),
    (64, 64, 3, 64, 48, True, 2),
    (192, 192, 3, 24, 18, True, 1),
    (48, 48, 3, 24, 24, True, 1),
]


@pytest.mark.parametrize('cin,cout,k,H,W,relu,nres', TC5_CASES)
def test_conv_tcgen05_vs_torch(cin, cout, k, H, W, relu, nres):
    """engine=2 forces the tcgen05/TMEM/TMA kernel (fails loudly if the shape is not covered)."""
    N = 5
    g = torch.Generator().manual_seed(cin * 7 + cout + k)
    x = torch.randn(N, cin, H, W, generator=g).bfloat16().float()
    w = (torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5).bfloat16().float()
    b = torch.randn(cout, generator=g) * 0.1
    r = torch.randn(N, cout, H, W, generator=g).bfloat16().float()
    pb = PlanBuilder(8, reuse=False)
    xin = pb.buf('x', H, W, cin + 8)
    rb = pb.buf('r', H, W, cout)
    ob = pb.buf('o', H, W, cout + 16)
    pb.conv(View(xin, 8, cin), w.double().numpy(), b.double().numpy(), relu=relu, dst=View(ob, 8, cout),
            res=[(View(rb), 0)] * nres, engine=2)
    h = _run(pb, N)
    pb.tensor_of(xin)[:N, ..., 8:] = x.permute(0, 2, 3, 1).cuda().bfloat16()
    pb.tensor_of(rb)[:N] = r.permute(0, 2, 3, 1).cuda().bfloat16()
    pb.tensor_of(ob).fill_(7.0)
    _exec(h, N)
    ref = F.conv2d(x, w, b, 1, k // 2) + nres * r
    if relu:
        ref = F.relu(ref)
    got = pb.tensor_of(ob)[:N, ..., 8:8 + cout].float().permute(0, 3, 1, 2).cpu()
    err = (got - ref).abs().max().item()
    assert err <= 2e-2 * max(ref.abs().max().item(), 1.0), err
    assert torch.all(pb.tensor_of(ob)[:N, ..., :8] == 7.0) and torch.all(pb.tensor_of(ob)[:N, ..., 8 + cout:] == 7.0)
    _lib.lib().rsg_plan_destroy(h)


@pytest.mark.parametrize('C_,S_hw', [(32, (64, 48)), (16, (24, 16)), (48, (24, 18)), (64, (9, 7))])
def test_trp_fp32_tail_vs_torch(C_, S_hw):
    """attention with fp32 output + the fp32 tail  GroupNorm(8, C)(W y + b)  (association.py:236-245, 288-300) on an
    adversarial input: nearly uniform attention weights, so y's mean over the positions is ~50x its spread -- the
    regime in which a bf16 y / z loses the normalised signal (measured 0.10 relative error on the tiny model)."""
    N = 3
    H, W = S_hw
    S = H * W
    g = torch.Generator().manual_seed(C_ + S)
    x = (torch.randn(N, S, C_, generator=g) * 0.08).abs().bfloat16().float()          # post-ReLU features, small logits
    gv = (torch.randn(N, S, C_, generator=g) + 0.7).bfloat16().float()
    wt = torch.randn(C_, C_, generator=g) / C_ ** 0.5
    bias, gamma, beta = torch.randn(C_, generator=g) * 0.1, torch.rand(C_, generator=g) + 0.5, torch.randn(C_, generator=g) * 0.1
    pb = PlanBuilder(4, reuse=False)
    xb = pb.buf('x', H, W, C_)
    gb = pb.buf('g', H, W, C_)
    yb = pb.buf('y', H, W, C_)
    y32 = pb.buf('y32', H, W, C_, itemsize=4)
    ob = pb.buf('o', H, W, 2 * C_)
    f32 = lambda a: pb.const(a.numpy().astype(np.float32))
    pb.simple('attention', dict(x=View(xb), g=View(gb), y=View(yb), y32=y32), [xb, gb], [yb, y32])
    pb.simple('trptail', dict(y32=y32, out=View(ob, C_, C_), w=f32(wt), bias=f32(bias), gamma=f32(gamma), beta=f32(beta),
                              groups=8, S=S, C=C_), [y32], [ob])
    h = _run(pb, N)
    pb.tensor_of(xb)[:N] = x.view(N, H, W, C_).cuda().bfloat16()
    pb.tensor_of(gb)[:N] = gv.view(N, H, W, C_).cuda().bfloat16()
    pb.tensor_of(ob).fill_(7.0)
    _exec(h, N)
    A = torch.sigmoid(x.double() @ x.double().transpose(1, 2))
    y = A @ gv.double()                                                               # [N,S,C]
    got_y = pb.tensor_of(y32)[:N].reshape(N, S, C_).cpu().double()
    assert (got_y - y).abs().max().item() <= 6e-3 * y.abs().max().item()             # P is rounded to bf16 before P.G
    z = got_y @ wt.double().t() + bias.double()
    ref = F.group_norm(z.transpose(1, 2).reshape(N, C_, H, W), 8, gamma.double(), beta.double(), 1e-5)
    got = pb.tensor_of(ob)[:N, ..., C_:].float().permute(0, 3, 1, 2).cpu().double()
    # the tail itself, given the kernel's own fp32 y: fp32 arithmetic + one bf16 output rounding
    assert (got - ref).abs().max().item() <= 6e-3 * ref.abs().max().item()
    assert torch.all(pb.tensor_of(ob)[:N, ..., :C_] == 7.0)
    spread = (y - y.mean(1, keepdim=True)).abs().max().item() / y.abs().max().item()
    print(f'C={C_} S={S}: spread/|mean| of y = {spread:.4f}')
    _lib.lib().rsg_plan_destroy(h)
