"""Training kernels (rsg_train_* C ABI, called through the tape) against torch CPU autograd in float64.

Every case builds the op on the tape, seeds the output gradient with a random tensor R (i.e. the scalar is sum(out * R)),
runs the tape backward and compares output and every input / parameter gradient with torch's double-precision result.
`precise` (3xTF32) cases carry fp32-class bars (1e-4 of max: the tensor core's accumulator is not a round-to-nearest fp32
adder; 4e-5 measured at a reduction length of 5400); the TF32 cases the bar of a 10-bit-mantissa product."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

DEV = 'cuda:0'


def _tape(precise=True):
    from rsgnet_b200.train.tape import Tape
    torch.cuda.set_device(0)
    return Tape(torch.device(DEV), precise)


def _node(a, req=True):
    from rsgnet_b200.train.tape import T
    return T(torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(DEV), req=req)


def _param(a):
    from rsgnet_b200.train.tape import T
    v = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(DEV)
    return T(v, req=True, g=torch.zeros_like(v))


def _rel(a, b):
    a = a.detach().cpu().double().numpy() if torch.is_tensor(a) else np.asarray(a, np.float64)
    b = b.detach().cpu().double().numpy() if torch.is_tensor(b) else np.asarray(b, np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def _run(tape, out, rs):
    R = rs.standard_normal(out.shape).astype(np.float32)
    out.g = torch.from_numpy(R).to(DEV)
    tape.backward()
    torch.cuda.synchronize()
    return torch.from_numpy(R).double()


def nhwc(t):      # torch NCHW double -> NHWC
    return t.permute(0, 2, 3, 1).contiguous()


def pack_conv(w):                 # OIHW -> [tap][ci(pad 4)][co]
    co, ci, kh, kw = w.shape
    cip = (ci + 3) // 4 * 4
    p = np.zeros((kh * kw, cip, co), np.float32)
    p[:, :ci, :] = w.transpose(2, 3, 1, 0).reshape(kh * kw, ci, co)
    return p


def unpack_conv(g, shape):
    co, ci, kh, kw = shape
    return g[:, :ci, :].reshape(kh, kw, ci, co).permute(3, 2, 0, 1)


CONV_CASES = [  # N, H, W, Ci, Co, k, stride, bias
    (2, 12, 10, 16, 32, 3, 1, False),
    (3, 13, 9, 32, 16, 3, 2, False),
    (2, 16, 12, 3, 64, 3, 2, False),        # the stem: 3 input channels padded to 4
    (2, 9, 7, 20, 17, 3, 1, True),          # odd channel counts, bias (final layers with a 3x3 kernel)
    (1, 8, 8, 600, 16, 3, 1, False),        # type_conv's reduction length
    (2, 24, 16, 64, 64, 3, 1, False),
    (3, 24, 16, 256, 16, 3, 1, False),      # transition1: many input channels, few output channels, several pixel splits
    (3, 24, 16, 256, 32, 3, 2, False),
    (2, 24, 16, 64, 256, 1, 1, False),      # layer1's 1x1 convs (plain GEMM form), N = 256
    (2, 24, 16, 256, 64, 1, 1, True),
    (1, 9, 5, 96, 288, 3, 1, False),        # more than 256 output channels: two column tiles
    (5, 7, 3, 32, 32, 3, 1, False),         # 105 pixels: one partial row tile
    (4, 40, 30, 32, 32, 3, 1, False),       # narrow-layer weight-gradient kernel, several pixel splits
    (3, 33, 21, 16, 24, 3, 2, True),
    (2, 128, 96, 3, 32, 3, 2, False),
    (2, 16, 12, 128, 128, 3, 1, False),     # tcgen05 weight gradients: 128-row accumulators
    (3, 8, 6, 256, 256, 3, 1, True),        # ... 2 x 4 tiles, 3 kernel rows, position splits
    (1, 5, 7, 64, 192, 3, 1, False),        # ... 64-row accumulators (M = 64 TMEM layout), one chunk
    (5, 32, 24, 128, 64, 3, 1, False),
    (2, 10, 8, 32, 64, 3, 1, False),        # ... partial tiles: 32 of 64 rows
    (2, 10, 8, 64, 32, 3, 1, True),         # ... 32 of 64 columns
    (2, 7, 9, 40, 72, 3, 1, False),         # ... 40 rows, 64 + 8 columns
    (1, 12, 9, 200, 24, 3, 1, False),       # ... 128 + 72 rows, 24 columns (N = 32 MMAs)
]


@pytest.mark.parametrize('precise,tol', [(1, 1e-4), (0, 4e-3), (2, 4e-3)])      # 3xTF32 | TF32 tcgen05 | TF32 mma.sync
@pytest.mark.parametrize('case', CONV_CASES)
def test_conv_forward_dgrad_wgrad(case, precise, tol):
    N, H, W, Ci, Co, k, s, bias = case
    rs = np.random.RandomState(1)
    x = rs.standard_normal((N, Ci, H, W)).astype(np.float32)
    w = (rs.standard_normal((Co, Ci, k, k)) / np.sqrt(Ci * k * k)).astype(np.float32)
    b = rs.standard_normal(Co).astype(np.float32) if bias else None
    cip = (Ci + 3) // 4 * 4
    xp = np.zeros((N, H, W, cip), np.float32)
    xp[..., :Ci] = x.transpose(0, 2, 3, 1)
    tape = _tape(precise)
    xn, wn = _node(xp), _param(pack_conv(w))
    bn = _param(b) if bias else None
    wT = wn.v.permute(0, 2, 1).contiguous()
    out = tape.conv(xn, wn, wT, k, s, k // 2, bn)
    R = _run(tape, out, rs)
    xt = torch.from_numpy(x).double().requires_grad_(True)
    wt = torch.from_numpy(w).double().requires_grad_(True)
    bt = torch.from_numpy(b).double().requires_grad_(True) if bias else None
    ref = F.conv2d(xt, wt, bt, s, k // 2)
    (nhwc(ref) * R).sum().backward()
    errs = (_rel(out.v, nhwc(ref)), _rel(xn.g[..., :Ci], nhwc(xt.grad)), _rel(unpack_conv(wn.g.cpu(), w.shape), wt.grad))
    print(f'conv {case} precise={precise}: fwd {errs[0]:.1e} dgrad {errs[1]:.1e} wgrad {errs[2]:.1e}')
    assert max(errs) < tol
    if bias:
        assert _rel(bn.g, bt.grad) < tol


@pytest.mark.parametrize('precise,tol', [(1, 1e-4), (0, 4e-3), (2, 4e-3)])
@pytest.mark.parametrize('case', [(3, 20, 14, 32, 32, 3), (2, 24, 16, 64, 64, 3), (2, 9, 7, 20, 20, 3), (2, 12, 10, 64, 64, 1),
                                  (2, 16, 12, 128, 128, 3)])
def test_conv_input_gradients_of_a_shared_input_add_up(case, precise, tol):
    """Two convs read the same x (the residual pattern of every BasicBlock): the first backward closure creates x.g, the second
    one's input gradient is added to it.  (Accumulating in the conv epilogue instead -- beta = 1 on the flat kernels -- saved
    113 launches per step and was measured 0.3 % SLOWER on one box: the read-modify-write sits on the critical path.)"""
    N, H, W, Ci, Co, k = case
    rs = np.random.RandomState(5)
    x = rs.standard_normal((N, Ci, H, W)).astype(np.float32)
    w1 = (rs.standard_normal((Co, Ci, k, k)) / np.sqrt(Ci * k * k)).astype(np.float32)
    w2 = (rs.standard_normal((Co, Ci, k, k)) / np.sqrt(Ci * k * k)).astype(np.float32)
    cip = (Ci + 3) // 4 * 4
    xp = np.zeros((N, H, W, cip), np.float32)
    xp[..., :Ci] = x.transpose(0, 2, 3, 1)
    tape = _tape(precise)
    xn = _node(xp)
    outs = []
    for w in (w1, w2):
        wn = _param(pack_conv(w))
        outs.append(tape.conv(xn, wn, wn.v.permute(0, 2, 1).contiguous(), k, 1, k // 2, None))
    out = tape.add(outs)
    R = _run(tape, out, rs)
    xt = torch.from_numpy(x).double().requires_grad_(True)
    ref = F.conv2d(xt, torch.from_numpy(w1).double(), None, 1, k // 2) + F.conv2d(xt, torch.from_numpy(w2).double(), None, 1, k // 2)
    (nhwc(ref) * R).sum().backward()
    err = _rel(xn.g[..., :Ci], nhwc(xt.grad))
    print(f'conv shared input {case} precise={precise}: dgrad sum {err:.1e}')
    assert err < tol


@pytest.mark.parametrize('precise,tol', [(True, 1e-4), (False, 4e-3)])
@pytest.mark.parametrize('case', [(2, 6, 5, 16, 16), (1, 12, 8, 32, 32), (3, 4, 4, 8, 20)])
def test_conv_transpose_4_2_1(case, precise, tol):
    N, h, w_, Ci, Co = case
    rs = np.random.RandomState(2)
    x = rs.standard_normal((N, Ci, h, w_)).astype(np.float32)
    w = (rs.standard_normal((Ci, Co, 4, 4)) / np.sqrt(Ci * 4)).astype(np.float32)
    tape = _tape(precise)
    xn = _node(x.transpose(0, 2, 3, 1))
    wn = _param(w.transpose(2, 3, 0, 1).reshape(16, Ci, Co))
    out = tape.conv_transpose(xn, wn, wn.v.permute(0, 2, 1).contiguous(), 4, 2, 1, 0)
    R = _run(tape, out, rs)
    xt = torch.from_numpy(x).double().requires_grad_(True)
    wt = torch.from_numpy(w).double().requires_grad_(True)
    ref = F.conv_transpose2d(xt, wt, None, 2, 1)
    (nhwc(ref) * R).sum().backward()
    assert _rel(out.v, nhwc(ref)) < tol
    assert _rel(xn.g, nhwc(xt.grad)) < tol
    assert _rel(wn.g.cpu().reshape(4, 4, Ci, Co).permute(2, 3, 0, 1), wt.grad) < tol


@pytest.mark.parametrize('precise,tol', [(True, 1e-4), (False, 4e-3)])
@pytest.mark.parametrize('case', [(384, 32, 32, True), (17, 600, 600, False), (700, 17, 600, False), (300, 96, 14, True),
                                  (18, 32, 32, True), (1000, 4, 32, False)])
def test_linear(case, precise, tol):
    M, I, O, bias = case
    rs = np.random.RandomState(3)
    x = rs.standard_normal((M, I)).astype(np.float32)
    w = (rs.standard_normal((O, I)) / np.sqrt(I)).astype(np.float32)
    b = rs.standard_normal(O).astype(np.float32) if bias else None
    tape = _tape(precise)
    xn, wn = _node(x), _param(w)
    bn = _param(b) if bias else None
    out = tape.linear(xn, wn, bn)
    R = _run(tape, out, rs)
    xt, wt = torch.from_numpy(x).double().requires_grad_(True), torch.from_numpy(w).double().requires_grad_(True)
    bt = torch.from_numpy(b).double().requires_grad_(True) if bias else None
    ref = F.linear(xt, wt, bt)
    (ref * R).sum().backward()
    assert _rel(out.v, ref) < tol
    assert _rel(xn.g, xt.grad) < tol
    assert _rel(wn.g, wt.grad) < tol
    if bias:
        assert _rel(bn.g, bt.grad) < tol


@pytest.mark.parametrize('case', [(18, 17, 32), (768, 17, 600), (130, 70, 33)])
def test_matmul_both_gradients(case):
    M, K, N = case
    rs = np.random.RandomState(4)
    a, b = rs.standard_normal((M, K)).astype(np.float32), rs.standard_normal((K, N)).astype(np.float32)
    tape = _tape(True)
    an, bn = _node(a), _node(b)                 # b is a COMPUTED weight: its gradient buffer is created on demand
    out = tape.matmul(an, bn)
    R = _run(tape, out, rs)
    at, bt = torch.from_numpy(a).double().requires_grad_(True), torch.from_numpy(b).double().requires_grad_(True)
    ref = at @ bt
    (ref * R).sum().backward()
    assert _rel(out.v, ref) < 2e-5 and _rel(an.g, at.grad) < 2e-5 and _rel(bn.g, bt.grad) < 2e-5


@pytest.mark.parametrize('relu', [False, True])
@pytest.mark.parametrize('M,C', [(2 * 24 * 16, 32), (17, 600), (3 * 7 * 5, 17), (4096, 256), (1000, 4), (640, 96)])
def test_batchnorm_train(M, C, relu):
    rs = np.random.RandomState(5)
    x = (rs.standard_normal((M, C)) * rs.uniform(0.5, 2, C) + rs.uniform(-1, 1, C)).astype(np.float32)
    gam, bet = rs.uniform(0.5, 1.5, C).astype(np.float32), rs.standard_normal(C).astype(np.float32)
    rm, rv = rs.standard_normal(C).astype(np.float32), rs.uniform(0.5, 1.5, C).astype(np.float32)
    tape = _tape()
    xn, gn, bn = _node(x), _param(gam), _param(bet)
    rmd, rvd = torch.from_numpy(rm).to(DEV), torch.from_numpy(rv).to(DEV)
    out = tape.batchnorm(xn, gn, bn, rmd, rvd, relu, 1e-5, 0.1)
    R = _run(tape, out, rs)
    xt = torch.from_numpy(x).double().requires_grad_(True)
    gt, bt = torch.from_numpy(gam).double().requires_grad_(True), torch.from_numpy(bet).double().requires_grad_(True)
    rmt, rvt = torch.from_numpy(rm).double(), torch.from_numpy(rv).double()
    ref = F.batch_norm(xt, rmt, rvt, gt, bt, True, 0.1, 1e-5)
    if relu:
        ref = F.relu(ref)
    (ref * R).sum().backward()
    assert _rel(out.v, ref) < 1e-5
    assert _rel(xn.g, xt.grad) < 5e-5
    assert _rel(gn.g, gt.grad) < 2e-5 and _rel(bn.g, bt.grad) < 2e-5
    assert _rel(rmd, rmt) < 1e-6 and _rel(rvd, rvt) < 1e-5


@pytest.mark.parametrize('B,S,C', [(2, 384, 16), (3, 96, 32), (1, 3072, 32), (2, 100, 48)])
def test_groupnorm(B, S, C):
    rs = np.random.RandomState(6)
    x = (rs.standard_normal((B, S, C)) + 3.0).astype(np.float32)
    gam, bet = rs.uniform(0.5, 1.5, C).astype(np.float32), rs.standard_normal(C).astype(np.float32)
    tape = _tape()
    xn, gn, bn = _node(x), _param(gam), _param(bet)
    out = tape.groupnorm(xn, gn, bn, 8, 1e-5)
    R = _run(tape, out, rs)
    xt = torch.from_numpy(x).double().requires_grad_(True)
    gt, bt = torch.from_numpy(gam).double().requires_grad_(True), torch.from_numpy(bet).double().requires_grad_(True)
    ref = F.group_norm(xt.permute(0, 2, 1), 8, gt, bt, 1e-5).permute(0, 2, 1)
    (ref * R).sum().backward()
    assert _rel(out.v, ref) < 1e-5 and _rel(xn.g, xt.grad) < 5e-5
    assert _rel(gn.g, gt.grad) < 2e-5 and _rel(bn.g, bt.grad) < 2e-5


def test_elementwise_ops_and_fanout():
    rs = np.random.RandomState(7)
    a, b, c = (rs.standard_normal((2, 6, 5, 16)).astype(np.float32) for _ in range(3))
    mask = (rs.uniform(size=(18, 17)) > 0.5).astype(np.float32)
    m = rs.standard_normal((18, 17)).astype(np.float32)
    tape = _tape()
    an, bn, cn, mn = _node(a), _node(b), _node(c), _param(m)
    s1 = tape.add([an, bn, cn], relu=True)
    s2 = tape.add([s1, an], relu=False)                      # `an` is used twice: gradient accumulation
    y = tape.leaky_relu(tape.sigmoid(s2), 0.02)
    z = tape.cat([y, s1])
    mm = tape.mul_const(mn, torch.from_numpy(mask).to(DEV))
    Rz = rs.standard_normal(z.shape).astype(np.float32)
    Rm = rs.standard_normal(mm.shape).astype(np.float32)
    z.g, mm.g = torch.from_numpy(Rz).to(DEV), torch.from_numpy(Rm).to(DEV)
    tape.backward()
    at, bt, ct = (torch.from_numpy(v).double().requires_grad_(True) for v in (a, b, c))
    mt = torch.from_numpy(m).double().requires_grad_(True)
    r1 = F.relu(at + bt + ct)
    r2 = r1 + at
    ry = F.leaky_relu(torch.sigmoid(r2), 0.02)
    rz = torch.cat([ry, r1], -1)
    rmm = mt * torch.from_numpy(mask).double()
    ((rz * torch.from_numpy(Rz).double()).sum() + (rmm * torch.from_numpy(Rm).double()).sum()).backward()
    assert _rel(z.v, rz) < 1e-6 and _rel(mm.v, rmm) < 1e-6
    for n, t in ((an, at), (bn, bt), (cn, ct), (mn, mt)):
        assert _rel(n.g, t.grad) < 1e-5


@pytest.mark.parametrize('f', [2, 4, 8])
def test_nearest_upsample(f):
    rs = np.random.RandomState(8)
    x = rs.standard_normal((2, 3, 4, 8)).astype(np.float32)
    tape = _tape()
    xn = _node(x)
    out = tape.upsample_nearest(xn, f)
    R = _run(tape, out, rs)
    xt = torch.from_numpy(x).double().requires_grad_(True)
    ref = F.interpolate(xt.permute(0, 3, 1, 2), scale_factor=f, mode='nearest').permute(0, 2, 3, 1)
    (ref * R).sum().backward()
    assert _rel(out.v, ref) == 0.0 and _rel(xn.g, xt.grad) < 1e-6


@pytest.mark.parametrize('shape', [(2, 24, 16, 17), (1, 5, 7, 14), (2, 64, 48, 18)])
def test_bilinear2x_align_corners(shape):
    rs = np.random.RandomState(9)
    x = rs.standard_normal(shape).astype(np.float32)
    tape = _tape()
    xn = _node(x)
    out = tape.bilinear2x(xn)
    R = _run(tape, out, rs)
    xt = torch.from_numpy(x).double().requires_grad_(True)
    ref = F.interpolate(xt.permute(0, 3, 1, 2), scale_factor=2, mode='bilinear', align_corners=True).permute(0, 2, 3, 1)
    (ref * R).sum().backward()
    assert _rel(out.v, ref) < 1e-5 and _rel(xn.g, xt.grad) < 2e-5      # fp32 source coordinates, as torch's own fp32 kernel


def test_maxpool_repeat_layout():
    rs = np.random.RandomState(10)
    x = rs.standard_normal((2, 6, 8, 16)).astype(np.float32)
    l = rs.standard_normal((1, 5, 4, 8)).astype(np.float32)
    img = rs.standard_normal((2, 3, 8, 6)).astype(np.float32)
    tape = _tape()
    xn, ln = _node(x), _node(l)
    mp = tape.maxpool2(xn)
    rp = tape.repeat_batch(ln, 3)
    nch = tape.to_nchw(mp)
    im = tape.from_nchw(torch.from_numpy(img).to(DEV), 4)
    Rn = rs.standard_normal(nch.shape).astype(np.float32)
    Rr = rs.standard_normal(rp.shape).astype(np.float32)
    nch.g, rp.g = torch.from_numpy(Rn).to(DEV), torch.from_numpy(Rr).to(DEV)
    tape.backward()
    xt, lt = torch.from_numpy(x).double().requires_grad_(True), torch.from_numpy(l).double().requires_grad_(True)
    rmp = F.max_pool2d(xt.permute(0, 3, 1, 2), 2)
    rrp = lt.repeat(3, 1, 1, 1)
    ((rmp * torch.from_numpy(Rn).double()).sum() + (rrp * torch.from_numpy(Rr).double()).sum()).backward()
    assert _rel(nch.v, rmp) == 0.0 and _rel(rp.v, rrp) == 0.0
    assert _rel(xn.g, xt.grad) < 1e-6 and _rel(ln.g, lt.grad) < 1e-5
    want = np.zeros((2, 8, 6, 4), np.float32)
    want[..., :3] = img.transpose(0, 2, 3, 1)
    assert np.array_equal(im.v.cpu().numpy(), want)


@pytest.mark.parametrize('B,S,C,mode', [(2, 384, 16, 'vec'), (2, 96, 32, 'full'), (1, 200, 32, 'direct'), (2, 130, 16, 'none')])
def test_trp_attention_and_relation_loss(B, S, C, mode):
    """association.py:288-299 + pose_rsgnet.py:1014-1018: y = sigmoid(x x^T) g; the relation loss 0.001 * mean_b mean_ij (T - P)^2
    enters through the hook (rank-1 factor or full target) or as a direct gradient on P."""
    rs = np.random.RandomState(11)
    x = (0.3 * rs.standard_normal((B, S, C))).astype(np.float32)
    g = rs.standard_normal((B, S, C)).astype(np.float32)
    v = rs.uniform(0, 1, (B, S)).astype(np.float32)
    tape = _tape()
    xn, gn = _node(x), _node(g)
    out, Pn, hook = tape.trp_attention(xn, gn)
    vt = torch.from_numpy(v).double()
    Tt = vt[:, :, None] * vt[:, None, :]
    coef = 0.001 / B * 2.0 / (S * S)
    RP = None
    if mode == 'vec':
        hook['rel'] = (None, torch.from_numpy(v).to(DEV), torch.full((B,), coef, device=DEV))
    elif mode == 'full':
        hook['rel'] = (Tt.float().to(DEV).contiguous(), None, torch.full((B,), coef, device=DEV))
    elif mode == 'direct':
        RP = rs.standard_normal((B, S, S)).astype(np.float32)
        Pn.g = torch.from_numpy(RP).to(DEV)
    R = _run(tape, out, rs)
    xt, gt = torch.from_numpy(x).double().requires_grad_(True), torch.from_numpy(g).double().requires_grad_(True)
    P = torch.sigmoid(xt @ xt.transpose(1, 2))
    ref = P @ gt
    loss = (ref * R).sum()
    if mode in ('vec', 'full'):
        loss = loss + 0.001 * ((Tt - P) ** 2).mean(dim=(1, 2)).mean()
    if mode == 'direct':
        loss = loss + (P * torch.from_numpy(RP).double()).sum()
    loss.backward()
    assert _rel(Pn.v, P) < 1e-5 and _rel(out.v, ref) < 2e-5
    assert _rel(gn.g, gt.grad) < 2e-5 and _rel(xn.g, xt.grad) < 5e-5
    # the forward value of the relation loss
    acc = torch.zeros(B, dtype=torch.float64, device=DEV)
    from rsgnet_b200.train.tape import _p
    if mode == 'vec':
        d_v = torch.from_numpy(v).to(DEV)
        tape.call('rsg_train_relation_mse', _p(Pn.v), None, _p(d_v), B, S, _p(acc))
        assert _rel(acc, ((Tt - P) ** 2).mean(dim=(1, 2))) < 1e-5


def test_losses_person_mask_adam():
    import ctypes as C
    from rsgnet_b200 import _lib
    from rsgnet_b200.train.tape import _p
    rs = np.random.RandomState(12)
    B, K, H, W, L = 3, 17, 24, 16, 18
    tape = _tape()
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(DEV)
    pred, tgt = rs.standard_normal((B, K, H, W)), rs.uniform(0, 1, (B, K, H, W))
    tw = (rs.uniform(size=(B, K, 1)) > 0.3).astype(np.float32)
    acc = torch.zeros(4, dtype=torch.float64, device=DEV)
    grad = torch.empty(B, K, H, W, device=DEV)
    d_pred, d_tgt, d_tw = dev(pred), dev(tgt), dev(tw)        # keep the device copies alive across the asynchronous calls
    tape.call('rsg_train_mse_joints', _p(d_pred), _p(d_tgt), _p(d_tw), B, K, H * W, 1.0, _p(acc), _p(grad))
    pt = torch.from_numpy(pred).double().requires_grad_(True)
    from oracle import train_oracle
    ref = train_oracle.joints_mse(pt, torch.from_numpy(tgt).double(), torch.from_numpy(tw).double())
    ref.backward()
    assert abs(float(acc[0]) - float(ref)) < 1e-6 * float(ref) and _rel(grad, pt.grad) < 1e-5
    p = rs.uniform(0.001, 0.999, (B, L, H, W))
    p[0, 0, 0, :3] = (0.0, 1.0, 1e-30)                       # the clamped-log corner of BCELoss
    t = rs.uniform(0, 1, (B, L, H, W))
    g2 = torch.empty(B, L, H, W, device=DEV)
    d_p, d_t = dev(p), dev(t)
    tape.call('rsg_train_bce', _p(d_p), _p(d_t), p.size, 0.01, 1.0, C.c_void_p(acc.data_ptr() + 8), _p(g2))
    p32 = torch.from_numpy(p.astype(np.float32)).double().requires_grad_(True)
    rb = 0.01 * F.binary_cross_entropy(p32, torch.from_numpy(t.astype(np.float32)).double())
    rb.backward()
    assert abs(float(acc[1]) - float(rb)) < 2e-6 * float(rb)
    ok = (p > 1e-6) & (p < 1 - 1e-6)
    assert _rel(g2.cpu()[torch.from_numpy(ok)], p32.grad[torch.from_numpy(ok)]) < 1e-4
    # relation-target factor (lib/core/function.py:256-267)
    vec = torch.empty(B, (H // 2) * (W // 2), device=DEV)
    tape.call('rsg_train_person_mask', _p(d_tgt), B, K, H, W, _p(vec))
    person = torch.from_numpy(tgt.astype(np.float32)).max(dim=1)[0].reshape(B, 1, H, W)
    want = F.interpolate(person, scale_factor=0.5, mode='bilinear', align_corners=True).reshape(B, -1)
    assert _rel(vec, want) < 1e-6
    # Adam against torch.optim.Adam for three steps
    n = 10007
    w0 = rs.standard_normal(n).astype(np.float32)
    pw = torch.nn.Parameter(torch.from_numpy(w0.copy()))
    opt = torch.optim.Adam([pw], lr=1e-3)
    dw, m, v = dev(w0), torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    for step in range(1, 4):
        gnp = rs.standard_normal(n).astype(np.float32)
        pw.grad = torch.from_numpy(gnp.copy())
        opt.step()
        d_g = dev(gnp * 2)
        _lib.check(_lib.lib().rsg_train_adam(tape.st, _p(dw), _p(d_g), _p(m), _p(v), n, 1e-3, 0.9, 0.999, 1e-8, step, 0.5))
        torch.cuda.synchronize()
    assert _rel(dw, pw.data) < 1e-6


@pytest.mark.parametrize('M,C', [(2 * 24 * 16, 32), (4096, 256), (777, 64)])
def test_batchnorm_add_relu_fused(M, C):
    """relu(bn(x) + res) and its backward (BasicBlock / Bottleneck tails, pose_rsgnet.py:47-52, 88-93) in the BN's two launches."""
    rs = np.random.RandomState(15)
    x = (rs.standard_normal((M, C)) * rs.uniform(0.5, 2, C) + rs.uniform(-1, 1, C)).astype(np.float32)
    r = rs.standard_normal((M, C)).astype(np.float32)
    gam, bet = rs.uniform(0.5, 1.5, C).astype(np.float32), rs.standard_normal(C).astype(np.float32)
    rm, rv = rs.standard_normal(C).astype(np.float32), rs.uniform(0.5, 1.5, C).astype(np.float32)
    tape = _tape()
    xn, rn, gn, bn = _node(x), _node(r), _param(gam), _param(bet)
    rmd, rvd = torch.from_numpy(rm).to(DEV), torch.from_numpy(rv).to(DEV)
    out = tape.batchnorm_add_relu(xn, gn, bn, rmd, rvd, rn, 1e-5, 0.1)
    R = _run(tape, out, rs)
    xt, rt = torch.from_numpy(x).double().requires_grad_(True), torch.from_numpy(r).double().requires_grad_(True)
    gt, bt = torch.from_numpy(gam).double().requires_grad_(True), torch.from_numpy(bet).double().requires_grad_(True)
    rmt, rvt = torch.from_numpy(rm).double(), torch.from_numpy(rv).double()
    ref = F.relu(F.batch_norm(xt, rmt, rvt, gt, bt, True, 0.1, 1e-5) + rt)
    (ref * R).sum().backward()
    assert _rel(out.v, ref) < 1e-5
    assert _rel(xn.g, xt.grad) < 5e-5 and _rel(rn.g, rt.grad) < 1e-6
    assert _rel(gn.g, gt.grad) < 2e-5 and _rel(bn.g, bt.grad) < 2e-5
    assert _rel(rmd, rmt) < 1e-6 and _rel(rvd, rvt) < 1e-5
