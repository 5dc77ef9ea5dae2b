"""Host side of the model path: fold + pack the reference-named parameters, describe the network
to librsg_b200 as a flat op list (the "plan"), and run it.

What is folded at pack time (exact algebraic identities in eval mode):
  * BatchNorm into the preceding bias-free conv (pose_rsgnet.py:38-54 etc.);
  * the type branch  type_fc -> BN1d -> ReLU -> scores^T.T -> type_conv  into one K->C0 3x3 conv
    (pose_rsgnet.py:966-977);
  * the loc branch  loc_conv(loc_features)  into a constant map stored once in the loc slice of the
    contact_conv input (pose_rsgnet.py:979-980);
  * KTMachine(final_layer.weight) into a constant 1x1 conv (pose_rsgnet.py:592-600, 1004-1005);
  * torch.cat by writing producers into channel slices of one buffer (pose_rsgnet.py:982, 991).
Packing is host-side NumPy on the parameters only; every activation is computed on the device by
the library.  Nothing here falls back to PyTorch ops for the forward.
"""
import ctypes as C
import threading

import numpy as np
import torch

from . import _lib
from ._lib import ConvDesc, Ref, Res, abs_ref, ext_ref, null_ref
from .config import KIND_RSGNET

EPS = 1e-5
EXT_X, EXT_HEAT, EXT_MULTI, EXT_LIMBS, EXT_REL = 0, 1, 2, 3, 4
N_EXT = 5


def _np(t):
    return t.detach().to('cpu', torch.float64).numpy()


def _round_up(v, m):
    return (v + m - 1) // m * m


def _bf16_bits(a):
    """fp32 -> bf16 (round to nearest even) as uint16."""
    a = np.ascontiguousarray(a, np.float32)
    u = a.view(np.uint32).astype(np.uint64)
    r = ((u >> 16) & 1) + 0x7FFF
    return ((u + r) >> 16).astype(np.uint16)


class _Params:
    """state_dict accessor (fp64 NumPy) with prefixes."""

    def __init__(self, sd, prefix=''):
        self.sd, self.prefix = sd, prefix

    def sub(self, name):
        return _Params(self.sd, f'{self.prefix}{name}.')

    def has(self, name):
        return f'{self.prefix}{name}' in self.sd

    def __getitem__(self, name):
        return _np(self.sd[f'{self.prefix}{name}'])

    def bn(self, name):
        """(scale, shift) of an eval-mode BatchNorm."""
        p = self.sub(name)
        s = p['weight'] / np.sqrt(p['running_var'] + EPS)
        return s, p['bias'] - p['running_mean'] * s


class Buf:
    """Symbolic activation buffer: per forward `elems` items of `itemsize` bytes."""

    def __init__(self, name, H, W, C, itemsize=2, persistent=False):
        self.name, self.H, self.W, self.C, self.itemsize = name, H, W, C, itemsize
        self.persistent = persistent
        self.first = self.last = None
        self.ptr = None
        self.tensor = None

    @property
    def bytes_per_fwd(self):
        return self.H * self.W * self.C * self.itemsize


class View:
    def __init__(self, buf, co=0, C=None):
        self.buf, self.co, self.C = buf, co, (buf.C - co if C is None else C)

    @property
    def H(self):
        return self.buf.H

    @property
    def W(self):
        return self.buf.W


class _Arena:
    """Host-side staging of packed constants -> one device tensor."""

    def __init__(self):
        self.chunks, self.size = [], 0

    def add(self, arr):
        arr = np.ascontiguousarray(arr)
        off = self.size
        self.chunks.append((off, arr))
        self.size = _round_up(off + arr.nbytes, 256)
        return off

    def upload(self, device):
        host = np.zeros(max(self.size, 256), np.uint8)
        for off, arr in self.chunks:
            host[off:off + arr.nbytes] = arr.view(np.uint8).reshape(-1)
        return torch.from_numpy(host).to(device)


class PlanBuilder:
    def __init__(self, chunk, reuse=True):
        self.chunk = chunk
        self.reuse = reuse     # False: every buffer keeps its own memory (debug taps)
        self.ops = []          # (kind, payload, reads, writes, aux)
        self.bufs = []
        self.arena = _Arena()
        self.aux = False
        self.flops_per_fwd = 0   # executed MACs*2 (after folding)

    # -- buffers ---------------------------------------------------------------------------
    def tensor_of(self, x):
        """torch view of a Buf / View after allocate(): bf16 [chunk,H,W,C(slice)] or f32 [chunk,C,H,W]."""
        b = x.buf if isinstance(x, View) else x
        raw = b.tensor if b.tensor is not None else self.work[b.ptr - self.work.data_ptr():][:self.chunk * b.bytes_per_fwd]
        if b.itemsize == 4:
            return raw.view(torch.float32).view(self.chunk, b.C, b.H, b.W)
        t = raw.view(torch.bfloat16).view(self.chunk, b.H, b.W, b.C)
        return t[..., x.co:x.co + x.C] if isinstance(x, View) else t

    def buf(self, name, H, W, C, itemsize=2, persistent=False):
        b = Buf(name, H, W, C, itemsize, persistent or not self.reuse)
        self.bufs.append(b)
        return b

    def const(self, arr):
        return ('const', self.arena.add(arr))

    # -- ops -------------------------------------------------------------------------------
    def _add(self, kind, payload, reads, writes):
        idx = len(self.ops)
        for b in list(reads) + list(writes):
            if b is None:
                continue
            if b.first is None:
                b.first = idx
            b.last = idx
        self.ops.append((kind, payload, self.aux))

    def conv(self, src, w_oihw, bias, stride=1, relu=False, dst=None, res=(), taps=None,
             out_f32=None, omul=1, ooy=0, oox=0, out_hw=None, name='conv', cout=None, engine=0,
             pixel_shuffle_c=0):
        """w_oihw: folded fp64 [Cout, Cin, kh, kw] (or [ntaps, Cout, Cin] with explicit taps)."""
        if taps is None:
            co_, ci_, kh, kw = w_oihw.shape
            pad = kh // 2
            taps = [(ky - pad, kx - pad) for ky in range(kh) for kx in range(kw)]
            w_t = w_oihw.transpose(2, 3, 0, 1).reshape(kh * kw, co_, ci_)
        else:
            w_t = w_oihw
            _, co_, ci_ = w_t.shape
        assert ci_ <= src.C, (name, ci_, src.C)
        cin = _round_up(ci_, 8)
        assert cin <= src.buf.C - src.co
        cin_pad = _round_up(cin, 32)
        cout = co_ if cout is None else cout
        cout_pad = _round_up(cout, 32)
        wp = np.zeros((len(taps), cout_pad, cin_pad), np.float32)
        wp[:, :co_, :ci_] = w_t
        bp = np.zeros(cout_pad, np.float32)
        bp[:co_] = bias
        Hin, Win = src.H, src.W
        Hout = (Hin - 1) // stride + 1 if out_hw is None else out_hw[0]
        Wout = (Win - 1) // stride + 1 if out_hw is None else out_hw[1]
        w_tc5 = None
        tc5_ok = (stride in (1, 2) and (omul == 1 or not res) and out_f32 is None and cin % 16 == 0
                  and cout % 8 == 0 and all(abs(dy) <= 1 and abs(dx) <= 1 for dy, dx in taps) and len(taps) <= 9)
        ns = C.c_int()
        # GEMM-shaped layers (K = taps*Cin >= 1152, >= 128 output channels) on maps small enough for the
        # flat formulation: weight-streaming kernel
        # ... and the stride-2 down-paths whose OUTPUT map is small (<= 16x12: the 3x3 s2 convs of the fuse layers and
        # transitions onto branches 2 and 3), in the flat stride-2 formulation of the same kernel
        want_ws = (omul == 1 and
                   (engine in (3, 4) or (engine == 0 and stride == 1 and len(taps) * cin >= 1152 and cout_pad >= 128)
                    or (engine == 0 and stride == 2 and len(taps) == 9 and cout_pad >= 64 and Hout * Wout <= 192)))
        # the CTA-pair kernel is opt-in (engine=4): measured equal to the 1-CTA kernel in isolation and ~7 % slower inside
        # the step (profiles/r2_notes.md), so the automatic routing keeps conv_ws
        if (tc5_ok and want_ws and dst is not None and engine == 4
                and _lib.lib().rsg_conv_ws2_config(cin, cout_pad, len(taps), Hin, Win)):
            # CTA-pair kernel (tcgen05.mma.cta_group::2): each CTA streams the 64 output channels of its half
            # [slice][half][16-channel chunk][tap][2 planes][64][8]: one contiguous block (one bulk copy) per chunk; the
            # output channels of every group of 64 in accumulator-column order (the epilogue reads 16x256b TMEM blocks)
            wt = wp[:, _quad_perm(cout_pad), :cin].reshape(len(taps), cout_pad // 128, 2, 64, cin // 16, 2, 8)
            w_tc5 = self.const(_bf16_bits(wt.transpose(1, 2, 4, 0, 5, 3, 6)))
            engine = 4
        elif engine == 4:
            raise ValueError(f'{name}: shape not covered by the CTA-pair weight-streaming kernel')
        elif (tc5_ok and want_ws and dst is not None and engine != 4
                and _lib.lib().rsg_conv_ws_config2(cin, cout_pad, len(taps), Hin, Win, stride, C.byref(ns))):
            NS = ns.value
            # [slice][16-channel chunk][tap][2 planes][NS][8]: one contiguous block per chunk
            wt = wp[:, :, :cin].reshape(len(taps), cout_pad // NS, NS, cin // 16, 2, 8)
            w_tc5 = self.const(_bf16_bits(wt.transpose(1, 3, 0, 4, 2, 5)))
            engine = 3
        elif engine == 3:
            raise ValueError(f'{name}: shape not covered by the weight-streaming kernel')
        elif tc5_ok:
            mode = 2 if stride == 2 else int(any(dy != 0 or dx != 0 for dy, dx in taps))
            ns, kc, st = C.c_int(), C.c_int(), C.c_int()
            if _lib.lib().rsg_conv_tc5_config(cin, cout_pad, len(taps), mode, C.byref(ns), C.byref(kc), C.byref(st)):
                # [slice][tap][Cin/8][NS][8]: the K-major core-matrix order the kernel bulk-copies
                NS = ns.value
                wt = wp[:, :, :cin].reshape(len(taps), cout_pad // NS, NS, cin // 8, 8)
                w_tc5 = self.const(_bf16_bits(wt.transpose(1, 0, 3, 2, 4)))
        payload = dict(src=src, w=self.const(_bf16_bits(wp)), w_tc5=w_tc5, bias=self.const(bp), cin=cin,
                       cout=cout, cout_pad=cout_pad, taps=taps, stride=stride, Hout=Hout, Wout=Wout,
                       dst=dst, out_f32=out_f32, res=list(res), relu=relu, omul=omul, ooy=ooy,
                       oox=oox, name=name, engine=engine, psc=pixel_shuffle_c)
        if pixel_shuffle_c and w_tc5 is None:
            raise ValueError(f'{name}: pixel-shuffle output needs the tcgen05 kernel')
        reads = [src.buf] + [r[0].buf for r in res]
        writes = [dst.buf if dst is not None else None,
                  out_f32 if isinstance(out_f32, Buf) else None]
        self._add('conv', payload, reads, writes)
        self.flops_per_fwd += 2 * len(taps) * ci_ * co_ * Hout * Wout
        return dst

    def simple(self, kind, payload, reads, writes):
        self._add(kind, payload, reads, writes)

    # -- finalise --------------------------------------------------------------------------
    def allocate(self, device):
        """Lifetime-based first-fit placement of the non-persistent buffers in one arena."""
        live = []      # (offset, size, last)
        total = 0
        order = sorted([b for b in self.bufs if b.first is not None and not b.persistent],
                       key=lambda b: b.first)
        placed = {}
        for b in order:
            size = _round_up(b.bytes_per_fwd * self.chunk, 1024)
            live = [(o, s, l) for (o, s, l) in live if l >= b.first]
            live.sort()
            off = 0
            for (o, s, l) in live:
                if off + size <= o:
                    break
                off = max(off, o + s)
            placed[b] = off
            live.append((off, size, b.last))
            total = max(total, off + size)
        self.work = torch.zeros(max(total, 1024), dtype=torch.uint8, device=device)
        base = self.work.data_ptr()
        for b, off in placed.items():
            b.ptr = base + off
        for b in self.bufs:
            if b.persistent or b.first is None:
                b.tensor = torch.zeros(self.chunk * b.bytes_per_fwd, dtype=torch.uint8, device=device)
                b.ptr = b.tensor.data_ptr()
        self.consts = self.arena.upload(device)
        self.work_bytes = total
        return self


def _ref(x, consts_ptr):
    """Buf / View / ('const', off) / ('ext', slot, crop_stride) / None -> Ref."""
    if x is None:
        return null_ref()
    if isinstance(x, Buf):
        return abs_ref(x.ptr)
    if isinstance(x, View):
        return abs_ref(x.buf.ptr)
    if x[0] == 'const':
        return abs_ref(consts_ptr + x[1])
    if x[0] == 'ext':
        return ext_ref(x[1], 0, x[2])
    raise TypeError(x)


def _res(view, shift, consts_ptr, bs0=0):
    return Res(_ref(view, consts_ptr), view.buf.C, view.co, view.buf.H, view.buf.W, shift, bs0)


def emit(builder, plan):
    """Second pass: translate the symbolic ops into rsg_plan_add_* calls."""
    L = _lib.lib()
    cp = builder.consts.data_ptr()
    aux_started = False
    for kind, p, aux in builder.ops:
        if aux and not aux_started:
            _lib.check(L.rsg_plan_begin_aux(plan))
            aux_started = True
        if kind == 'conv':
            d = ConvDesc()
            src = p['src']
            d.inp = _ref(src, cp)
            d.in_cs, d.in_co, d.Hin, d.Win, d.Cin = src.buf.C, src.co, src.H, src.W, p['cin']
            d.w, d.bias = _ref(p['w'], cp), _ref(p['bias'], cp)
            d.w_tc5 = _ref(p['w_tc5'], cp)
            d.engine = int(p.get('engine', 0))
            d.Cout, d.CoutPad = p['cout'], p['cout_pad']
            d.ntaps = len(p['taps'])
            for t, (dy, dx) in enumerate(p['taps']):
                d.tap_dy[t], d.tap_dx[t] = dy, dx
            d.stride, d.Hout, d.Wout = p['stride'], p['Hout'], p['Wout']
            d.omul, d.ooy, d.oox = p['omul'], p['ooy'], p['oox']
            d.oH, d.oW = p['Hout'] * p['omul'], p['Wout'] * p['omul']
            dst = p['dst']
            if dst is not None:
                d.out = _ref(dst, cp)
                d.out_cs, d.out_co = dst.buf.C, dst.co
                assert (dst.buf.H, dst.buf.W) == (d.oH, d.oW), p['name']
            else:
                d.out = null_ref()
            d.out_f32 = _ref(p['out_f32'], cp)
            d.nres = len(p['res'])
            for q, (view, shift) in enumerate(p['res']):
                d.res[q] = _res(view, shift, cp)
            d.relu = int(p['relu'])
            d.pixel_shuffle_c = int(p.get('psc', 0))
            _lib.check(L.rsg_plan_add_conv(plan, C.byref(d)))
        elif kind == 'stem':
            _lib.check(L.rsg_plan_add_stem(plan, _ref(p['x'], cp), p['H'], p['W'], _ref(p['w'], cp),
                                           _ref(p['bias'], cp), _ref(p['out'], cp)))
        elif kind == 'fuse':
            terms = (Res * len(p['terms']))(*[_res(v, s, cp) for v, s in p['terms']])
            dst = p['dst']
            _lib.check(L.rsg_plan_add_fuse(plan, len(p['terms']), terms, _ref(dst, cp), dst.buf.C,
                                           dst.co, dst.H, dst.W, dst.C, int(p['relu'])))
        elif kind == 'maxpool':
            src = p['src']
            _lib.check(L.rsg_plan_add_maxpool(plan, _ref(src, cp), src.buf.C, src.co, src.H, src.W,
                                              src.C, _ref(p['dst'], cp)))
        elif kind == 'attention':
            x, g, y = p['x'], p['g'], p['y']
            if p.get('y32') is not None:
                _lib.check(L.rsg_plan_add_attention_f32(plan, _ref(x, cp), x.buf.C, x.co, _ref(g, cp), g.buf.C, g.co,
                                                        _ref(y, cp), y.buf.C, y.co, _ref(p['y32'], cp), x.H * x.W, x.C))
            else:
                _lib.check(L.rsg_plan_add_attention(plan, _ref(x, cp), x.buf.C, x.co, _ref(g, cp),
                                                    g.buf.C, g.co, _ref(y, cp), y.buf.C, y.co,
                                                    x.H * x.W, x.C))
        elif kind == 'trptail':
            y32, o = p['y32'], p['out']
            _lib.check(L.rsg_plan_add_trp_tail(plan, _ref(y32, cp), _ref(p['w'], cp), _ref(p['bias'], cp),
                                               _ref(p['gamma'], cp), _ref(p['beta'], cp), p['groups'], EPS,
                                               _ref(o, cp), o.buf.C, o.co, p['S'], p['C']))
        elif kind == 'relscores':
            x = p['x']
            _lib.check(L.rsg_plan_add_relation_scores(plan, _ref(x, cp), x.buf.C, x.co, x.H * x.W,
                                                      x.C, _ref(p['out'], cp)))
        elif kind == 'groupnorm':
            x, y = p['x'], p['y']
            _lib.check(L.rsg_plan_add_groupnorm(plan, _ref(x, cp), x.buf.C, x.co,
                                                _ref(p['gamma'], cp), _ref(p['beta'], cp),
                                                p['groups'], EPS, _ref(y, cp), y.buf.C, y.co,
                                                x.H * x.W, x.C))
        elif kind == 'bblock':
            src, dst = p['src'], p['dst']
            _lib.check(L.rsg_plan_add_basic_block(plan, _ref(src, cp), src.buf.C, src.co, src.H, src.W, p['C'],
                                                  _ref(p['w1'], cp), _ref(p['b1'], cp), _ref(p['w2'], cp),
                                                  _ref(p['b2'], cp), _ref(dst, cp), dst.buf.C, dst.co))
        elif kind == 'bneck':
            src, dst, r = p['src'], p['dst'], p['res']
            _lib.check(L.rsg_plan_add_bottleneck(plan, _ref(src, cp), src.buf.C, src.co, src.H, src.W, p['Cin'],
                                                 _ref(p['w1'], cp), _ref(p['b1'], cp), _ref(p['w2'], cp),
                                                 _ref(p['b2'], cp), _ref(p['w3'], cp), _ref(p['b3'], cp),
                                                 _ref(r, cp), r.buf.C, r.co, _ref(dst, cp), dst.buf.C, dst.co))
        elif kind == 'bilinear':
            _lib.check(L.rsg_plan_add_bilinear2x(plan, _ref(p['src'], cp), _ref(p['out'], cp),
                                                 p['C'], p['H'], p['W'], int(p['sigmoid'])))
        else:
            raise ValueError(kind)


# ---------------------------------------------------------------------------------------------
# Network description (mirrors pose_rsgnet.py:921-1021 / pose_hrnet.py:428-463)
# ---------------------------------------------------------------------------------------------
def _fold(w, bn):
    s, b = bn
    return w * s[:, None, None, None], b


def _cbr(pb, P, src, name, stride=1, relu=True, dst=None, res=(), conv='0', bn='1'):
    """Sequential(conv(bias=False), BN[, ReLU]) -> one fused conv op."""
    w, b = _fold(P.sub(name)[f'{conv}.weight'] if conv else P[f'{name}.weight'], P.sub(name).bn(bn))
    if dst is None:
        Hout = (src.H - 1) // stride + 1
        Wout = (src.W - 1) // stride + 1
        dst = View(pb.buf(name, Hout, Wout, _round_up(w.shape[0], 8)))
    return pb.conv(src, w, b, stride=stride, relu=relu, dst=dst, res=res, name=P.prefix + name)


def _conv_bn(pb, P, conv, bn, src, stride=1, relu=False, res=(), dst=None):
    w, b = _fold(P[f'{conv}.weight'], P.bn(bn))
    if dst is None:
        Hout = (src.H - 1) // stride + 1
        Wout = (src.W - 1) // stride + 1
        dst = View(pb.buf(P.prefix + conv, Hout, Wout, _round_up(w.shape[0], 8)))
    return pb.conv(src, w, b, stride=stride, relu=relu, dst=dst, res=res, name=P.prefix + conv)


def _pack_tc5(w):
    """[Cout, Cin, kh, kw] -> bf16 [taps][Cin/8][Cout][8]: the tcgen05 packing with one slice of NS = Cout."""
    co, ci, kh, kw = w.shape
    wt = w.transpose(2, 3, 0, 1).reshape(kh * kw, co, ci // 8, 8)
    return _bf16_bits(np.ascontiguousarray(wt.transpose(0, 2, 1, 3)))


def _quad_perm(n):
    """Accumulator column -> output channel for epilogues that read TMEM as 16x256b blocks (conv_bneck.cu, epilogue 3):
    inside each group of 64, column 8g + 2j + e holds channel 16j + 2g + e, so that lane quad member j owns 16
    consecutive channels."""
    c = np.arange(n)
    g, j, e = (c % 64) // 8, (c % 8) // 2, c % 2
    return (c // 64) * 64 + 16 * j + 2 * g + e


def _bottleneck(pb, P, x):
    w1, b1 = _fold(P['conv1.weight'], P.bn('bn1'))
    w2, b2 = _fold(P['conv2.weight'], P.bn('bn2'))
    w3, b3 = _fold(P['conv3.weight'], P.bn('bn3'))
    planes, cout = w1.shape[0], w3.shape[0]
    if (w1.shape == (planes, x.C, 1, 1) and w2.shape == (planes, planes, 3, 3) and w3.shape == (cout, planes, 1, 1)
            and _lib.lib().rsg_bottleneck_supported(x.C, planes, cout, x.H, x.W)):
        # one fused kernel: both 64-channel intermediates stay in shared memory; the residual is x itself or the
        # output of the (separate) 1x1 downsample conv
        r = _cbr(pb, P, x, 'downsample', relu=False) if P.has('downsample.0.weight') else x
        dst = View(pb.buf(P.prefix + 'block', x.H, x.W, cout))
        f32 = lambda b: pb.const(b.astype(np.float32))
        pb.simple('bneck', dict(src=x, dst=dst, res=r, Cin=x.C, w1=pb.const(_pack_tc5(w1)), b1=f32(b1),
                                w2=pb.const(_pack_tc5(w2)), b2=f32(b2), w3=pb.const(_pack_tc5(w3[_quad_perm(cout)])), b3=f32(b3),
                                name=P.prefix + 'block(fused)'),
                  [x.buf, r.buf], [dst.buf])
        pb.flops_per_fwd += 2 * (x.C * planes + 9 * planes * planes + planes * cout) * x.H * x.W
        return dst
    y = _conv_bn(pb, P, 'conv1', 'bn1', x, relu=True)
    y = _conv_bn(pb, P, 'conv2', 'bn2', y, relu=True)
    r = _cbr(pb, P, x, 'downsample', relu=False) if P.has('downsample.0.weight') else x
    return _conv_bn(pb, P, 'conv3', 'bn3', y, relu=True, res=[(r, 0)])


def _basic(pb, P, x):
    w1, b1 = _fold(P['conv1.weight'], P.bn('bn1'))
    w2, b2 = _fold(P['conv2.weight'], P.bn('bn2'))
    Cb = w1.shape[0]
    if (w1.shape == (Cb, Cb, 3, 3) and w2.shape == (Cb, Cb, 3, 3) and x.C == Cb
            and _lib.lib().rsg_basic_block_supported(Cb, x.H, x.W)):
        # one fused kernel: the intermediate stays in shared memory, x is read once (input patch = residual)
        dst = View(pb.buf(P.prefix + 'block', x.H, x.W, Cb))

        def pack(w):        # [9][1 slice][C/8][C][8]: the tcgen05 packing with NS = C
            wt = w.transpose(2, 3, 0, 1).reshape(9, 1, Cb, Cb // 8, 8)
            return pb.const(_bf16_bits(wt.transpose(1, 0, 3, 2, 4)))
        pb.simple('bblock', dict(src=x, dst=dst, C=Cb, w1=pack(w1), b1=pb.const(b1.astype(np.float32)),
                                 w2=pack(w2), b2=pb.const(b2.astype(np.float32)), name=P.prefix + 'block(fused)'),
                  [x.buf], [dst.buf])
        pb.flops_per_fwd += 2 * 2 * 9 * Cb * Cb * x.H * x.W
        return dst
    y = _conv_bn(pb, P, 'conv1', 'bn1', x, relu=True)
    return _conv_bn(pb, P, 'conv2', 'bn2', y, relu=True, res=[(x, 0)])


def _hr_module(pb, P, xs, n_out):
    nb = len(xs)
    xs = list(xs)
    for b in range(nb):
        i = 0
        while P.has(f'branches.{b}.{i}.conv1.weight'):
            xs[b] = _basic(pb, P.sub(f'branches.{b}.{i}'), xs[b])
            i += 1
    if nb == 1:
        return xs
    outs = []
    for i in range(n_out):
        terms = []                       # (view, up-shift)
        for j in range(nb):
            if j == i:
                terms.append((xs[j], 0))
            elif j > i:
                t = _cbr(pb, P, xs[j], f'fuse_layers.{i}.{j}', relu=False)
                terms.append((t, j - i))
        down = [j for j in range(nb) if j < i]
        # all but the last down-path produce plain tensors; the last one's final conv absorbs the
        # whole sum + ReLU in its epilogue
        for j in down[:-1]:
            t = xs[j]
            for k in range(i - j):
                t = _cbr(pb, P, t, f'fuse_layers.{i}.{j}.{k}', stride=2, relu=(k != i - j - 1))
            terms.append((t, 0))
        if down:
            j = down[-1]
            t = xs[j]
            for k in range(i - j - 1):
                t = _cbr(pb, P, t, f'fuse_layers.{i}.{j}.{k}', stride=2, relu=True)
            out = _cbr(pb, P, t, f'fuse_layers.{i}.{j}.{i - j - 1}', stride=2, relu=True, res=terms)
        else:
            x0 = xs[i]
            out = View(pb.buf(P.prefix + f'fuse{i}', x0.H, x0.W, x0.buf.C))
            pb.simple('fuse', dict(terms=terms, dst=out, relu=True),
                      [t.buf for t, _ in terms], [out.buf])
        outs.append(out)
    return outs


def _transition(pb, P, prev, n_cur):
    n_pre = len(prev)
    out = []
    for i in range(n_cur):
        if i < n_pre:
            out.append(_cbr(pb, P, prev[i], f'{i}') if P.has(f'{i}.0.weight') else prev[i])
        else:
            t = prev[-1]
            for j in range(i + 1 - n_pre):
                t = _cbr(pb, P, t, f'{i}.{j}', stride=2)
            out.append(t)
    return out


def _backbone(pb, P, spec, taps):
    H, W = spec.image_h, spec.image_w
    w1, b1 = _fold(P['conv1.weight'], P.bn('bn1'))
    c1 = pb.buf('conv1', H // 2, W // 2, 64)
    wst = np.ascontiguousarray(w1.transpose(1, 2, 3, 0).reshape(27, 64), np.float32)
    pb.simple('stem', dict(x=('ext', EXT_X, 0), H=H, W=W, w=pb.const(wst),
                           bias=pb.const(b1.astype(np.float32)), out=c1), [], [c1])
    pb.flops_per_fwd += 2 * 27 * 64 * (H // 2) * (W // 2)
    x = _conv_bn(pb, P, 'conv2', 'bn2', View(c1), stride=2, relu=True)
    taps['stem'] = x
    for i in range(4):
        x = _bottleneck(pb, P.sub(f'layer1.{i}'), x)
    taps['layer1'] = x
    ys = [x]
    for si, st in enumerate(spec.stages):
        s = si + 2
        xs = _transition(pb, P.sub(f'transition{s - 1}'), ys, st.num_branches)
        for m in range(st.num_modules):
            last = (si == len(spec.stages) - 1 and m == st.num_modules - 1)
            xs = _hr_module(pb, P.sub(f'stage{s}.{m}'), xs, 1 if last else st.num_branches)
        ys = xs
        for b, t in enumerate(ys):
            taps[f'stage{s}.{b}'] = t
    return ys


def _deconv4(pb, P, name, src, dst):
    """Sequential(ConvTranspose2d(k4,s2,p1,bias=False), BN, ReLU) as four 2x2 phase convs."""
    w = P[f'{name}.0.weight']                       # [Cin, Cout, 4, 4]
    s, b = P.sub(name).bn('1')
    if w.shape[2] != 4:
        raise ValueError('only FINAL_DECONV_KERNEL_SIZE=4 is supported')
    cin, cout = w.shape[0], w.shape[1]
    ns, kc, st = C.c_int(), C.c_int(), C.c_int()
    if (cout % 16 == 0 and cin % 16 == 0 and dst.co == 0 and dst.C == cout and
            _lib.lib().rsg_conv_tc5_config(cin, _round_up(4 * cout, 32), 9, 1, C.byref(ns), C.byref(kc), C.byref(st))):
        # ONE 3x3 conv with 4*Cout output columns (phase-major; taps a phase does not use are zero) whose
        # epilogue pixel-shuffles: the input is read once instead of four times
        taps9 = [(dy, dx) for dy in (-1, 0, 1) for dx in (-1, 0, 1)]
        wf = np.zeros((9, 4 * cout, cin))
        for py in (0, 1):
            ys = [(1, 0), (3, -1)] if py == 0 else [(0, 1), (2, 0)]
            for px in (0, 1):
                xs = [(1, 0), (3, -1)] if px == 0 else [(0, 1), (2, 0)]
                ph = 2 * py + px
                for ky, dy in ys:
                    for kx, dx in xs:
                        wf[taps9.index((dy, dx)), ph * cout:(ph + 1) * cout, :] = (w[:, :, ky, kx] * s[None, :]).T
        pb.conv(src, wf, np.tile(b, 4), relu=True, dst=dst, taps=taps9, omul=2, out_hw=(src.H, src.W),
                name=P.prefix + name + '.fused', pixel_shuffle_c=cout)
        pb.flops_per_fwd -= 2 * (9 * 4 - 16) * cin * cout * src.H * src.W      # zero taps are not credited
        return dst
    for py in (0, 1):
        ys = [(1, 0), (3, -1)] if py == 0 else [(0, 1), (2, 0)]
        for px in (0, 1):
            xs = [(1, 0), (3, -1)] if px == 0 else [(0, 1), (2, 0)]
            taps, mats = [], []
            for ky, dy in ys:
                for kx, dx in xs:
                    taps.append((dy, dx))
                    mats.append((w[:, :, ky, kx] * s[None, :]).T)      # [Cout, Cin]
            pb.conv(src, np.stack(mats), b, relu=True, dst=dst, taps=taps, omul=2, ooy=py, oox=px,
                    out_hw=(src.H, src.W), name=P.prefix + name + f'.p{py}{px}')
    return dst


def build_network(pb, sd, spec):
    """Returns a dict describing the external outputs."""
    P = _Params(sd)
    taps = {}
    ys = _backbone(pb, P, spec, taps)
    feat = ys[0]
    K = spec.num_joints
    h, w = spec.feat_h, spec.feat_w
    info = dict(K=K, heat_h=spec.heat_h, heat_w=spec.heat_w, taps=taps)

    def head1x1(name, src, out_f32, cout_w=None):
        wt, bs = P[f'{name}.weight'], P[f'{name}.bias']
        if wt.shape[-1] != 1:
            raise ValueError('FINAL_CONV_KERNEL must be 1')
        return wt, bs

    if spec.kind != KIND_RSGNET:
        wt, bs = head1x1('final_layer', feat, None)
        pb.conv(feat, wt, bs, out_f32=('ext', EXT_HEAT, K * h * w * 4), name='final_layer')
        return info

    C0 = spec.head_channels
    L = spec.num_limbs
    up = spec.up_scale
    if up not in (1, 2):
        raise ValueError('UP_SCALE must be 1 or 2')
    # multi_final_layer -> bf16 scores (padded to a multiple of 8 channels) for the type conv
    wm, bm = head1x1('multi_final_layer', feat, None)
    K8 = _round_up(K, 8)
    multi = View(pb.buf('multi_scores', h, w, K8))
    pb.conv(feat, wm, bm, dst=multi, cout=K8, name='multi_final_layer')

    # cat(vis, type, loc) lives in one persistent buffer; the loc slice is a pack-time constant
    cat1 = pb.buf('cat_vis_type_loc', h, w, 3 * C0, persistent=True)
    taps['vis'] = _cbr(pb, P, feat, 'vis_conv', dst=View(cat1, 0, C0))
    tf = P['type_features'] @ P['type_fc.0.weight'].T
    s1, b1 = P.sub('type_fc').bn('1')
    T = np.maximum(tf * s1[None, :] + b1[None, :], 0.0)                 # [K, 600]
    wc = P['type_conv.0.weight']                                       # [C0, 600, 3, 3]
    wfold = np.einsum('otyx,kt->okyx', wc, T)
    s2, b2 = P.sub('type_conv').bn('1')
    taps['type'] = pb.conv(multi, wfold * s2[:, None, None, None], b2, relu=True,
                           dst=View(cat1, C0, C0), name='type_conv(folded)')
    wl = P['loc_conv.0.weight'][:, :, 0, 0]                            # [C0, 4]
    s3, b3 = P.sub('loc_conv').bn('1')
    loc = np.einsum('oi,iyx->oyx', wl, P['loc_features'][0]) * s3[:, None, None] + b3[:, None, None]
    info['loc_const'] = (cat1, np.maximum(loc, 0.0).transpose(1, 2, 0).astype(np.float32), 2 * C0)

    fvc = _cbr(pb, P, View(cat1), 'contact_conv')
    cat2 = pb.buf('cat_rel_vis', h, w, 2 * C0)
    fv = _cbr(pb, P, fvc, 'predict_contact_net', dst=View(cat2, C0, C0))
    taps['final_vis'] = fv
    taps['relation'] = View(cat2, 0, C0)

    # ---- TRP (association.py:280-301)
    R = P.sub('relation_head')
    xs = fv
    if spec.relation_sub_sample:
        pooled = pb.buf('trp_pool', h // 2, w // 2, C0)
        pb.simple('maxpool', dict(src=fv, dst=pooled), [fv.buf], [pooled])
        xs = View(pooled)
    g = View(pb.buf('trp_g', xs.H, xs.W, C0))
    pb.conv(xs, R['g.weight'], R['g.bias'], dst=g, name='relation_head.g')
    yv = View(pb.buf('trp_y', xs.H, xs.W, C0))
    S = xs.H * xs.W
    pb.flops_per_fwd += 2 * 2 * S * S * C0
    info['S'] = S
    info['trp_x'] = xs
    f32 = lambda a: pb.const(np.ascontiguousarray(a, np.float32))
    if spec.relation_sub_sample:
        pb.simple('attention', dict(x=xs, g=g, y=yv), [xs.buf, g.buf], [yv.buf])
        yu = View(pb.buf('trp_up', h, w, C0))
        _deconv4(pb, R, 'W.0', yv, yu)                  # ConvT + BN + ReLU in between: the bf16 conv path
        Wt = R.sub('W.1')
        z = View(pb.buf('trp_z', h, w, C0))
        pb.conv(yu, Wt['0.weight'], Wt['0.bias'], dst=z, name='relation_head.W')
        pb.simple('groupnorm', dict(x=z, y=View(cat2, 0, C0), groups=8, gamma=f32(Wt['1.weight']), beta=f32(Wt['1.bias'])),
                  [z.buf], [cat2])
    else:
        # y -> 1x1 W -> GroupNorm stays in fp32: y's mean over the positions dwarfs its spread and GroupNorm removes it
        Wt = R.sub('W')
        y32 = pb.buf('trp_y32', xs.H, xs.W, C0, itemsize=4)
        pb.simple('attention', dict(x=xs, g=g, y=yv, y32=y32), [xs.buf, g.buf], [yv.buf, y32])
        pb.simple('trptail', dict(y32=y32, out=View(cat2, 0, C0), w=f32(Wt['0.weight'][:, :, 0, 0]), bias=f32(Wt['0.bias']),
                                  gamma=f32(Wt['1.weight']), beta=f32(Wt['1.bias']), groups=8, S=S, C=C0,
                                  name='relation_head.W + GroupNorm (fp32)'), [y32], [cat2])
        pb.flops_per_fwd += 2 * 2 * C0 * C0 * S          # z is computed twice (statistics pass + output pass), on CUDA cores

    # ---- keypoint head (pose_rsgnet.py:991-1000)
    kf = _cbr(pb, P, View(cat2), 'kpt_net')
    taps['kpt_net'] = kf
    if up > 1:
        kd = View(pb.buf('kpt_up', h * 2, w * 2, C0))
        kf = _deconv4(pb, P, 'predict_convtranspose', kf, kd)
    kf = _cbr(pb, P, kf, 'predict_net')
    taps['kpt_feat'] = kf
    wf, bf_ = head1x1('final_layer', kf, None)
    pb.conv(kf, wf, bf_, out_f32=('ext', EXT_HEAT, K * spec.heat_h * spec.heat_w * 4),
            name='final_layer')

    # ---- auxiliary outputs (unused by the eval loop, function.py:389): multi_kpt_scores,
    # limbs_scores (SGM, pose_rsgnet.py:1003-1013), relation_scores
    pb.aux = True
    if up > 1:
        m32 = pb.buf('multi_f32', h, w, K, itemsize=4)
        pb.conv(feat, wm, bm, out_f32=m32, name='multi_final_layer(aux)')
        pb.simple('bilinear', dict(src=m32, out=('ext', EXT_MULTI, K * 4 * h * w * 4), C=K, H=h, W=w,
                                   sigmoid=False), [m32], [])
    else:
        pb.conv(feat, wm, bm, out_f32=('ext', EXT_MULTI, K * h * w * 4), name='multi_final_layer(aux)')
    lf = _cbr(pb, P, fv, 'limbs_net')
    KT = P.sub('kt_machine')
    mm = KT['matrix_limb'] * KT['real_matrix_limb']
    t = mm @ wf.reshape(K, -1)
    t = t @ KT['kpt_transformer.0.weight'].T + KT['kpt_transformer.0.bias']
    t = np.where(t > 0, t, 0.02 * t)
    t = t @ KT['kpt_transformer.2.weight'].T + KT['kpt_transformer.2.bias']
    refine = t.reshape(L, C0, 1, 1)
    if up > 1:
        l32 = pb.buf('limbs_f32', h, w, L, itemsize=4)
        pb.conv(lf, refine, np.zeros(L), out_f32=l32, name='limbs(dynamic 1x1)')
        pb.simple('bilinear', dict(src=l32, out=('ext', EXT_LIMBS, L * 4 * h * w * 4), C=L, H=h, W=w,
                                   sigmoid=True), [l32], [])
    else:
        raise ValueError('UP_SCALE=1 RSGNet heads are not supported')
    pb.simple('relscores', dict(x=xs, out=('ext', EXT_REL, S * S * 4)), [xs.buf], [])
    info['L'] = L
    return info


# ---------------------------------------------------------------------------------------------
class Engine:
    """Per (module, device) packed network + plan.  Rebuilt when parameters change."""

    def __init__(self, module, device, chunk=32, reuse=True):
        _lib.require_cuda()
        self.device = torch.device(device)
        self.spec = module.spec
        self.chunk = chunk
        sd = {k: v for k, v in module.state_dict().items()}
        with torch.cuda.device(self.device):
            pb = PlanBuilder(chunk, reuse)
            self.info = build_network(pb, sd, self.spec)
            pb.allocate(self.device)
            if 'loc_const' in self.info:
                buf, loc, co = self.info.pop('loc_const')
                t = buf.tensor.view(torch.bfloat16).view(chunk, buf.H, buf.W, buf.C)
                t[..., co:co + loc.shape[-1]] = torch.from_numpy(loc).to(self.device, torch.bfloat16)
            self.pb = pb
            handle = C.c_void_p()
            _lib.check(_lib.lib().rsg_plan_create(C.byref(handle), chunk))
            self.plan = handle
            emit(pb, self.plan)
        self.flops_per_fwd = pb.flops_per_fwd
        self.lock = threading.Lock()
        # the activation arena, the concat buffers and the graph cache belong to this engine: runs issued from different
        # CUDA streams are ordered against each other with an event (see run)
        self._last_stream = None
        self._events = [torch.cuda.Event(), torch.cuda.Event()]
        self._ev_i = 0
        self._static = {}        # module_forward's static input / output buffers, keyed by with_aux

    def __del__(self):
        try:
            if getattr(self, 'plan', None):
                _lib.lib().rsg_plan_destroy(self.plan)
                self.plan = None
        except Exception:
            pass

    def out_shapes(self, n_fwd):
        s, i = self.spec, self.info
        shapes = {EXT_HEAT: (n_fwd, i['K'], s.heat_h, s.heat_w)}
        if s.kind == KIND_RSGNET:
            shapes[EXT_MULTI] = (n_fwd, i['K'], s.heat_h, s.heat_w)
            shapes[EXT_LIMBS] = (n_fwd, i['L'], s.heat_h, s.heat_w)
            shapes[EXT_REL] = (n_fwd, i['S'], i['S'])
        return shapes

    def run(self, x, heat, n_fwd, n_crops, aux=None, use_graph=False):
        """x: f32 CUDA [n_crops,3,H,W] contiguous; heat: f32 CUDA [n_fwd,K,Hh,Wh]; aux: dict slot->tensor."""
        ext = (C.c_void_p * N_EXT)()
        ext[EXT_X] = x.data_ptr()
        ext[EXT_HEAT] = heat.data_ptr()
        with_aux = 0
        if aux:
            with_aux = 1
            for slot, t in aux.items():
                ext[slot] = t.data_ptr()
        with self.lock, torch.cuda.device(self.device):
            st = torch.cuda.current_stream(self.device)
            if self._last_stream is not None and self._last_stream != st:
                # another stream ran this engine last: its work must finish before the arena is overwritten
                st.wait_event(self._events[self._ev_i])
            _lib.check(_lib.lib().rsg_plan_run(self.plan, _lib.stream_ptr(self.device), ext, N_EXT,
                                               n_fwd, n_crops, with_aux, int(use_graph)))
            self._ev_i ^= 1
            self._events[self._ev_i].record(st)
            self._last_stream = st

    def static_buffers(self, with_aux):
        """Persistent input / output tensors of module_forward (pointer-stable, so the whole forward replays as one
        CUDA graph): x [chunk,3,H,W], heat, and with_aux the three auxiliary outputs."""
        key = bool(with_aux)
        if key not in self._static:
            sp, dev = self.spec, self.device
            shapes = self.out_shapes(self.chunk)
            base = self._static.get(False)
            bufs = dict(base) if (key and base) else dict(
                x=torch.empty((self.chunk, 3, sp.image_h, sp.image_w), dtype=torch.float32, device=dev),
                heat=torch.empty(shapes[EXT_HEAT], dtype=torch.float32, device=dev))
            if key:
                for slot in (EXT_MULTI, EXT_LIMBS, EXT_REL):
                    bufs[slot] = torch.empty(shapes[slot], dtype=torch.float32, device=dev)
            self._static[key] = bufs
        return self._static[key]

    def profile(self, x, heat, nb, n_crops):
        """Per-op device times of one chunk (eager, event pair around every op).
        Returns (ms f32[n_ops], kind i32[n_ops], flops f64[n_ops], names)."""
        L = _lib.lib()
        n = L.rsg_plan_num_ops(self.plan)
        ms = np.zeros(n, np.float32)
        kind = np.zeros(n, np.int32)
        flops = np.zeros(n, np.float64)
        ext = (C.c_void_p * N_EXT)()
        ext[EXT_X] = x.data_ptr()
        ext[EXT_HEAT] = heat.data_ptr()
        with self.lock, torch.cuda.device(self.device):
            _lib.check(L.rsg_plan_profile(self.plan, _lib.stream_ptr(self.device), ext, N_EXT, nb,
                                          n_crops, 0, ms.ctypes.data, kind.ctypes.data,
                                          flops.ctypes.data))
        names, shapes, nbytes = [], [], []
        for k, p, _ in self.pb.ops:
            names.append(p.get('name', k) if isinstance(p, dict) else k)
            if k == 'conv':
                shapes.append(f"conv {p['cin']}->{p['cout']} taps{len(p['taps'])} s{p['stride']} "
                              f"{p['Hout']}x{p['Wout']}{' +res' * len(p['res'])}")
                # algorithmic bytes per forward: input read once, output written once, every residual term read
                # once at its own resolution (weights are negligible and L2-resident)
                src = p['src']
                by = src.H * src.W * p['cin'] * 2
                by += p['Hout'] * p['Wout'] * p['omul'] ** 2 * p['cout'] * (4 if p['out_f32'] is not None else 2) \
                    // (4 if p.get('psc') else 1)
                by += sum(v.H * v.W * p['cout'] * 2 for v, _ in p['res'])
            else:
                shapes.append(k)
                if k == 'stem':
                    by = 3 * p['H'] * p['W'] * 4 + (p['H'] // 2) * (p['W'] // 2) * 64 * 2
                elif k == 'fuse':
                    by = sum(v.H * v.W * p['dst'].C * 2 for v, _ in p['terms']) + p['dst'].H * p['dst'].W * p['dst'].C * 2
                elif k == 'bblock':
                    by = 2 * p['src'].H * p['src'].W * p['C'] * 2
                elif k == 'bneck':       # x read once, residual read once, output written once
                    by = p['src'].H * p['src'].W * (p['Cin'] + 256 + 256) * 2
                elif k == 'attention':
                    by = 3 * p['x'].H * p['x'].W * p['x'].C * 2
                elif k == 'groupnorm':
                    by = 2 * p['x'].H * p['x'].W * p['x'].C * 2
                else:
                    by = 0
            nbytes.append(float(by) * nb)
        self.last_profile_bytes = np.asarray(nbytes)
        return ms, kind, flops, names, shapes

    def last_launches(self):
        return _lib.lib().rsg_plan_last_launches(self.plan)


# ---------------------------------------------------------------------------------------------
_engines_lock = threading.Lock()
_CHUNKS = (32, 64, 128, 256, 512)


def _source(module):
    """The module that owns the real parameters: a DataParallel replica (torch.nn.parallel.replicate gives it an EMPTY
    _parameters dict and per-forward broadcast copies) points back at the module it was replicated from."""
    return module.__dict__.get('_rsg_source') or module


def _param_version(module):
    ts = module.__dict__.get('_rsg_tensors')
    if ts is None:
        ts = module.__dict__['_rsg_tensors'] = list(module.parameters()) + list(module.buffers())
    return sum(t._version for t in ts), len(ts)


def engine_for(module, device, chunk=None):
    """Engine cache on the parameter-owning module, keyed by (device, chunk); invalidated when any parameter / buffer
    changes (load_state_dict, .to(), train() drop the cache through EngineOwner; in-place edits bump the tensors'
    version counters).  Safe to call from DataParallel's per-device threads."""
    module = _source(module)
    chunk = int(chunk or getattr(module, 'chunk', 32))
    key = (str(torch.device(device)), chunk)
    with _engines_lock:
        cache = module.__dict__.setdefault('_rsg_engines', {})
        ver = _param_version(module)
        ent = cache.get(key)
        if ent is None or ent[0] != ver:
            for k in [k for k, e in cache.items() if e[0] != ver]:
                del cache[k]
            cache[key] = ent = (ver, Engine(module, device, chunk))
        return ent[1]


class LazyOutput:
    """An output of RSGNet.forward that the eval loop never reads (function.py:389 uses outputs[1] only), computed by
    a second run of the plan WITH the auxiliary ops when it is first touched (``model.lazy_aux = True``).  Behaves like
    the tensor it stands for: attribute access, indexing and torch.* functions materialise it."""

    def __init__(self, group, slot, shape, device):
        self._group, self._slot, self._shape, self._device = group, slot, tuple(shape), device

    def materialize(self):
        return self._group.get(self._slot)

    @property
    def shape(self):
        return torch.Size(self._shape)

    @property
    def device(self):
        return self._device

    @property
    def dtype(self):
        return torch.float32

    def size(self, *a):
        return torch.Size(self._shape) if not a else self._shape[a[0]]

    def dim(self):
        return len(self._shape)

    def __len__(self):
        return self._shape[0]

    def __getitem__(self, idx):
        return self.materialize()[idx]

    def __getattr__(self, name):
        if name.startswith('__'):
            raise AttributeError(name)
        return getattr(self.materialize(), name)

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        un = lambda a: a.materialize() if isinstance(a, LazyOutput) else a
        args = tuple([un(v) for v in a] if isinstance(a, (list, tuple)) else un(a) for a in args)
        return func(*args, **{k: un(v) for k, v in (kwargs or {}).items()})


class _LazyAuxGroup:
    """Shared state of the three lazy outputs of one forward: the (cloned) input and, once touched, the tensors."""

    def __init__(self, module, x, relation_target):
        self.module, self.x, self.target, self.out = module, x, relation_target, None

    def get(self, slot):
        if self.out is None:
            full = _forward_impl(self.module, self.x, self.target, with_aux=True)
            self.out = {EXT_MULTI: full[0], EXT_LIMBS: full[2], EXT_REL: full[3]}
            self.x = None
        return self.out[slot]


def _forward_impl(module, x, relation_target, with_aux):
    src = _source(module)
    spec = src.spec
    if x.is_cuda:
        dev = x.device
    else:
        dev = module.__dict__.get('_rsg_device') or next(src.parameters()).device
    if dev.type != 'cuda':
        raise _lib.RsgError('move the model to a CUDA device first (model.cuda()); no CPU path')
    B = int(x.shape[0])
    chunk = next((c for c in _CHUNKS if c >= B), _CHUNKS[-1])
    eng = engine_for(src, dev, chunk)
    rsg = spec.kind == KIND_RSGNET
    outs = []
    with torch.cuda.device(dev):
        st = torch.cuda.current_stream(dev)
        graph_ok = st.cuda_stream != 0            # capture needs a non-default stream; else the same plan runs eagerly
        for lo in range(0, B, chunk):
            nb = min(chunk, B - lo)
            bufs = eng.static_buffers(with_aux)
            bufs['x'][:nb].copy_(x[lo:lo + nb], non_blocking=True)       # H2D for DataParallel's CPU scatter, else D2D
            aux = {s: bufs[s] for s in (EXT_MULTI, EXT_LIMBS, EXT_REL)} if (rsg and with_aux) else None
            eng.run(bufs['x'], bufs['heat'], nb, nb, aux=aux, use_graph=graph_ok)
            # results leave the static buffers as fresh tensors (the caller keeps `output` across the flipped forward)
            got = {EXT_HEAT: bufs['heat'][:nb].clone()}
            if aux:
                for s, t in aux.items():
                    got[s] = t[:nb].clone()
            outs.append(got)
    cat = lambda s: outs[0][s] if len(outs) == 1 else torch.cat([o[s] for o in outs])
    heat = cat(EXT_HEAT)
    if not rsg:
        return heat
    if not with_aux:
        return None, heat, None, None
    rel = cat(EXT_REL)
    if relation_target is not None:
        rel = ((relation_target.to(dev) - rel) ** 2).mean(dim=(1, 2))
    return cat(EXT_MULTI), heat, cat(EXT_LIMBS), rel


def module_forward(module, x, relation_target=None):
    """nn.Module.forward of the drop-in models: the reference's return convention (pose_rsgnet.py:955-1021,
    pose_hrnet.py:428-463).  The input is copied into a pointer-stable buffer and the whole forward replays as one CUDA
    graph per (batch size, device); DataParallel replicas build their engines from the source module's parameters."""
    src = _source(module)
    if module.training or src.training:
        # train mode: batch-statistics BatchNorm + a differentiable output (lib/core/function.py:271 calls
        # model(input, relation_target) and then loss.backward()): rsgnet_b200/train, not the folded inference plan
        if src is not module:
            raise _lib.RsgError('training mode under nn.DataParallel replicas is not supported: run one process per GPU '
                                '(rsgnet_b200.train.TrainStep all-reduces the gradients over NCCL)')
        from .train.step import module_forward_train
        return module_forward_train(module, x, relation_target)
    _lib.require_cuda()
    spec = src.spec
    if x.dim() != 4 or x.shape[1] != 3 or x.shape[2] != spec.image_h or x.shape[3] != spec.image_w:
        raise ValueError(f'expected input [B,3,{spec.image_h},{spec.image_w}], got {tuple(x.shape)}')
    if x.dtype != torch.float32:
        x = x.float()
    if spec.kind != KIND_RSGNET:
        return _forward_impl(module, x, None, False)
    if getattr(src, 'lazy_aux', False) and relation_target is None:
        _, heat, _, _ = _forward_impl(module, x, None, False)
        B, i = int(x.shape[0]), engine_for(src, heat.device, next((c for c in _CHUNKS if c >= x.shape[0]), _CHUNKS[-1])).info
        group = _LazyAuxGroup(module, x.detach().clone(), None)
        shp = {EXT_MULTI: (B, i['K'], spec.heat_h, spec.heat_w), EXT_LIMBS: (B, i['L'], spec.heat_h, spec.heat_w),
               EXT_REL: (B, i['S'], i['S'])}
        lazy = {s: LazyOutput(group, s, shp[s], heat.device) for s in shp}
        return lazy[EXT_MULTI], heat, lazy[EXT_LIMBS], lazy[EXT_REL]
    return _forward_impl(module, x, relation_target, True)
