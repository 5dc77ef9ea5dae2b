"""A minimal reverse-mode tape over the library's training kernels (include/rsg_b200.h, 'Training step').

The reference gets its backward pass from torch autograd over cuDNN / cuBLAS (lib/core/function.py:240-363 calls
``loss.backward()``).  Here every operation of the train-mode forward (lib/models/pose_rsgnet.py:955-1021) is one or a few
calls into librsg_b200.so, and records a closure that issues the matching backward kernels.  torch supplies device memory
(``torch.empty``), streams and events; the only torch kernels on this path are bookkeeping: the ``num_batches_tracked``
counters (one ``_foreach_add_``), input copies into the graph's static buffers, and the gradient snapshot of the drop-in route.

Activations are fp32 NHWC tensors ``[N, H, W, C]`` (or plain matrices ``[M, C]``).
"""
import ctypes as C

import torch

from .. import _lib


class T:
    """A tape value: ``v`` the tensor, ``g`` its gradient (None until something flows back)."""
    __slots__ = ('v', 'g', 'req', 'own', 'leaf')

    def __init__(self, v, req=True, g=None, leaf=False):
        self.v = v
        self.g = g
        self.req = req
        self.leaf = leaf              # a parameter: nothing in the backward pass reads its gradient
        self.own = g is not None      # whether g may be updated in place (False when the tensor is shared with another node)

    @property
    def shape(self):
        return tuple(self.v.shape)


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


_SIDE_STREAMS = {}


def side_streams(device, n=3):
    """Process-wide side streams of a device for Tape.parallel (created once, outside any graph capture)."""
    key = (device.index if device.index is not None else torch.cuda.current_device())
    st = _SIDE_STREAMS.get(key)
    if st is None or len(st) < n:
        with torch.cuda.device(device):
            st = [torch.cuda.Stream(device) for _ in range(n)]
        _SIDE_STREAMS[key] = st
    return st


class Tape:
    def __init__(self, device, precise=False, side=None, wstream=None):
        self.device = device
        self.lib = _lib.lib()
        self.ops = []
        self.default_precise = int(precise)
        self.precise = int(precise)       # 0 TF32 (tcgen05 kernel where it applies), 1 3xTF32 (fp32-class), 2 TF32 on mma.sync
        self.launches = 0
        self.flops = 0.0              # executed multiply-adds * 2 of the matrix kernels (3x in the 3xTF32 mode not counted)
        self.prof = None              # list of (kernel, start event, end event, flops) when profiling
        self.side = side or []         # side streams for parallel(); empty = everything on the current stream
        self.wstream = wstream         # weight gradients run here, off the critical path (None = on the current stream)
        self._w_keep, self._w_event = [], None
        self._ws = {}                  # per-stream reduction scratch (doubles), kept ZERO between calls (rsg_b200.h, BatchNorm)
        self.ws(4096)

    # ------------------------------------------------------------------ plumbing
    @property
    def st(self):
        return _lib.stream_ptr(self.device)

    def new(self, *shape):
        return torch.empty(shape, dtype=torch.float32, device=self.device)

    def call(self, name, *args, n=1, flops=0.0, tag=''):
        if self.prof is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _lib.check(getattr(self.lib, name)(self.st, *args))
            e1.record()
            self.prof.append((name + tag, e0, e1, flops))
        else:
            _lib.check(getattr(self.lib, name)(self.st, *args))
        self.launches += n
        self.flops += flops

    def record(self, fn):
        self.ops.append(fn)

    def backward(self):
        while self.ops:
            self.ops.pop()()
        self.join_wgrad()

    def acc(self, t, g, shared=False):
        """Add gradient tensor g into node t.  `shared`: g is also handed to another node (must not be modified)."""
        if not t.req:
            return
        if t.g is None:
            t.g = g
            t.own = not shared
        elif t.own:
            self.call('rsg_train_ew', 6, _p(g), None, 0.0, g.numel(), _p(t.g))
        else:
            out = self.new(*t.g.shape)
            ptrs = (C.c_void_p * 2)(t.g.data_ptr(), g.data_ptr())
            self.call('rsg_train_add', 2, ptrs, 0, g.numel(), _p(out))
            t.g = out
            t.own = True

    def grad_buf(self, t):
        """The gradient buffer of node t for kernels that ACCUMULATE into it (wgrad atomics): parameters own a persistent
        one; computed weights (the type vectors, the KTMachine's limb filters) get a zeroed buffer on first use."""
        if t.g is None:
            t.g = self.new(*t.v.shape)
            self.call('rsg_train_zero', _p(t.g), C.c_size_t(4 * t.g.numel()))
            t.own = True
        elif not t.own:
            g = self.new(*t.g.shape)
            self.call('rsg_train_copy2d', _p(t.g), 0, _p(g), 0, 1, t.g.numel(), 0)
            t.g, t.own = g, True
        return t.g

    def ws(self, n):
        key = torch.cuda.current_stream(self.device).cuda_stream
        buf = self._ws.get(key)
        if buf is None or buf.numel() < n:
            buf = torch.empty(max(n, 4096), dtype=torch.float64, device=self.device)
            self.call('rsg_train_zero', _p(buf), C.c_size_t(8 * buf.numel()))
            self._ws[key] = buf
        return buf

    def parallel(self, fns):
        """Run the callables concurrently: fns[0] on the current stream, fns[i] on side stream i - 1 (fork / join with events;
        captured into a CUDA graph these become parallel branches).  Each callable records its own sub-tape; ONE composite
        closure on the main tape replays the sub-tapes backwards with the same fork / join, so the backward passes of the
        branches overlap too.  The HRNet branches of a HighResolutionModule (pose_rsgnet.py:254-260) are independent chains of
        BasicBlocks, and on the 16x12 and 8x6 maps their kernels fill a fraction of the SMs.
        Memory safety without record_stream: every tensor that crosses streams is produced before a fork or consumed after a
        join, and torch's caching allocator reuses a freed block only on the stream it was allocated on."""
        if not self.side or len(fns) < 2:
            return [fn() for fn in fns]
        dev = self.device
        main = torch.cuda.current_stream(dev)
        n = min(len(fns), len(self.side) + 1)
        fork = torch.cuda.Event()
        fork.record(main)
        saved = self.ops
        results, recs = [], []
        for i, fn in enumerate(fns):
            st = main if (i == 0 or i >= n) else self.side[i - 1]
            self.ops = []
            if st is not main:
                st.wait_event(fork)
            with torch.cuda.stream(st):
                results.append(fn())
            recs.append((st is not main, st, self.ops))
        self.ops = saved
        for on_side, st, _ in recs:
            if on_side:
                ev = torch.cuda.Event()
                ev.record(st)
                main.wait_event(ev)

        def bwd():
            cur = torch.cuda.current_stream(dev)
            fk = torch.cuda.Event()
            fk.record(cur)
            for on_side, st, ops in recs:
                run_on = st if on_side else cur
                if on_side:
                    run_on.wait_event(fk)
                with torch.cuda.stream(run_on):
                    while ops:
                        ops.pop()()
            for on_side, st, _ in recs:
                if on_side:
                    ev = torch.cuda.Event()
                    ev.record(st)
                    cur.wait_event(ev)
        self.record(bwd)
        return results

    # ------------------------------------------------------------------ matrix ops
    def _gemm(self, pr, A, B, Cm, bias, M, Nc, Ca, lda, ldb, ldc, mode=0, transA=0, transB=0, beta=0, geom=None, batch=1,
              sA=0, sB=0, sC=0):
        g = (C.c_int * 8)(*geom) if geom is not None else None
        taps = geom[4] * geom[5] if geom is not None else 1
        self.call('rsg_train_gemm', _p(A), _p(B), _p(Cm), _p(bias), M, Nc, Ca, lda, ldb, ldc, batch, sA, sB, sC, mode,
                  transA, transB, beta, g, pr, flops=2.0 * M * Nc * Ca * taps * batch,
                  tag=f' M{M} N{Nc} K{Ca} taps{taps} mode{mode} s{geom[6] if geom is not None else 1} b{batch} tA{transA} tB{transB} p{pr}' if self.prof is not None else '')

    def _wgrad(self, pr, X, dY, dW, M, Ca, Nc, mode=0, geom=None, leaf=False):
        """Weight gradients of PARAMETERS are leaves of the backward pass: nothing waits for them before the step's gradient
        un-packing, so they go to a dedicated stream that only waits for `dY` and is joined once, at the end of backward().
        X and dY are kept referenced until that join: the caching allocator must not hand their blocks to the main stream
        meanwhile.  Gradients of COMPUTED weights (the type vectors, the KTMachine's limb filters) are read by later closures
        and stay on the current stream."""
        if leaf and self.wstream is not None and self.prof is None:
            cur = torch.cuda.current_stream(self.device)
            ev = torch.cuda.Event()
            ev.record(cur)
            self.wstream.wait_event(ev)
            self._w_keep.append((X, dY))
            with torch.cuda.stream(self.wstream):
                self._wgrad_now(pr, X, dY, dW, M, Ca, Nc, mode, geom)
                self._w_event = torch.cuda.Event()
                self._w_event.record(self.wstream)
            return
        self._wgrad_now(pr, X, dY, dW, M, Ca, Nc, mode, geom)

    def join_wgrad(self):
        if self._w_event is not None:
            torch.cuda.current_stream(self.device).wait_event(self._w_event)
            self._w_event = None
        self._w_keep = []

    def _wgrad_now(self, pr, X, dY, dW, M, Ca, Nc, mode=0, geom=None):
        g = (C.c_int * 8)(*geom) if geom is not None else None
        taps = geom[4] * geom[5] if geom is not None else 1
        self.call('rsg_train_wgrad', _p(X), _p(dY), _p(dW), M, Ca, Nc, Ca, Nc, mode, g, pr, flops=2.0 * M * Ca * Nc * taps,
                  tag=f' M{M} ci{Ca} co{Nc} taps{taps} mode{mode} s{geom[6] if geom is not None else 1}' if self.prof is not None else '')

    def precision(self, precise):
        """Context manager: matrix ops RECORDED inside use 3xTF32 products (forward and backward), whatever the tape's
        default -- the TRP needs it: the GroupNorm behind it amplifies product rounding ~1000x (profiles/r2_notes.md §7)."""
        tape = self

        class _Ctx:
            def __enter__(self):
                self.old = tape.precise
                tape.precise = 1 if precise else 0

            def __exit__(self, *a):
                tape.precise = self.old
        return _Ctx()

    def conv(self, x, wp, wT, k, stride=1, pad=None, bias=None):
        """x [N,H,W,Ci] -> [N,Ho,Wo,Co] (nn.Conv2d).  wp: node of the packed weight [k*k, Ci, Co] (carries the gradient);
        wT: the same weights as [k*k, Co, Ci] (plain tensor), the K-major operand of the forward pass."""
        pr = self.precise
        if pad is None:
            pad = k // 2
        N, H, W, Ci = x.shape
        taps, Cip, Co = wp.shape
        assert taps == k * k and Cip == Ci and tuple(wT.shape) == (taps, Co, Ci), (wp.shape, wT.shape, x.shape, k)
        Ho, Wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
        y = self.new(N, Ho, Wo, Co)
        M = N * Ho * Wo
        plain = k == 1 and stride == 1 and pad == 0
        fg = None if plain else (H, W, Ho, Wo, k, k, stride, pad)
        self._gemm(pr, x.v, wT, y, bias.v if bias is not None else None, M, Co, Ci, Ci, Ci, Co, mode=0 if plain else 1,
                   transB=1, geom=fg)
        out = T(y)

        def bwd():
            dy = out.g
            if dy is None:
                return
            if wp.req:
                self._wgrad(pr, x.v, dy, self.grad_buf(wp), M, Ci, Co, mode=0 if plain else 1, geom=fg, leaf=wp.leaf)
            if bias is not None and bias.req:
                self.call('rsg_train_colsum', _p(dy), M, Co, _p(self.grad_buf(bias)), 1, _p(self.ws(Co)), n=2)
            if x.req:
                dx = self.new(N, H, W, Ci)
                self._gemm(pr, dy, wp.v, dx, None, N * H * W, Ci, Co, Co, Co, Ci, mode=0 if plain else 2, transB=1,
                           geom=None if plain else (Ho, Wo, H, W, k, k, stride, pad))
                self.acc(x, dx)
        self.record(bwd)
        return out

    def conv_transpose(self, x, wp, wT, k, stride, pad, opad=0):
        """x [N,h,w,Ci] -> [N,H,W,Co] (nn.ConvTranspose2d, weight [Ci,Co,k,k]); wp [k*k, Ci, Co] (node), wT [k*k, Co, Ci]."""
        pr = self.precise
        N, h, w, Ci = x.shape
        taps, Cip, Co = wp.shape
        assert taps == k * k and Cip == Ci and tuple(wT.shape) == (taps, Co, Ci)
        H, W = (h - 1) * stride - 2 * pad + k + opad, (w - 1) * stride - 2 * pad + k + opad
        y = self.new(N, H, W, Co)
        M = N * H * W
        tg = (h, w, H, W, k, k, stride, pad)
        self._gemm(pr, x.v, wT, y, None, M, Co, Ci, Ci, Ci, Co, mode=2, transB=1, geom=tg)
        out = T(y)

        def bwd():
            dy = out.g
            if dy is None:
                return
            if wp.req:
                self._wgrad(pr, x.v, dy, self.grad_buf(wp), M, Ci, Co, mode=2, geom=tg, leaf=wp.leaf)
            if x.req:
                dx = self.new(N, h, w, Ci)
                self._gemm(pr, dy, wp.v, dx, None, N * h * w, Ci, Co, Co, Co, Ci, mode=1, transB=1,
                           geom=(H, W, h, w, k, k, stride, pad))
                self.acc(x, dx)
        self.record(bwd)
        return out

    def linear(self, x, w, bias=None):
        """x [..., I], w raw [O, I] (nn.Linear weight or a 1x1 conv's OIHW weight) -> [..., O]."""
        pr = self.precise
        I = x.shape[-1]
        O = w.v.shape[0]
        assert w.v.numel() == O * I, (w.v.shape, x.shape)
        M = x.v.numel() // I
        y = self.new(*x.shape[:-1], O)
        self._gemm(pr, x.v, w.v, y, bias.v if bias is not None else None, M, O, I, I, I, O, transB=1)
        out = T(y)

        def bwd():
            dy = out.g
            if dy is None:
                return
            if w.req:
                self._wgrad(pr, dy, x.v, self.grad_buf(w), M, O, I, leaf=w.leaf)
            if bias is not None and bias.req:
                self.call('rsg_train_colsum', _p(dy), M, O, _p(self.grad_buf(bias)), 1, _p(self.ws(O)), n=2)
            if x.req:
                dx = self.new(*x.shape)
                self._gemm(pr, dy, w.v, dx, None, M, I, O, O, I, I)
                self.acc(x, dx)
        self.record(bwd)
        return out

    def matmul(self, a, b):
        """a [M, K] @ b [K, N] with gradients to both (KTMachine: matrix_relation @ final_layer.weight)."""
        pr = self.precise
        M, K = a.shape
        N = b.v.numel() // K
        y = self.new(M, N)
        self._gemm(pr, a.v, b.v, y, None, M, N, K, K, N, N)
        out = T(y)

        def bwd():
            dy = out.g
            if dy is None:
                return
            if b.req:
                self._wgrad(pr, a.v, dy, self.grad_buf(b), M, K, N, leaf=b.leaf)
            if a.req:
                da = self.new(M, K)
                self._gemm(pr, dy, b.v, da, None, M, K, N, N, N, K, transB=1)
                self.acc(a, da)
        self.record(bwd)
        return out

    # ------------------------------------------------------------------ normalisation
    def batchnorm(self, x, gamma, beta, running_mean, running_var, relu=False, eps=1e-5, momentum=0.1):
        Cn = x.shape[-1]
        M = x.v.numel() // Cn
        y = self.new(*x.shape)
        mean, invstd = self.new(Cn), self.new(Cn)
        self.call('rsg_train_bn_fwd', _p(x.v), M, Cn, _p(gamma.v), _p(beta.v), eps, momentum, _p(running_mean),
                  _p(running_var), 1 if relu else 0, _p(y), _p(mean), _p(invstd), _p(self.ws(3 * Cn + 4)), n=2)
        out = T(y)

        def bwd():
            dy = out.g
            if dy is None:
                return
            dx = self.new(*x.shape) if x.req else None
            self.call('rsg_train_bn_bwd', _p(x.v), _p(y), _p(dy), M, Cn, _p(gamma.v), _p(mean), _p(invstd),
                      1 if relu else 0, _p(dx), _p(gamma.g) if gamma.req else None, _p(beta.g) if beta.req else None,
                      _p(self.ws(3 * Cn + 4)), n=2)
            if dx is not None:
                self.acc(x, dx)
        self.record(bwd)
        return out

    def batchnorm_add_relu(self, x, gamma, beta, running_mean, running_var, res, eps=1e-5, momentum=0.1):
        """relu(bn(x) + res): the tail of a BasicBlock / Bottleneck (pose_rsgnet.py:47-52, 88-93) in the BN's own two launches --
        the residual add + ReLU and, backward, the ReLU mask are folded into the apply kernels."""
        Cn = x.shape[-1]
        if Cn % 4:
            return self.add([self.batchnorm(x, gamma, beta, running_mean, running_var, False, eps, momentum), res], relu=True)
        M = x.v.numel() // Cn
        y = self.new(*x.shape)
        mean, invstd = self.new(Cn), self.new(Cn)
        self.call('rsg_train_bn_fwd_res', _p(x.v), M, Cn, _p(gamma.v), _p(beta.v), eps, momentum, _p(running_mean),
                  _p(running_var), 1, _p(res.v), _p(y), _p(mean), _p(invstd), _p(self.ws(3 * Cn + 4)), n=2)
        out = T(y)

        def bwd():
            dy = out.g
            if dy is None:
                return
            dx = self.new(*x.shape) if x.req else None
            dres = self.new(*x.shape) if res.req else None
            self.call('rsg_train_bn_bwd_res', _p(x.v), _p(y), _p(dy), M, Cn, _p(gamma.v), _p(mean), _p(invstd), 1, _p(dx), _p(dres),
                      _p(gamma.g) if gamma.req else None, _p(beta.g) if beta.req else None, _p(self.ws(3 * Cn + 4)), n=2)
            if dres is not None:
                self.acc(res, dres)
            if dx is not None:
                self.acc(x, dx)
        self.record(bwd)
        return out

    def groupnorm(self, x, gamma, beta, groups=8, eps=1e-5):
        B = x.shape[0]
        Cn = x.shape[-1]
        S = x.v.numel() // (B * Cn)
        y = self.new(*x.shape)
        mean, rstd = self.new(B * groups), self.new(B * groups)
        self.call('rsg_train_gn_fwd', _p(x.v), B, S, Cn, groups, _p(gamma.v), _p(beta.v), eps, _p(y), _p(mean), _p(rstd))
        out = T(y)

        def bwd():
            dy = out.g
            if dy is None:
                return
            dx = self.new(*x.shape) if x.req else None
            self.call('rsg_train_gn_bwd', _p(x.v), _p(dy), B, S, Cn, groups, _p(gamma.v), _p(mean), _p(rstd), _p(dx),
                      _p(gamma.g) if gamma.req else None, _p(beta.g) if beta.req else None)
            if dx is not None:
                self.acc(x, dx)
        self.record(bwd)
        return out

    # ------------------------------------------------------------------ element-wise
    def add(self, xs, relu=False):
        y = self.new(*xs[0].shape)
        ptrs = (C.c_void_p * len(xs))(*[x.v.data_ptr() for x in xs])
        self.call('rsg_train_add', len(xs), ptrs, 1 if relu else 0, y.numel(), _p(y))
        out = T(y)

        def bwd():
            dy = out.g
            if dy is None:
                return
            if relu:
                g = self.new(*y.shape)
                self.call('rsg_train_ew', 0, _p(dy), _p(y), 0.0, y.numel(), _p(g))
            else:
                g = dy
            for x in xs:
                self.acc(x, g, shared=True)
        self.record(bwd)
        return out

    def _unary(self, x, fop, bop, slope=0.0, save_input=False):
        y = self.new(*x.shape)
        self.call('rsg_train_ew', fop, _p(x.v), None, slope, y.numel(), _p(y))
        out = T(y)

        def bwd():
            dy = out.g
            if dy is None or not x.req:
                return
            dx = self.new(*x.shape)
            self.call('rsg_train_ew', bop, _p(dy), _p(x.v if save_input else y), slope, y.numel(), _p(dx))
            self.acc(x, dx)
        self.record(bwd)
        return out

    def sigmoid(self, x):
        return self._unary(x, 1, 2)

    def leaky_relu(self, x, slope):
        return self._unary(x, 3, 4, slope, save_input=True)

    def mul_const(self, x, c):
        """x * c with c a constant tensor of the same shape (KTMachine's real_matrix_limb mask)."""
        y = self.new(*x.shape)
        self.call('rsg_train_ew', 5, _p(x.v), _p(c), 0.0, y.numel(), _p(y))
        out = T(y)

        def bwd():
            if out.g is None or not x.req:
                return
            dx = self.new(*x.shape)
            self.call('rsg_train_ew', 5, _p(out.g), _p(c), 0.0, y.numel(), _p(dx))
            self.acc(x, dx)
        self.record(bwd)
        return out

    def cat(self, xs):
        """Concatenate along the channel (last) dimension."""
        lead = xs[0].shape[:-1]
        M = xs[0].v.numel() // xs[0].shape[-1]
        Ct = sum(x.shape[-1] for x in xs)
        y = self.new(*lead, Ct)
        off = 0
        offs = []
        for x in xs:
            c = x.shape[-1]
            self.call('rsg_train_copy2d', _p(x.v), c, C.c_void_p(y.data_ptr() + 4 * off), Ct, M, c, 0)
            offs.append(off)
            off += c
        out = T(y)

        def bwd():
            dy = out.g
            if dy is None:
                return
            for x, o in zip(xs, offs):
                if not x.req:
                    continue
                c = x.shape[-1]
                dx = self.new(*x.shape)
                self.call('rsg_train_copy2d', C.c_void_p(dy.data_ptr() + 4 * o), Ct, _p(dx), c, M, c, 0)
                self.acc(x, dx)
        self.record(bwd)
        return out

    def repeat_batch(self, x, n):
        """x [1, ...] -> [n, ...] (pose_rsgnet.py:980 loc_features.repeat)."""
        cols = x.v.numel()
        y = self.new(n, *x.shape[1:])
        self.call('rsg_train_copy2d', _p(x.v), 0, _p(y), cols, n, cols, 0)
        out = T(y)

        def bwd():
            if out.g is None or not x.req:
                return
            dx = self.new(*x.shape)
            self.call('rsg_train_colsum', _p(out.g), n, cols, _p(dx), 0, _p(self.ws(cols)), n=3)
            self.acc(x, dx)
        self.record(bwd)
        return out

    def _resample(self, x, kf, kb, f):
        N, h, w, Cn = x.shape
        y = self.new(N, h * f, w * f, Cn)
        self.call('rsg_train_resample', kf, _p(x.v), N, h, w, Cn, f, _p(y))
        out = T(y)

        def bwd():
            if out.g is None or not x.req:
                return
            dx = self.new(*x.shape)
            self.call('rsg_train_resample', kb, _p(out.g), N, h, w, Cn, f, _p(dx), n=1 if kb == 1 else 2)
            self.acc(x, dx)
        self.record(bwd)
        return out

    def upsample_nearest(self, x, f):
        return self._resample(x, 0, 1, f)

    def bilinear2x(self, x):
        return self._resample(x, 2, 3, 2)

    def maxpool2(self, x):
        N, H, W, Cn = x.shape
        y = self.new(N, H // 2, W // 2, Cn)
        idx = torch.empty(y.shape, dtype=torch.uint8, device=self.device)
        self.call('rsg_train_maxpool', 0, _p(x.v), _p(idx), N, H, W, Cn, _p(y))
        out = T(y)

        def bwd():
            if out.g is None or not x.req:
                return
            dx = self.new(*x.shape)
            self.call('rsg_train_maxpool', 1, _p(out.g), _p(idx), N, H, W, Cn, _p(dx))
            self.acc(x, dx)
        self.record(bwd)
        return out

    # ------------------------------------------------------------------ layout
    def view(self, x, *shape):
        """Reshape without copying; the gradient of the view flows back to x as a view of the same storage."""
        out = T(x.v.view(*shape), req=x.req)

        def bwd():
            if out.g is not None and x.req:
                self.acc(x, out.g.view(*x.shape), shared=not out.own)
        self.record(bwd)
        return out

    def from_nchw(self, x, cpad=None):
        """torch NCHW tensor -> NHWC node (channels zero-padded to cpad); no gradient flows to the input."""
        N, Cn, H, W = x.shape
        cp = cpad or Cn
        y = self.new(N, H, W, cp)
        self.call('rsg_train_permute3', _p(x), _p(y), N, H * W, cp, Cn * H * W, 1, H * W, H * W, Cn, 0)
        return T(y, req=False)

    def to_nchw(self, x):
        """NHWC node -> NCHW node (the tensors the reference's caller sees)."""
        N, H, W, Cn = x.shape
        y = self.new(N, Cn, H, W)
        self.call('rsg_train_permute3', _p(x.v), _p(y), N, Cn, H * W, H * W * Cn, 1, Cn, Cn, H * W, 0)
        out = T(y)

        def bwd():
            if out.g is None or not x.req:
                return
            dx = self.new(*x.shape)
            self.call('rsg_train_permute3', _p(out.g), _p(dx), N, H * W, Cn, Cn * H * W, 1, H * W, H * W, Cn, 0)
            self.acc(x, dx)
        self.record(bwd)
        return out

    # ------------------------------------------------------------------ TRP (association.py:280-301)
    def trp_attention(self, x, g):
        """x [B,S,C] (theta = phi = x), g [B,S,Cg] -> (y [B,S,Cg] = sigmoid(x x^T) g, P [B,S,S] as a node).
        The gradient of a loss on P (the relation MSE) is added by the caller through ``P.g`` or ``rel_hook``."""
        pr = self.precise
        # dP = dy g^T and dx = (dA + dA^T) x are ordinary gradient GEMMs: they follow the tape's default precision.  dg = P dy
        # stays with the forward's 3xTF32: the GroupNorm backward makes sum_j dy_j ~ 0 and every row of P is nearly constant,
        # so dg is what is left of a cancellation (the g weights' gradient moved by 3e-3 with TF32 products)
        pb = self.default_precise
        B, S, Cx = x.shape
        Cg = g.shape[-1]
        P = self.new(B, S, S)
        self._gemm(pr, x.v, x.v, P, None, S, S, Cx, Cx, Cx, S, transB=1, batch=B, sA=S * Cx, sB=S * Cx, sC=S * S)
        self.call('rsg_train_ew', 1, _p(P), None, 0.0, P.numel(), _p(P))
        y = self.new(B, S, Cg)
        self._gemm(pr, P, g.v, y, None, S, Cg, S, S, Cg, Cg, batch=B, sA=S * S, sB=S * Cg, sC=S * Cg)
        out, Pn = T(y), T(P)
        Pn.g = None
        hook = {}

        def bwd():
            dy = out.g
            dP = None
            if dy is not None:
                if g.req:
                    dg = self.new(B, S, Cg)          # P is symmetric: P^T dy = P dy
                    self._gemm(pr, P, dy, dg, None, S, Cg, S, S, Cg, Cg, batch=B, sA=S * S, sB=S * Cg, sC=S * Cg)
                    self.acc(g, dg)
                dP = self.new(B, S, S)
                self._gemm(pb, dy, g.v, dP, None, S, S, Cg, Cg, Cg, S, transB=1, batch=B, sA=S * Cg, sB=S * Cg, sC=S * S)
            if Pn.g is not None:                      # a caller differentiated through the returned scores directly
                if dP is None:
                    dP = Pn.g
                else:
                    self.call('rsg_train_ew', 6, _p(Pn.g), None, 0.0, dP.numel(), _p(dP))
            rel = hook.get('rel')                     # (T_full | None, v | None, coef [B])
            if dP is None and rel is None:
                return
            if not x.req:
                return
            dA = dP if dP is not None else self.new(B, S, S)
            self.call('rsg_train_trp_dscore', _p(P), _p(dP), _p(rel[0]) if rel else None, _p(rel[1]) if rel else None,
                      _p(rel[2]) if rel else None, B, S, _p(dA))
            dx = self.new(B, S, Cx)                   # A = x x^T: dx = dA x + dA^T x
            self._gemm(pb, dA, x.v, dx, None, S, Cx, S, S, Cx, Cx, batch=B, sA=S * S, sB=S * Cx, sC=S * Cx)
            self._gemm(pb, dA, x.v, dx, None, S, Cx, S, S, Cx, Cx, transA=1, beta=1, batch=B, sA=S * S, sB=S * Cx, sC=S * Cx)
            self.acc(x, dx)
        self.record(bwd)
        return out, Pn, hook
