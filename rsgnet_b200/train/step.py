"""Training step of the drop-in models: forward (batch-statistics BN) + losses + backward + all-reduce + Adam.

Replaces /root/reference/lib/core/function.py:240-363 (rsgnet_train) / :33-100 (train) for one iteration:
    relation target (:256-269) -> model(input, relation_target) -> JointsMSELoss x2 + 0.01 BCE + 0.001 relation (:283-313)
    -> loss.backward() -> optimizer.step() (Adam, lib/utils/utils.py:70-74)
All arithmetic runs in librsg_b200.so (rsgnet_b200/csrc/train_*.cu) through the tape of ``rsgnet_b200.train.tape``.

Memory layout: every parameter of the module becomes a view into ONE flat fp32 buffer (``flat_p``); gradients, and Adam's
two moments, are flat buffers of the same layout, so the optimiser is one launch and the data-parallel gradient exchange
is one NCCL all-reduce (117 MB for W32; SURVEY.md §8e 'Training').  Convolution weights with a kernel larger than 1x1 are
re-packed once per step to ``[tap][Cin][Cout]`` (one batched launch); their gradients are produced in that layout and
un-packed into the flat gradient buffer (one batched launch).
"""
import ctypes as C

import torch
import torch.nn as nn

from .. import _lib
from ..config import KIND_RSGNET
from .net import Net
from .tape import T, Tape, _p, side_streams


def _ceil4(n):
    return (n + 3) // 4 * 4


class ParamStore:
    """Flat parameter / gradient storage of one module on one device + the packed conv weights."""

    def __init__(self, module, device):
        self.device = device
        self.lib = _lib.lib()
        params = []
        seen = set()
        for p in module.parameters():
            if id(p) not in seen:
                seen.add(id(p))
                params.append(p)
        self.params = params
        total = sum(_ceil4(p.numel()) for p in params)
        self.flat_p = torch.zeros(total, dtype=torch.float32, device=device)
        self.flat_g = torch.zeros(total, dtype=torch.float32, device=device)
        self.leaf = {}
        self.offsets = {}
        off = 0
        for p in params:
            n = p.numel()
            view = self.flat_p[off:off + n].view(p.shape)
            view.copy_(p.data)
            p.data = view
            self.leaf[id(p)] = T(view, req=p.requires_grad, g=self.flat_g[off:off + n].view(p.shape), leaf=True)
            self.offsets[id(p)] = (off, n)
            off += _ceil4(n)
        self.total = total
        # Packed conv weights and their gradients.  Every Conv2d / ConvTranspose2d weight is kept in the two K-major forms
        # the implicit-GEMM kernels read: wT [tap][Cout][CinPad] (reduction over Cin contiguous: the forward pass) and
        # w [tap][CinPad][Cout] (reduction over Cout contiguous: the input-gradient pass; also the layout the weight-gradient
        # kernel writes).  For a 1x1 Conv2d, wT is the parameter itself.
        convs = [m for m in module.modules() if isinstance(m, (nn.ConvTranspose2d, nn.Conv2d))]
        sizes = []
        for m in convs:
            w = m.weight
            tr = isinstance(m, nn.ConvTranspose2d)
            ci, co = (w.shape[0], w.shape[1]) if tr else (w.shape[1], w.shape[0])
            sizes.append((m, tr, ci, _ceil4(ci), co, w.shape[2] * w.shape[3]))
        ptotal = sum(t * cip * co for (_, _, _, cip, co, t) in sizes)
        self.packed_w = torch.zeros(max(ptotal, 4), dtype=torch.float32, device=device)
        self.packed_wT = torch.zeros(max(ptotal, 4), dtype=torch.float32, device=device)
        self.packed_g = torch.zeros(max(ptotal, 4), dtype=torch.float32, device=device)
        self.packed = {}
        pack, unpack = [], []
        off = 0
        for (m, tr, ci, cip, co, t) in sizes:
            n = t * cip * co
            wv = self.packed_w[off:off + n].view(t, cip, co)
            gv = self.packed_g[off:off + n].view(t, cip, co)
            src = m.weight.data.data_ptr()
            poff, _ = self.offsets[id(m.weight)]
            gdst = self.flat_g.data_ptr() + 4 * poff
            if not tr and t == 1 and cip == ci:
                wT = m.weight.data.view(1, co, ci)              # [Cout][Cin] is already the forward operand
            else:
                wT = self.packed_wT[off:off + n].view(t, co, cip)
            node = T(wv, req=m.weight.requires_grad, g=gv, leaf=True)
            self.packed[id(m)] = (node, wT)
            if not tr:      # OIHW: w[tap][ci][co] = W[co*Ci*T + ci*T + tap];  wT[tap][co][ci] likewise
                pack.append(_lib.PermEntry(src, wv.data_ptr(), t, cip, co, ci, co, 1, t, ci * t, 0))
                if wT.data_ptr() != src:
                    pack.append(_lib.PermEntry(src, wT.data_ptr(), t, co, cip, co, ci, 1, ci * t, t, 0))
                # grad OIHW [co][ci][tap] += packed_g[tap*CiP*Co + ci*Co + co]
                unpack.append(_lib.PermEntry(gv.data_ptr(), gdst, co, ci, t, ci, t, 1, co, cip * co, 1))
            else:           # ConvTranspose2d [Ci][Co][kh][kw]: w[tap][ci][co] = W[ci*Co*T + co*T + tap]
                pack.append(_lib.PermEntry(src, wv.data_ptr(), t, cip, co, ci, co, 1, co * t, t, 0))
                pack.append(_lib.PermEntry(src, wT.data_ptr(), t, co, cip, co, ci, 1, t, co * t, 0))
                unpack.append(_lib.PermEntry(gv.data_ptr(), gdst, ci, co, t, co, t, co, 1, cip * co, 1))
            off += n
        self.n_packed = len(sizes)

        def table(entries):
            arr = (_lib.PermEntry * max(len(entries), 1))(*entries)
            return torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).clone().to(device), len(entries)
        self.pack_table, self.n_pack = table(pack)
        self.unpack_table, self.n_unpack = table(unpack)
        self.bn_counters = [m.num_batches_tracked for m in module.modules()
                            if isinstance(m, (nn.BatchNorm2d, nn.BatchNorm1d)) and m.num_batches_tracked is not None]

    def raw(self, p):
        return self.leaf[id(p)]

    def pack_all(self, tape):
        if self.n_pack:
            tape.call('rsg_train_permute3_batch', _p(self.pack_table), self.n_pack, 16)

    def unpack_grads(self, tape):
        if self.n_unpack:
            tape.call('rsg_train_permute3_batch', _p(self.unpack_table), self.n_unpack, 16)

    def zero_grads(self, tape):
        tape.call('rsg_train_zero', _p(self.flat_g), C.c_size_t(4 * self.flat_g.numel()))
        tape.call('rsg_train_zero', _p(self.packed_g), C.c_size_t(4 * self.packed_g.numel()))


class _StoreView:
    """What Net needs from the store."""

    def __init__(self, store):
        self.s = store

    def raw(self, p):
        return self.s.leaf[id(p)]

    def packed(self, conv):
        return self.s.packed[id(conv)]


def store_for(module, device):
    st = module.__dict__.get('_rsg_train_store')
    if st is None or st.device != device or any(p.data.data_ptr() != st.leaf[id(p)].v.data_ptr() for p in st.params[:4]):
        st = ParamStore(module, device)
        module.__dict__['_rsg_train_store'] = st
    return st


class Losses:
    """Loss bookkeeping on the device: acc = [multi, target, skeleton, relation_0 .. relation_{B-1}] in doubles."""

    def __init__(self, tape, B):
        self.tape, self.B = tape, B
        self.acc = torch.empty(3 + B, dtype=torch.float64, device=tape.device)
        tape.call('rsg_train_zero', _p(self.acc), C.c_size_t(8 * (3 + B)))

    def slot(self, i):
        return C.c_void_p(self.acc.data_ptr() + 8 * i)

    def read(self):
        a = self.acc.cpu().numpy()            # the step's one D2H copy (the reference's loss.item(), function.py:323)
        rel = 0.001 * float(a[3:].mean()) if self.B else 0.0
        return dict(multi_loss=float(a[0]), target_loss=float(a[1]), skeleton_loss=float(a[2]), relation_loss=rel,
                    loss=float(a[0] + a[1] + a[2]) + rel)


class TrainStep:
    """One optimisation step per call.  ``world_size > 1``: one process per GPU, gradients summed with one NCCL all-reduce
    of the flat buffer and scaled by 1/world_size inside the Adam kernel (DistributedDataParallel's mean)."""

    def __init__(self, module, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, process_group=None, precise=False, device=None):
        _lib.require_cuda()
        self.module = module
        self.device = device or next(module.parameters()).device
        if self.device.type != 'cuda':
            raise _lib.RsgError('TrainStep: move the model to a CUDA device first (no CPU path)')
        self.store = store_for(module, self.device)
        self.lr, self.betas, self.eps = lr, betas, eps
        self.m = torch.zeros_like(self.store.flat_p)
        self.v = torch.zeros_like(self.store.flat_p)
        self.t = 0
        self.pg = process_group
        self.precise = precise
        self.world = 1
        if process_group is not None or (torch.distributed.is_available() and torch.distributed.is_initialized()):
            self.world = torch.distributed.get_world_size(process_group)
        self.last_launches = 0
        self.profile = False          # True: CUDA events around every library call of the next step (see tape.prof)
        self.concurrent = True        # HRNet branches on side streams (Tape.parallel); False = one stream
        self.side = side_streams(self.device, 4)    # 3 for parallel chains, 1 for the weight gradients
        # frozen parameters (loc_features, kt_machine.real_matrix_limb) keep a zero gradient and are skipped by masking lr:
        # Adam with g = 0, m = v = 0 leaves them unchanged (update = 0 / (0 + eps) = 0)

    # ---- pieces ------------------------------------------------------------------------------------------------------
    def forward_backward(self, x, target, target_weight, all_target=None, all_target_weight=None, target_limbs=None,
                         relation_target=None):
        """x [B,3,H,W]; target / all_target [B,K,2h,2w]; *_weight [B,K,1]; target_limbs [B,L,2h,2w]; all CUDA fp32.
        ``relation_target``: None = built on the device from `target` as the reference does on the host
        (function.py:256-269), or the reference's explicit [B,S,S] tensor.  Returns (losses dict after sync, outputs)."""
        mod, dev = self.module, self.device
        with torch.cuda.device(dev):
            tape = Tape(dev, self.precise, side=self.side[:3] if self.concurrent else None,
                        wstream=self.side[3] if self.concurrent else None)
            if self.profile:
                tape.prof = []
            self.tape = tape
            st = self.store
            st.zero_grads(tape)
            st.pack_all(tape)
            net = Net(mod, _StoreView(st), tape)
            x = x.to(dev, torch.float32).contiguous()
            B = int(x.shape[0])
            losses = Losses(tape, B if mod.spec.kind == KIND_RSGNET else 0)
            cont = lambda t: t.to(dev, torch.float32).contiguous()
            keep = []            # device copies of the targets must outlive the asynchronous launches that read them
            if mod.spec.kind != KIND_RSGNET:
                out = net.hrnet(x)
                K, HW = out.shape[1], out.shape[2] * out.shape[3]
                out.g = tape.new(*out.shape)
                target, target_weight = cont(target), cont(target_weight)
                keep += [target, target_weight]
                tape.call('rsg_train_mse_joints', _p(out.v), _p(target), _p(target_weight), B, K, HW, 1.0,
                          losses.slot(1), _p(out.g))
                outputs = (out.v,)
            else:
                multi, kpt, limbs, P, hook = net.rsgnet(x)
                K, HW = kpt.shape[1], kpt.shape[2] * kpt.shape[3]
                target, target_weight = cont(target), cont(target_weight)
                all_target, all_target_weight, target_limbs = cont(all_target), cont(all_target_weight), cont(target_limbs)
                keep += [target, target_weight, all_target, all_target_weight, target_limbs]
                kpt.g = tape.new(*kpt.shape)
                tape.call('rsg_train_mse_joints', _p(kpt.v), _p(target), _p(target_weight), B, K, HW, 1.0,
                          losses.slot(1), _p(kpt.g))
                multi.g = tape.new(*multi.shape)
                tape.call('rsg_train_mse_joints', _p(multi.v), _p(all_target), _p(all_target_weight), B, K, HW,
                          1.0, losses.slot(0), _p(multi.g))
                limbs.g = tape.new(*limbs.shape)
                tape.call('rsg_train_bce', _p(limbs.v), _p(target_limbs), limbs.v.numel(), 0.01, 1.0, losses.slot(2),
                          _p(limbs.g))
                S = P.shape[1]
                full = vec = None
                if relation_target is None:
                    vec = tape.new(B, S)
                    Hh, Wh = target.shape[2], target.shape[3]
                    assert (Hh // 2) * (Wh // 2) == S, 'relation target: heat-map size must be 2x the feature map'
                    tape.call('rsg_train_person_mask', _p(target), B, K, Hh, Wh, _p(vec))
                else:
                    full = cont(relation_target)
                tape.call('rsg_train_relation_mse', _p(P.v), _p(full), _p(vec), B, S, losses.slot(3))
                hook['rel'] = (full, vec, self._rel_coef(B, S))
                outputs = (multi.v, kpt.v, limbs.v, P.v)
            for o in (multi, kpt, limbs) if mod.spec.kind == KIND_RSGNET else (out,):
                o.own = True
            tape.backward()
            st.unpack_grads(tape)
            self.last_launches = tape.launches
            self.last_flops = tape.flops
            losses.keep = keep
            if st.bn_counters:
                torch._foreach_add_(st.bn_counters, 1)
        return losses, outputs

    def _rel_coef(self, B, S):
        key = (B, S)
        c = getattr(self, '_coef_cache', {})
        if key not in c:
            c[key] = torch.full((B,), 0.001 / B * 2.0 / (float(S) * float(S)), dtype=torch.float32, device=self.device)
            self._coef_cache = c
        return c[key]

    def all_reduce(self):
        if self.world > 1:
            torch.distributed.all_reduce(self.store.flat_g, group=self.pg)

    def adam(self):
        self.t += 1
        st = self.store
        _lib.check(self.lib_adam()(_lib.stream_ptr(self.device), _p(st.flat_p), _p(st.flat_g), _p(self.m), _p(self.v),
                                   st.total, self.lr, self.betas[0], self.betas[1], self.eps, self.t, 1.0 / self.world))
        self.last_launches += 1

    def lib_adam(self):
        return _lib.lib().rsg_train_adam

    def __call__(self, x, target, target_weight, all_target=None, all_target_weight=None, target_limbs=None,
                 relation_target=None, sync=True):
        losses, outputs = self.forward_backward(x, target, target_weight, all_target, all_target_weight, target_limbs,
                                                relation_target)
        with torch.cuda.device(self.device):
            self.all_reduce()
            self.adam()
        return (losses.read() if sync else losses), outputs

    # ---- CUDA-graph replay -------------------------------------------------------------------------------------------------
    # A step is ~4 000 small launches issued from Python: eagerly the host is the bottleneck.  The sequence is static for a
    # fixed batch shape, so forward + losses + backward are captured ONCE into a CUDA graph over pointer-stable input
    # buffers and replayed; the gradient all-reduce and the Adam launch follow the replay on the same stream.
    def build_graph(self, x, target, target_weight, all_target=None, all_target_weight=None, target_limbs=None):
        dev = self.device
        args = (x, target, target_weight, all_target, all_target_weight, target_limbs)
        with torch.cuda.device(dev):
            self._static = [a.to(dev, torch.float32).contiguous().clone() if a is not None else None for a in args]
            bufs = [b for b in self.module.buffers()]
            saved = [b.detach().clone() for b in bufs]
            side = torch.cuda.Stream(dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):                   # one eager pass: lazy initialisation, allocator warm-up
                self.forward_backward(*self._static)
            torch.cuda.current_stream(dev).wait_stream(side)
            torch.cuda.synchronize(dev)
            for b, sv in zip(bufs, saved):                  # the warm-up pass must not count as a training step
                b.copy_(sv)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                losses, outputs = self.forward_backward(*self._static)
            self.tape = None
            self._graph = (graph, losses, outputs)
        return self

    def step_graph(self, x, target, target_weight, all_target=None, all_target_weight=None, target_limbs=None, sync=True):
        """One optimisation step by graph replay (build_graph first; same batch shape every call)."""
        graph, losses, outputs = self._graph
        with torch.cuda.device(self.device):
            for s, a in zip(self._static, (x, target, target_weight, all_target, all_target_weight, target_limbs)):
                if s is not None:
                    s.copy_(a, non_blocking=True)
            graph.replay()
            self.all_reduce()
            self.adam()
        return (losses.read() if sync else losses), outputs


# ----------------------------------------------------------------------------------------------------------------------
# Drop-in route: model.train(); outputs = model(input, relation_target); loss.backward() as lib/core/function.py:271-319
# does it.  One autograd node spans the whole network: its forward runs the tape forward, its backward the tape backward.
class _NetFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, x, relation_target, *params):
        dev = params[0].device
        st = store_for(module, dev)
        with torch.cuda.device(dev):
            tape = Tape(dev, False)
            st.zero_grads(tape)
            st.pack_all(tape)
            net = Net(module, _StoreView(st), tape)
            x = x.detach().to(dev, torch.float32).contiguous()
            ctx.tape, ctx.store, ctx.nparams = tape, st, len(params)
            ctx.params = params
            if module.spec.kind != KIND_RSGNET:
                out = net.hrnet(x)
                ctx.nodes = (out,)
                ctx.rel = None
                res = (out.v,)
            else:
                multi, kpt, limbs, P, hook = net.rsgnet(x)
                B, S = P.shape[0], P.shape[1]
                ctx.nodes = (multi, kpt, limbs)
                if relation_target is not None:
                    full = relation_target.detach().to(dev, torch.float32).contiguous()
                    acc = torch.empty(B, dtype=torch.float64, device=dev)
                    tape.call('rsg_train_zero', _p(acc), C.c_size_t(8 * B))
                    tape.call('rsg_train_relation_mse', _p(P.v), _p(full), None, B, S, _p(acc))
                    rel = tape.new(B)
                    tape.call('rsg_train_d2f', _p(acc), 1.0, B, 0, _p(rel))
                    ctx.rel = (P, hook, full, S)
                    res = (multi.v, kpt.v, limbs.v, rel)
                else:
                    ctx.rel = (P, hook, None, S)
                    res = (multi.v, kpt.v, limbs.v, P.v)
            if st.bn_counters:
                torch._foreach_add_(st.bn_counters, 1)
        return res

    @staticmethod
    def backward(ctx, *grads):
        tape, st = ctx.tape, ctx.store
        with torch.cuda.device(st.device):
            for node, g in zip(ctx.nodes, grads):
                if g is not None:
                    node.g = g.to(torch.float32).contiguous()
                    node.own = False
            if ctx.rel is not None and len(grads) > 3 and grads[3] is not None:
                P, hook, full, S = ctx.rel
                g = grads[3].to(torch.float32).contiguous()
                if full is not None:        # d out[b] / dP = 2 (P - T) / S^2
                    coef = tape.new(g.numel())
                    tape.call('rsg_train_ew', 7, _p(g), None, 2.0 / (float(S) * float(S)), g.numel(), _p(coef))
                    hook['rel'] = (full, None, coef)
                else:
                    P.g = g
                    P.own = False
            tape.backward()
            st.unpack_grads(tape)
        snap = st.flat_g.clone()            # ONE copy: the next forward zeroes flat_g, p.grad must survive it
        out = []
        for p in ctx.params:
            off, n = st.offsets[id(p)]
            out.append(snap[off:off + n].view(p.shape) if p.requires_grad else None)
        return (None, None, None, *out)


def module_forward_train(module, x, relation_target=None):
    """forward() of the drop-in modules in training mode (the reference's `model(input, relation_target)`,
    lib/core/function.py:271): outputs carry an autograd node whose backward is the tape backward."""
    _lib.require_cuda()
    dev = next(module.parameters()).device
    if dev.type != 'cuda':
        raise _lib.RsgError('move the model to a CUDA device first (model.cuda()); no CPU path')
    store_for(module, dev)                      # flatten before autograd sees the parameters
    params = tuple(p for p in module.parameters())
    res = _NetFunction.apply(module, x, relation_target, *params)
    return res[0] if module.spec.kind != KIND_RSGNET else tuple(res)


class FusedAdam(torch.optim.Optimizer):
    """torch.optim.Adam replacement for the reference's get_optimizer (lib/utils/utils.py:70-74): one launch of
    rsg_train_adam per parameter tensor (p.grad as produced by backward()).  TrainStep is the faster route (one launch
    for the whole model); this class exists so that the reference's training loop runs on the library end to end."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))

    @torch.no_grad()
    def step(self, closure=None):
        lib = _lib.lib()
        for group in self.param_groups:
            for p in group['params']:
                if p.grad is None:
                    continue
                stt = self.state[p]
                if not stt:
                    stt['step'] = 0
                    stt['m'] = torch.zeros_like(p.data)
                    stt['v'] = torch.zeros_like(p.data)
                stt['step'] += 1
                g = p.grad.contiguous()
                with torch.cuda.device(p.device):
                    _lib.check(lib.rsg_train_adam(_lib.stream_ptr(p.device), _p(p.data), _p(g), _p(stt['m']), _p(stt['v']),
                                                  p.numel(), group['lr'], group['betas'][0], group['betas'][1],
                                                  group['eps'], stt['step'], 1.0))
