"""Host side of the data-parallel training step on CPU (no compute call): the flat parameter / gradient layout of
rsgnet_b200.train.step.ParamStore, its pack / unpack tables, and the world_size-2 gradient exchange over gloo exactly as
TrainStep.all_reduce issues it (one all-reduce of the flat buffer, mean taken by Adam's 1/world scale)."""
import ctypes
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn

from rsgnet_b200 import _lib, presets
from rsgnet_b200.models import _params, pose_rsgnet


def _store():
    from rsgnet_b200.train.step import ParamStore
    cfg = presets.preset('tiny_cp')
    net = pose_rsgnet.get_pose_net(cfg, True)
    net.load_state_dict(_params.synth_state_dict(net, seed=1))
    return net, ParamStore(net, torch.device('cpu'))


def test_flat_layout_aliases_the_module_parameters():
    net, st = _store()
    seen = 0
    for p in net.parameters():
        off, n = st.offsets[id(p)]
        assert off % 4 == 0 and n == p.numel()                     # 16-byte aligned views
        assert p.data.data_ptr() == st.flat_p.data_ptr() + 4 * off  # the module's tensors ARE the flat buffer
        leaf = st.leaf[id(p)]
        assert leaf.g.data_ptr() == st.flat_g.data_ptr() + 4 * off and leaf.req == p.requires_grad
        seen += n
    assert seen <= st.total < seen + 4 * len(st.params)
    # load_state_dict copies in place: aliasing survives, and the flat buffer follows
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    sd['final_layer.weight'] = sd['final_layer.weight'] + 1.0
    net.load_state_dict(sd)
    off, n = st.offsets[id(net.final_layer.weight)]
    assert torch.equal(st.flat_p[off:off + n].view_as(net.final_layer.weight), sd['final_layer.weight'])
    assert not net.loc_features.requires_grad and not st.leaf[id(net.loc_features)].req


def test_pack_tables_cover_every_conv():
    """Every Conv2d / ConvTranspose2d weight is packed to [tap][Cin][Cout] (input-gradient operand, gradient layout) and to
    [tap][Cout][Cin] (forward operand; a 1x1 Conv2d's parameter already is that).  The tables are emulated on the host."""
    net, st = _store()
    convs = [m for m in net.modules() if isinstance(m, (nn.ConvTranspose2d, nn.Conv2d))]
    assert st.n_packed == len(convs) and len(st.packed) == len(convs) and st.n_unpack == len(convs)
    by_src = {m.weight.data.data_ptr(): m for m in convs}
    raw = bytes(st.pack_table.numpy().tobytes())
    ent = (_lib.PermEntry * st.n_pack).from_buffer_copy(raw)
    seen = set()
    for e in ent:
        m = by_src[e.src]
        node, wT = st.packed[id(m)]
        t = m.kernel_size[0] * m.kernel_size[1]
        tr = isinstance(m, nn.ConvTranspose2d)
        w = m.weight.detach()
        ci, co = (w.shape[0], w.shape[1]) if tr else (w.shape[1], w.shape[0])
        w_tcc = (w.permute(2, 3, 0, 1) if tr else w.permute(2, 3, 1, 0)).reshape(t, ci, co)       # [tap][ci][co]
        flat = w.reshape(-1)
        i0, i1, i2 = torch.meshgrid(torch.arange(e.D0), torch.arange(e.V1), torch.arange(e.V2), indexing='ij')
        got = flat[(i0 * e.s0 + i1 * e.s1 + i2 * e.s2).reshape(-1)].reshape(e.D0, e.V1, e.V2)
        assert e.accumulate == 0 and e.D0 == t
        if e.dst == node.v.data_ptr():
            assert (e.D1, e.D2, e.V1, e.V2) == (node.v.shape[1], co, ci, co) and e.D1 % 4 == 0 and torch.equal(got, w_tcc)
            seen.add((id(m), 'w'))
        else:
            assert e.dst == wT.data_ptr() and (e.D1, e.D2, e.V1, e.V2) == (co, wT.shape[2], co, ci)
            assert torch.equal(got, w_tcc.permute(0, 2, 1))
            seen.add((id(m), 'wT'))
    for m in convs:
        node, wT = st.packed[id(m)]
        assert (id(m), 'w') in seen
        if (id(m), 'wT') not in seen:          # 1x1 Conv2d with Cin % 4 == 0: the parameter itself
            assert wT.data_ptr() == m.weight.data.data_ptr() and m.kernel_size == (1, 1)
            assert torch.equal(wT, m.weight.detach().view(1, m.weight.shape[0], m.weight.shape[1]))
    assert ctypes.sizeof(_lib.PermEntry) == 72


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from rsgnet_b200.train.step import TrainStep
    net, st = _store()
    ts = TrainStep.__new__(TrainStep)              # the exchange logic only: no CUDA device in this test
    ts.store, ts.pg, ts.world = st, None, dist.get_world_size()
    st.flat_g.fill_(float(rank + 1))
    off, n = st.offsets[id(net.final_layer.weight)]
    st.flat_g[off:off + n] = torch.arange(n, dtype=torch.float32) * (rank + 1)
    ts.all_reduce()
    scale = 1.0 / ts.world                         # what TrainStep.adam passes to rsg_train_adam as grad_scale
    mean = st.flat_g * scale
    ok = bool(torch.allclose(mean[off:off + n], torch.arange(n, dtype=torch.float32) * 1.5)) and float(mean[0]) == 1.5
    q.put((rank, ok))
    dist.barrier()
    dist.destroy_process_group()


def test_gradient_exchange_world2_gloo():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=300) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert got == [(0, True), (1, True)]
