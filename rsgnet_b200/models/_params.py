"""Parameter containers whose ``state_dict`` keys and shapes equal the reference's.

These modules own the fp32 ``nn.Parameter``s / BN buffers so that reference checkpoints load with
``load_state_dict`` (names per SURVEY.md App. A.4, i.e. /root/reference/lib/models/
pose_rsgnet.py:611-775 and association.py:228-268).  They are never *called*: inference runs
through the CUDA library (``rsgnet_b200._engine``), which reads these tensors, folds and packs them.
"""
import os
import pickle

import numpy as np
import torch
import torch.nn as nn


def _holder(**children):
    m = nn.Module()
    for k, v in children.items():
        setattr(m, k, v)
    return m


def conv_bn(cin, cout, k, stride=1, relu=False, transposed=False):
    """Sequential(conv(no bias), BN[, ReLU]) -> keys '0.weight', '1.{weight,bias,running_*}'."""
    if transposed:
        pad, opad = {4: (1, 0), 3: (1, 1), 2: (0, 0)}[k]
        conv = nn.ConvTranspose2d(cin, cout, k, 2, pad, opad, bias=False)
    else:
        conv = nn.Conv2d(cin, cout, k, stride, k // 2, bias=False)
    layers = [conv, nn.BatchNorm2d(cout)]
    if relu:
        layers.append(nn.ReLU(inplace=True))
    return nn.Sequential(*layers)


def bottleneck(cin, planes, with_proj):
    m = _holder(conv1=nn.Conv2d(cin, planes, 1, bias=False), bn1=nn.BatchNorm2d(planes),
                conv2=nn.Conv2d(planes, planes, 3, 1, 1, bias=False), bn2=nn.BatchNorm2d(planes),
                conv3=nn.Conv2d(planes, planes * 4, 1, bias=False),
                bn3=nn.BatchNorm2d(planes * 4))
    if with_proj:
        m.downsample = conv_bn(cin, planes * 4, 1)
    return m


def basic_block(c):
    return _holder(conv1=nn.Conv2d(c, c, 3, 1, 1, bias=False), bn1=nn.BatchNorm2d(c),
                   conv2=nn.Conv2d(c, c, 3, 1, 1, bias=False), bn2=nn.BatchNorm2d(c))


def hr_module(channels, num_blocks, multi_scale_output):
    nb = len(channels)
    m = nn.Module()
    m.branches = nn.ModuleList(
        nn.Sequential(*[basic_block(channels[b]) for _ in range(num_blocks[b])])
        for b in range(nb))
    if nb > 1:
        rows = []
        for i in range(nb if multi_scale_output else 1):
            row = []
            for j in range(nb):
                if j == i:
                    row.append(None)
                elif j > i:
                    row.append(conv_bn(channels[j], channels[i], 1))
                else:
                    hops = i - j
                    row.append(nn.Sequential(*[
                        conv_bn(channels[j], channels[i] if h == hops - 1 else channels[j],
                                3, stride=2, relu=(h != hops - 1)) for h in range(hops)]))
            rows.append(nn.ModuleList(row))
        m.fuse_layers = nn.ModuleList(rows)
    return m


def transition(prev_channels, cur_channels):
    n_pre = len(prev_channels)
    layers = []
    for i, c in enumerate(cur_channels):
        if i < n_pre:
            layers.append(conv_bn(prev_channels[i], c, 3, relu=True)
                          if c != prev_channels[i] else None)
        else:
            hops = i + 1 - n_pre
            cin = prev_channels[-1]
            layers.append(nn.Sequential(*[
                conv_bn(cin, c if h == hops - 1 else cin, 3, stride=2, relu=True)
                for h in range(hops)]))
    return nn.ModuleList(layers)


def add_backbone(net, spec):
    """Attach conv1..stage4 to `net`; returns the channel list of the last stage."""
    net.conv1 = nn.Conv2d(3, 64, 3, 2, 1, bias=False)
    net.bn1 = nn.BatchNorm2d(64)
    net.conv2 = nn.Conv2d(64, 64, 3, 2, 1, bias=False)
    net.bn2 = nn.BatchNorm2d(64)
    net.layer1 = nn.Sequential(*[bottleneck(64 if i == 0 else 256, 64, i == 0)
                                 for i in range(4)])
    prev = [256]
    for si, st in enumerate(spec.stages):
        s = si + 2
        cur = list(st.num_channels)
        setattr(net, f'transition{s - 1}', transition(prev, cur))
        last_stage = (si == len(spec.stages) - 1)
        mods = [hr_module(cur, st.num_blocks,
                          not (last_stage and m == st.num_modules - 1))
                for m in range(st.num_modules)]
        setattr(net, f'stage{s}', nn.Sequential(*mods))
        prev = cur
    return prev


# CrowdPose-14 skeleton used by the reference's KTMachine for every config
# (/root/reference/lib/models/pose_rsgnet.py:523-573); joints are indexed
# l/r shoulder 0/1, elbow 2/3, wrist 4/5, hip 6/7, knee 8/9, ankle 10/11, head 12, neck 13.
LIMB_RULES = ((12, 13), (1, 13), (0, 13), (0, 2), (1, 3), (2, 4), (3, 5),
              (6, 7), (6, 8), (7, 9), (8, 10), (9, 11), (0, 6), (1, 7))


def limb_incidence(num_limbs, num_joints):
    m = np.zeros((num_limbs, num_joints), np.float32)
    for i, (a, b) in enumerate(LIMB_RULES):
        if i < num_limbs and a < num_joints and b < num_joints:
            m[i, a] = 1
            m[i, b] = 1
    return m


def geometry_map(feat_w, feat_h):
    """[1,4,h,w]: x/w, y/h, x-w/2, y-h/2 (pose_rsgnet.py:800-815), fp32 arithmetic."""
    xs = np.arange(feat_w, dtype=np.float32)
    ys = np.arange(feat_h, dtype=np.float32)
    g = np.zeros((4, feat_h, feat_w), np.float64)
    g[0] += (xs / feat_w)[None, :]
    g[1] += (ys / feat_h)[:, None]
    g[2] += (xs - feat_w / 2.)[None, :]
    g[3] += (ys - feat_h / 2.)[:, None]
    return torch.from_numpy(g[None].astype(np.float32))


def load_type_embeddings(num_joints, dim):
    """The reference opens 'kpt_word_embs.pkl' relative to the CWD (pose_rsgnet.py:675).  Use it
    when it is there and has the right shape; otherwise start from zeros -- checkpoints carry
    `type_features`, so load_state_dict overwrites it."""
    path = 'kpt_word_embs.pkl'
    if os.path.isfile(path):
        try:
            with open(path, 'rb') as f:
                arr = np.asarray(pickle.load(f))
            if arr.shape == (num_joints, dim):
                return torch.from_numpy(arr).float()
        except Exception:
            pass
    return torch.zeros(num_joints, dim)


def add_rsgnet_heads(net, spec, c0):
    K, L, T = spec.num_joints, spec.num_limbs, spec.type_dim
    fk = spec.final_conv_kernel
    hc = spec.head_channels
    net.multi_final_layer = nn.Conv2d(c0, K, fk, 1, 1 if fk == 3 else 0)
    net.vis_conv = conv_bn(hc, hc, 3, relu=True)
    net.type_features = nn.Parameter(load_type_embeddings(K, T))
    net.type_fc = nn.Sequential(nn.Linear(T, T, bias=False), nn.BatchNorm1d(T),
                                nn.ReLU(inplace=True))
    net.type_conv = conv_bn(T, hc, 3, relu=True)
    net.loc_features = nn.Parameter(geometry_map(spec.feat_w, spec.feat_h), requires_grad=False)
    net.loc_conv = conv_bn(4, hc, 1, relu=True)
    net.contact_conv = conv_bn(hc * 3, hc * 3, 3, relu=True)
    net.predict_contact_net = conv_bn(hc * 3, hc, 3, relu=True)

    rh = nn.Module()
    rh.g = nn.Conv2d(hc, hc, 1)
    w_tail = nn.Sequential(nn.Conv2d(hc, hc, 1), nn.GroupNorm(8, hc))
    if spec.relation_sub_sample:
        rh.W = nn.Sequential(conv_bn(hc, hc, 4, relu=True, transposed=True), w_tail)
    else:
        rh.W = w_tail
    for name, p in rh.named_parameters():
        if 'bias' in name:
            nn.init.zeros_(p)
        else:
            nn.init.normal_(p, std=1e-3)
    net.relation_head = rh

    net.kpt_net = conv_bn(hc * 2, hc, 3, relu=True)
    if spec.up_scale > 1:
        net.predict_convtranspose = conv_bn(hc, hc, spec.deconv_kernel, relu=True,
                                            transposed=True)
    net.predict_net = conv_bn(hc, hc, 3, relu=True)
    net.final_layer = nn.Conv2d(c0, K, fk, 1, 1 if fk == 3 else 0)

    wsize = c0 * fk * fk
    kt = nn.Module()
    kt.matrix_limb = nn.Parameter(torch.randn(L, K))
    kt.real_matrix_limb = nn.Parameter(torch.from_numpy(limb_incidence(L, K)),
                                       requires_grad=False)
    kt.kpt_transformer = nn.Sequential(nn.Linear(wsize, wsize), nn.LeakyReLU(0.02),
                                       nn.Linear(wsize, wsize))
    net.kt_machine = kt
    net.limbs_net = conv_bn(hc, hc, 3, relu=True)


_HEAD_GAIN = {'predict_contact_net.1.weight': 0.03, 'type_conv.1.weight': 0.3, 'limbs_net.1.weight': 0.2,
              'predict_net.1.weight': 0.3}


def _norm_gain(name):
    """Scale of a synthetic BN weight.  Eval-mode BN with synthetic running statistics does not normalise, so a
    He-initialised residual network doubles its variance at every block and quadruples it at every fuse sum: after
    8 HighResolutionModules the activations were O(1e5) and every sigmoid downstream (TRP affinity, limbs) saturated
    to a 0/1 mask -- a vacuous parity check.  Like a trained network (zero-init-residual-style small last gammas),
    the last BN of a residual branch and the BNs of the fuse paths get small weights, which keeps every stage O(1)."""
    if '.branches.' in name and name.endswith('.bn2.weight'):
        return 0.15
    if name.startswith('layer1.') and name.endswith('.bn3.weight'):
        return 0.2
    if '.fuse_layers.' in name:
        return 0.12
    return _HEAD_GAIN.get(name, 1.0)


def synth_state_dict(module, seed=0):
    """Deterministic, platform-independent, 'trained-like' weights for tests and the bench:
    He-scaled conv/linear weights, BN statistics and affine terms spread over realistic ranges so
    that folding is exercised.  Values depend only on (key, shape, seed)."""
    import zlib
    out = {}
    sd = module.state_dict()
    for name, t in sd.items():
        rs = np.random.RandomState((zlib.crc32(name.encode()) + 7919 * seed) & 0x7fffffff)
        shape = tuple(t.shape)
        leaf = name.rsplit('.', 1)[-1]
        if leaf == 'num_batches_tracked':
            out[name] = torch.zeros((), dtype=torch.long)
            continue
        if name in ('loc_features', 'kt_machine.real_matrix_limb'):
            out[name] = t.detach().clone().float()
            continue
        if leaf == 'running_mean':
            v = rs.normal(0, 0.1, shape)
        elif leaf == 'running_var':
            v = rs.uniform(0.5, 1.5, shape)
        elif name == 'type_features':
            v = rs.normal(0, 0.37, shape)
        elif name == 'kt_machine.matrix_limb':
            v = rs.normal(0, 1.0, shape)
        elif t.dim() == 1:
            # BN / GN affine or a conv/linear bias: the owning module decides
            owner = name.rsplit('.', 1)[0]
            is_norm = (owner + '.running_mean') in sd or owner.endswith('relation_head.W.1') \
                or owner.endswith('relation_head.W.1.1')
            if is_norm and leaf == 'weight':
                v = rs.uniform(0.5, 1.5, shape) * _norm_gain(name)
            else:
                v = rs.normal(0, 0.1, shape)
        else:
            fan_in = int(np.prod(shape[1:])) if t.dim() > 1 else shape[0]
            if 'convtranspose' in name or (t.dim() == 4 and 'relation_head.W.0.0' in name):
                fan_in = shape[0] * shape[2] * shape[3] // 4
            v = rs.normal(0, np.sqrt(2.0 / max(fan_in, 1)), shape)
        out[name] = torch.from_numpy(np.asarray(v, np.float32).reshape(shape))
    return out
