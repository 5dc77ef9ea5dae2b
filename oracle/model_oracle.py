"""CPU oracle for the RSGNet / HRNet per-crop forward.   *** TEST INFRASTRUCTURE ONLY ***

A functional fp32 restatement (torch CPU ops) of the reference's inference forward, driven
directly by a reference-named ``state_dict``.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s cpu_baseline / ``--impl reference`` leg may import this file; the product
package ``rsgnet_b200`` never does.

Parity pin: ``oracle/gen_golden.py`` runs the UNMODIFIED reference (imported from
/root/reference with the three import shims of SURVEY.md App. C) on the same seeded weights
and inputs and commits its outputs under ``tests/golden/``; ``tests/test_oracle_golden.py``
checks this restatement against them.  The reference ships no golden vectors of its own
(SURVEY.md §4), so the reference-executed fixtures are the pin.

Reference lines followed (paths under /root/reference):
  lib/models/pose_rsgnet.py:57-95    Bottleneck
  lib/models/pose_rsgnet.py:25-54    BasicBlock
  lib/models/pose_rsgnet.py:194-272  HighResolutionModule fuse rule + forward
  lib/models/pose_rsgnet.py:817-856  transition layers
  lib/models/pose_rsgnet.py:921-953  _forward_visual_encoder
  lib/models/pose_rsgnet.py:955-1021 RSGNet.forward (heads, TRP call, SGM)
  lib/models/pose_rsgnet.py:592-600  KTMachine.forward
  lib/models/association.py:280-301  SpatialRelationHead.forward (TRP)
  lib/models/pose_hrnet.py:428-463   vanilla HRNet forward
"""
import torch
import torch.nn.functional as F

EPS = 1e-5


def _get(cfg, *path):
    cur = cfg
    for p in path:
        cur = cur[p] if isinstance(cur, dict) else getattr(cur, p)
    return cur


class _W:
    """state_dict accessor with a name prefix."""

    def __init__(self, sd, prefix=""):
        self.sd, self.prefix = sd, prefix

    def sub(self, name):
        return _W(self.sd, f"{self.prefix}{name}.")

    def has(self, name):
        return f"{self.prefix}{name}" in self.sd

    def __getitem__(self, name):
        return self.sd[f"{self.prefix}{name}"].to(DTYPE)


DTYPE = torch.float32  # oracle/train_oracle.py runs the same graph in float64 as a yard-stick
TRAIN = False        # oracle/train_oracle.py flips this: batch statistics + running-stat updates (momentum 0.1), as
                     # nn.BatchNorm2d(momentum=BN_MOMENTUM) does in train mode (pose_rsgnet.py:16, 32-36)


def _bn(w, x):
    if TRAIN:
        return F.batch_norm(x, w.sd[w.prefix + "running_mean"], w.sd[w.prefix + "running_var"], w["weight"], w["bias"],
                            True, 0.1, EPS)
    return F.batch_norm(x, w["running_mean"], w["running_var"], w["weight"], w["bias"],
                        False, 0.0, EPS)


def _conv_bn(w, conv, bn, x, stride=1, pad=None, relu=False):
    k = w.sub(conv)["weight"]
    if pad is None:
        pad = k.shape[-1] // 2
    y = _bn(w.sub(bn), F.conv2d(x, k, None, stride, pad))
    return F.relu(y) if relu else y


def _seq_cbr(w, x, stride=1, relu=True):
    """Sequential(Conv(bias=False), BN[, ReLU]) stored as indices 0, 1."""
    return _conv_bn(w, "0", "1", x, stride=stride, relu=relu)


def _bottleneck(w, x):
    y = _conv_bn(w, "conv1", "bn1", x, relu=True)
    y = _conv_bn(w, "conv2", "bn2", y, relu=True)
    y = _conv_bn(w, "conv3", "bn3", y)
    r = _seq_cbr(w.sub("downsample"), x, relu=False) if w.has("downsample.0.weight") else x
    return F.relu(y + r)


def _basic(w, x):
    y = _conv_bn(w, "conv1", "bn1", x, relu=True)
    y = _conv_bn(w, "conv2", "bn2", y)
    return F.relu(y + x)


def _hr_module(w, xs, n_out):
    nb = len(xs)
    xs = list(xs)
    for b in range(nb):
        i = 0
        while w.has(f"branches.{b}.{i}.conv1.weight"):
            xs[b] = _basic(w.sub(f"branches.{b}.{i}"), xs[b])
            i += 1
    if nb == 1:
        return xs
    outs = []
    for i in range(n_out):
        acc = None
        for j in range(nb):
            if j == i:
                t = xs[j]
            elif j > i:
                t = _seq_cbr(w.sub(f"fuse_layers.{i}.{j}"), xs[j], relu=False)
                t = F.interpolate(t, scale_factor=2 ** (j - i), mode="nearest")
            else:
                t = xs[j]
                for k in range(i - j):
                    t = _seq_cbr(w.sub(f"fuse_layers.{i}.{j}.{k}"), t, stride=2,
                                 relu=(k != i - j - 1))
            acc = t if acc is None else acc + t
        outs.append(F.relu(acc))
    return outs


def _transition(w, prev, n_cur):
    """prev: list of tensors from the previous stage; returns the n_cur inputs of the next."""
    n_pre = len(prev)
    out = []
    for i in range(n_cur):
        if i < n_pre:
            if w.has(f"{i}.0.weight"):
                out.append(_seq_cbr(w.sub(f"{i}"), prev[i]))
            else:
                out.append(prev[i])
        else:
            t = prev[-1]
            for j in range(i + 1 - n_pre):
                t = _seq_cbr(w.sub(f"{i}.{j}"), t, stride=2)
            out.append(t)
    return out


def backbone(sd, cfg, x, stages=None):
    """HRNet encoder.  Returns the last stage's output list.  `stages` (dict) collects taps."""
    w = _W(sd)
    extra = _get(cfg, "MODEL", "EXTRA")
    x = _conv_bn(w, "conv1", "bn1", x, stride=2, relu=True)
    x = _conv_bn(w, "conv2", "bn2", x, stride=2, relu=True)
    if stages is not None:
        stages["stem"] = x
    for i in range(4):
        x = _bottleneck(w.sub(f"layer1.{i}"), x)
    if stages is not None:
        stages["layer1"] = x
    ys = [x]
    for s in (2, 3, 4):
        sc = extra[f"STAGE{s}"]
        nb, nm = int(sc["NUM_BRANCHES"]), int(sc["NUM_MODULES"])
        xs = _transition(w.sub(f"transition{s - 1}"), ys, nb)
        for m in range(nm):
            last = (s == 4 and m == nm - 1)
            xs = _hr_module(w.sub(f"stage{s}.{m}"), xs, 1 if last else nb)
        ys = xs
        if stages is not None:
            for b, t in enumerate(ys):
                stages[f"stage{s}.{b}"] = t
    return ys


def hrnet_forward(sd, cfg, x, stages=None):
    """lib/models/pose_hrnet.py:428-463: backbone + final_layer (with bias)."""
    ys = backbone(sd, cfg, x, stages)
    w = _W(sd)
    k = w["final_layer.weight"]
    return F.conv2d(ys[0], k, w["final_layer.bias"], 1, k.shape[-1] // 2)


def trp(w, x, sub_sample):
    """lib/models/association.py:280-301 (sigmoid non-local block)."""
    B, C = x.shape[:2]
    if sub_sample:
        x = F.max_pool2d(x, 2)
    g = F.conv2d(x, w["g.weight"], w["g.bias"]).view(B, C, -1).permute(0, 2, 1)
    th = x.reshape(B, C, -1)
    score = torch.sigmoid(th.permute(0, 2, 1) @ th)
    y = (score @ g).permute(0, 2, 1).reshape(B, C, *x.shape[2:])
    if sub_sample:
        up = w.sub("W.0")
        y = F.relu(_bn(up.sub("1"), F.conv_transpose2d(y, up["0.weight"], None, 2, 1)))
        wz = w.sub("W.1")
    else:
        wz = w.sub("W")
    z = F.conv2d(y, wz["0.weight"], wz["0.bias"])
    z = F.group_norm(z, 8, wz["1.weight"], wz["1.bias"], EPS)
    return z, score


def kt_machine(w, final_w):
    """lib/models/pose_rsgnet.py:592-600."""
    n_out, n_in, kh, kw = final_w.shape
    m = w["matrix_limb"] * w["real_matrix_limb"]
    t = m @ final_w.reshape(n_out, -1)
    t = F.linear(t, w["kpt_transformer.0.weight"], w["kpt_transformer.0.bias"])
    t = F.leaky_relu(t, 0.02)
    t = F.linear(t, w["kpt_transformer.2.weight"], w["kpt_transformer.2.bias"])
    return t.reshape(t.shape[0], n_in, kh, kw)


def rsgnet_forward(sd, cfg, x, stages=None, relation_target=None):
    """lib/models/pose_rsgnet.py:955-1021.  Returns the reference's 4-tuple."""
    w = _W(sd)
    up_scale = int(_get(cfg, "MODEL", "UP_SCALE"))
    sub = bool(_get(cfg, "MODEL", "RELATION_SUB_SAMPLE"))
    ys = backbone(sd, cfg, x, stages)
    feat = ys[0]
    B = feat.shape[0]

    kfin = w["multi_final_layer.weight"]
    multi = F.conv2d(feat, kfin, w["multi_final_layer.bias"], 1, kfin.shape[-1] // 2)
    vis = _seq_cbr(w.sub("vis_conv"), feat)

    # type branch: scores^T . relu(bn1d(type_features W^T)) spread over space, then 3x3 conv
    tf = F.linear(w["type_features"], w["type_fc.0.weight"])
    tf = F.relu(_bn(w.sub("type_fc.1"), tf))
    Kj, H, Wd = multi.shape[1:]
    t = multi.reshape(B, Kj, H * Wd).permute(0, 2, 1) @ tf            # B,S,T
    t = t.permute(0, 2, 1).reshape(B, tf.shape[1], H, Wd)
    typ = _seq_cbr(w.sub("type_conv"), t)

    loc = _seq_cbr(w.sub("loc_conv"), w["loc_features"]).repeat(B, 1, 1, 1)

    fv = torch.cat((vis, typ, loc), 1)
    fv = _seq_cbr(w.sub("contact_conv"), fv)
    fv = _seq_cbr(w.sub("predict_contact_net"), fv)

    rel, rel_scores = trp(w.sub("relation_head"), fv, sub)

    kf = _seq_cbr(w.sub("kpt_net"), torch.cat((rel, fv), 1))
    if up_scale > 1:
        dw = w["predict_convtranspose.0.weight"]
        dk = dw.shape[-1]
        pad, opad = {4: (1, 0), 3: (1, 1), 2: (0, 0)}[dk]
        kf = F.conv_transpose2d(kf, dw, None, 2, pad, opad)
        kf = F.relu(_bn(w.sub("predict_convtranspose.1"), kf))
    kf = _seq_cbr(w.sub("predict_net"), kf)
    kfin2 = w["final_layer.weight"]
    kpt = F.conv2d(kf, kfin2, w["final_layer.bias"], 1, kfin2.shape[-1] // 2)

    lf = _seq_cbr(w.sub("limbs_net"), fv)
    limbs = F.conv2d(lf, kt_machine(w.sub("kt_machine"), kfin2))
    if up_scale > 1:
        multi = F.interpolate(multi, scale_factor=2, mode="bilinear", align_corners=True)
        limbs = F.interpolate(limbs, scale_factor=2, mode="bilinear", align_corners=True)
    limbs_logits = limbs
    limbs = torch.sigmoid(limbs)
    if relation_target is not None:
        rel_scores = ((relation_target - rel_scores) ** 2).mean(dim=(1, 2))
    if stages is not None:
        stages.update(vis=vis, type=typ, final_vis=fv, relation=rel, kpt_feat=kf, limbs_logits=limbs_logits)
    return multi, kpt, limbs, rel_scores


def forward(sd, cfg, x, **kw):
    name = _get(cfg, "MODEL", "NAME")
    with torch.no_grad():
        if name == "pose_hrnet":
            return hrnet_forward(sd, cfg, x, **kw)
        return rsgnet_forward(sd, cfg, x, **kw)
