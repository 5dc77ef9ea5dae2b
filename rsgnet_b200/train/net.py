"""Train-mode forward of the drop-in RSGNet / HRNet modules on the tape (batch-statistics BatchNorm).

Follows /root/reference/lib/models/pose_rsgnet.py: Bottleneck :57-95, BasicBlock :25-54, HighResolutionModule :194-272,
transitions :817-856, _forward_visual_encoder :921-953, RSGNet.forward :955-1021, KTMachine.forward :592-600, and
lib/models/association.py:280-301 (SpatialRelationHead.forward); vanilla HRNet: lib/models/pose_hrnet.py:428-463.
The parameter containers are the ones of ``rsgnet_b200.models._params`` (reference names and shapes).
"""
import torch.nn as nn

from ..config import KIND_RSGNET


class Net:
    def __init__(self, module, store, tape):
        self.m, self.s, self.t = module, store, tape

    # ---- building blocks -------------------------------------------------------------------------------------------
    def conv(self, x, conv, bias=True):
        """nn.Conv2d with its own parameters (both packed forms of the weight come from the store)."""
        k = conv.kernel_size[0]
        b = self.s.raw(conv.bias) if (bias and conv.bias is not None) else None
        wp, wT = self.s.packed(conv)
        return self.t.conv(x, wp, wT, k, conv.stride[0], conv.padding[0], b)

    def bn(self, x, bn, relu=False):
        return self.t.batchnorm(x, self.s.raw(bn.weight), self.s.raw(bn.bias), bn.running_mean, bn.running_var, relu,
                                bn.eps, bn.momentum)

    def bn_add_relu(self, x, bn, res):
        return self.t.batchnorm_add_relu(x, self.s.raw(bn.weight), self.s.raw(bn.bias), bn.running_mean, bn.running_var, res,
                                         bn.eps, bn.momentum)

    def cbr(self, x, seq, relu):
        """Sequential(conv, BN[, ReLU]) as built by _params.conv_bn."""
        conv = seq[0]
        if isinstance(conv, nn.ConvTranspose2d):
            wp, wT = self.s.packed(conv)
            y = self.t.conv_transpose(x, wp, wT, conv.kernel_size[0], conv.stride[0], conv.padding[0],
                                      conv.output_padding[0])
        else:
            y = self.conv(x, conv)
        return self.bn(y, seq[1], relu)

    def bottleneck(self, x, b):
        y = self.bn(self.conv(x, b.conv1), b.bn1, True)
        y = self.bn(self.conv(y, b.conv2), b.bn2, True)
        r = self.cbr(x, b.downsample, False) if hasattr(b, 'downsample') else x
        return self.bn_add_relu(self.conv(y, b.conv3), b.bn3, r)

    def basic(self, x, b):
        y = self.bn(self.conv(x, b.conv1), b.bn1, True)
        return self.bn_add_relu(self.conv(y, b.conv2), b.bn2, x)

    def hr_module(self, xs, mod):
        nb = len(xs)
        xs = list(xs)

        def branch(b):
            x = xs[b]
            for blk in mod.branches[b]:
                x = self.basic(x, blk)
            return x
        xs = self.t.parallel([lambda b=b: branch(b) for b in range(nb)])       # independent chains: one stream each
        if nb == 1:
            return xs
        # The fused outputs are independent of each other too, but they READ the same branch results: every output gets its
        # own alias nodes (made before the fork), so that the gradients of a shared input are summed on the main stream,
        # after the join, by the aliases' link closures -- never by two streams at once.
        rows = list(mod.fuse_layers)
        alias = [[self.t.view(xs[j], *xs[j].shape) for j in range(nb)] for _ in rows]

        def fuse(i):
            row, src = rows[i], alias[i]
            terms = []
            for j in range(nb):
                if j == i:
                    terms.append(src[j])
                elif j > i:
                    terms.append(self.t.upsample_nearest(self.cbr(src[j], row[j], False), 2 ** (j - i)))
                else:
                    t = src[j]
                    hops = row[j]
                    for k, hop in enumerate(hops):
                        t = self.cbr(t, hop, relu=(k != len(hops) - 1))
                    terms.append(t)
            return self.t.add(terms, relu=True)
        return self.t.parallel([lambda i=i: fuse(i) for i in range(len(rows))])

    def transition(self, prev, layers):
        n_pre = len(prev)
        out = []
        for i, layer in enumerate(layers):
            if i < n_pre:
                out.append(self.cbr(prev[i], layer, True) if layer is not None else prev[i])
            else:
                t = prev[-1]
                for hop in layer:
                    t = self.cbr(t, hop, True)
                out.append(t)
        return out

    def backbone(self, x):
        m = self.m
        x = self.bn(self.conv(x, m.conv1), m.bn1, True)
        x = self.bn(self.conv(x, m.conv2), m.bn2, True)
        for blk in m.layer1:
            x = self.bottleneck(x, blk)
        ys = [x]
        for s in (2, 3, 4):
            xs = self.transition(ys, getattr(m, f'transition{s - 1}'))
            for mod in getattr(m, f'stage{s}'):
                xs = self.hr_module(xs, mod)
            ys = xs
        return ys

    # ---- heads -----------------------------------------------------------------------------------------------------
    def hrnet(self, x_nchw):
        t = self.t
        ys = self.backbone(t.from_nchw(x_nchw, 4))
        return t.to_nchw(self.conv(ys[0], self.m.final_layer))

    def rsgnet(self, x_nchw):
        """Returns (multi, kpt, limbs) as NCHW nodes, the relation-score node P [B,S,S] and the TRP hook."""
        t, m, s = self.t, self.m, self.s
        spec = m.spec
        feat = self.backbone(t.from_nchw(x_nchw, 4))[0]
        B, h, w, c0 = feat.shape
        multi = self.conv(feat, m.multi_final_layer)                                   # [B,h,w,K]
        vis = self.cbr(feat, m.vis_conv, True)
        # type branch (pose_rsgnet.py:968-977): scores [B,S,K] @ relu(bn1d(type_features W^T)) [K,T] is already NHWC
        tf = t.linear(s.raw(m.type_features), s.raw(m.type_fc[0].weight))
        tf = self.bn(tf, m.type_fc[1], True)
        typ = self._type_matmul(multi, tf)
        typ = self.cbr(typ, m.type_conv, True)
        # location branch (:979-980): the [1,4,h,w] coordinate map as NHWC, 1x1 conv + BN (batch of ONE) + ReLU, repeated
        loc = t.from_nchw(m.loc_features.detach(), None)
        loc = self.cbr(loc, m.loc_conv, True)
        loc = t.repeat_batch(loc, B)
        fv = t.cat([vis, typ, loc])
        fv = self.cbr(fv, m.contact_conv, True)
        fv = self.cbr(fv, m.predict_contact_net, True)
        rel, P, hook = self.trp(fv)
        kf = self.cbr(t.cat([rel, fv]), m.kpt_net, True)
        if spec.up_scale > 1:
            kf = self.cbr(kf, m.predict_convtranspose, True)
        kf = self.cbr(kf, m.predict_net, True)
        kpt = self.conv(kf, m.final_layer)
        # limbs branch (:1002-1005): the 1x1 limb filters come from the KTMachine applied to final_layer.weight
        lf = self.cbr(fv, m.limbs_net, True)
        refine = self.kt_machine()
        limbs = t.linear(lf, refine)
        if spec.up_scale > 1:
            multi = t.bilinear2x(multi)
            limbs = t.bilinear2x(limbs)
        limbs = t.sigmoid(limbs)
        return t.to_nchw(multi), t.to_nchw(kpt), t.to_nchw(limbs), P, hook

    def _type_matmul(self, multi, tf):
        """[B,h,w,K] scores x [K,T] type vectors -> [B,h,w,T] (torch.matmul at pose_rsgnet.py:974)."""
        B, h, w, K = multi.shape
        y = self.t.matmul(self.t.view(multi, B * h * w, K), tf)
        return self.t.view(y, B, h, w, tf.shape[-1])

    def trp(self, x):
        """association.py:280-301.  The relation features are the GroupNorm of W (P g): every row of P g is close to
        (S/2) mean(g) -- a mean hundreds of times its spread -- so product rounding in P g, in the 1x1 W conv or in g
        becomes O(1) noise behind the GroupNorm.  These few small products always use the 3xTF32 (fp32-class) mode."""
        t, rh = self.t, self.m.relation_head
        spec = self.m.spec
        if spec.relation_sub_sample:
            x = t.maxpool2(x)
        B, h, w, c = x.shape
        with t.precision(True):
            g = self.conv(x, rh.g)
            y, P, hook = t.trp_attention(t.view(x, B, h * w, c), t.view(g, B, h * w, c))
            y = t.view(y, B, h, w, c)
            if spec.relation_sub_sample:
                y = self.cbr(y, rh.W[0], True)
                tail = rh.W[1]
            else:
                tail = rh.W
            z = self.conv(y, tail[0])
        gn = tail[1]
        z = t.groupnorm(z, self.s.raw(gn.weight), self.s.raw(gn.bias), gn.num_groups, gn.eps)
        return z, P, hook

    def kt_machine(self):
        t, kt, s = self.t, self.m.kt_machine, self.s
        fw = self.m.final_layer.weight
        assert fw.shape[2] == 1 and fw.shape[3] == 1, 'training path: FINAL_CONV_KERNEL must be 1'
        rel = t.mul_const(s.raw(kt.matrix_limb), kt.real_matrix_limb.detach())
        x = t.matmul(rel, s.raw(fw))
        lin0, lin2 = kt.kpt_transformer[0], kt.kpt_transformer[2]
        x = t.linear(x, s.raw(lin0.weight), s.raw(lin0.bias))
        x = t.leaky_relu(x, kt.kpt_transformer[1].negative_slope)
        return t.linear(x, s.raw(lin2.weight), s.raw(lin2.bias))

    def forward(self, x_nchw):
        if self.m.spec.kind == KIND_RSGNET:
            return self.rsgnet(x_nchw)
        return self.hrnet(x_nchw)
