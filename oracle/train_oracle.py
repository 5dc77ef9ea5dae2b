"""CPU oracle for the training step.   *** TEST INFRASTRUCTURE ONLY ***

A functional fp32 restatement (torch CPU ops + torch autograd) of one iteration of the reference's
``rsgnet_train`` loop, driven by a reference-named ``state_dict``.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s cpu_baseline / ``--impl reference`` legs may import this file.

Parity pin: ``oracle/gen_golden.py`` executes the UNMODIFIED ``lib/core/function.py:rsgnet_train`` on the unmodified
reference model (one batch, CPU) and commits its printed losses, gradients, updated parameters and BN running statistics as
``tests/golden/train_*.npz``; ``tests/test_train_oracle_golden.py`` checks this restatement against them.

Reference lines followed (paths under /root/reference):
  lib/core/function.py:256-269   relation target = outer product of the down-sampled person mask
  lib/core/function.py:271       model(input, relation_target)  (train mode: batch-statistics BatchNorm)
  lib/models/pose_rsgnet.py:1014-1018  relation_scores -> per-sample MSE
  lib/core/function.py:283-313   target / multi / skeleton (0.01 BCE) / relation (0.001 mean) losses
  lib/core/loss.py:14-38         JointsMSELoss
  lib/utils/utils.py:70-74       Adam(lr)
"""
import torch
import torch.nn.functional as F

from . import model_oracle

FROZEN = ('loc_features', 'kt_machine.real_matrix_limb')


def _is_buffer(name):
    leaf = name.rsplit('.', 1)[-1]
    return leaf in ('running_mean', 'running_var', 'num_batches_tracked')


def joints_mse(output, target, target_weight):
    """lib/core/loss.py:20-38 with use_target_weight=True."""
    B, K = output.shape[:2]
    pred = output.reshape(B, K, -1)
    gt = target.reshape(B, K, -1)
    loss = 0
    for k in range(K):
        w = target_weight[:, k]
        loss = loss + 0.5 * F.mse_loss(pred[:, k] * w, gt[:, k] * w)
    return loss / K


def relation_target(target):
    """lib/core/function.py:256-269."""
    person, _ = torch.max(target, dim=1)
    b, h, w = person.shape
    person = F.interpolate(person.reshape(b, 1, h, w), scale_factor=1 / 2, mode='bilinear', align_corners=True)
    person = person.reshape(b, 1, -1)
    return torch.matmul(person.permute(0, 2, 1), person)


def forward_backward(sd, cfg, batch, dtype=torch.float32):
    """Returns (losses, grads {name: tensor}, new running buffers {name: tensor}, outputs).  `batch`: dict of torch CPU
    tensors (keys of rsgnet_b200.synth.train_batch).  dtype=torch.float64 runs the same graph in double precision."""
    params = {}
    work = {}
    batch = {k: v.to(dtype) for k, v in batch.items()}
    for k, v in sd.items():
        if _is_buffer(k):
            work[k] = v.clone().to(dtype) if v.is_floating_point() else v.clone()
        else:
            params[k] = v.clone().to(dtype).requires_grad_(k not in FROZEN)
            work[k] = params[k]
    model_oracle.DTYPE = dtype
    name = model_oracle._get(cfg, 'MODEL', 'NAME')
    x = batch['input']
    model_oracle.TRAIN = True
    try:
        if name == 'pose_hrnet':
            out = model_oracle.hrnet_forward(work, cfg, x)
            loss = joints_mse(out, batch['target'], batch['target_weight'])
            losses = dict(target_loss=float(loss.detach()), loss=float(loss.detach()))
            outputs = (out.detach(),)
        else:
            rt = relation_target(batch['target'])
            multi, kpt, limbs, rel = model_oracle.rsgnet_forward(work, cfg, x, relation_target=rt)
            target_loss = joints_mse(kpt, batch['target'], batch['target_weight'])
            multi_loss = joints_mse(multi, batch['all_ins_target'], batch['all_ins_target_weight'])
            skel = 0.01 * F.binary_cross_entropy(limbs, batch['target_limbs'])
            relation = 0.001 * torch.mean(rel)
            loss = multi_loss + target_loss + skel + relation
            losses = dict(multi_loss=float(multi_loss.detach()), target_loss=float(target_loss.detach()), skeleton_loss=float(skel.detach()),
                          relation_loss=float(relation.detach()), loss=float(loss.detach()))
            outputs = (multi.detach(), kpt.detach(), limbs.detach(), rel.detach())
        loss.backward()
    finally:
        model_oracle.TRAIN = False
        model_oracle.DTYPE = torch.float32
    grads = {k: (p.grad.detach() if p.grad is not None else torch.zeros_like(p)) for k, p in params.items() if p.requires_grad}
    buffers = {k: v for k, v in work.items() if _is_buffer(k)}
    return losses, grads, buffers, outputs


def adam_step(sd, grads, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, state=None):
    """torch.optim.Adam's update written out (first moment m, second moment v, bias corrections); returns (new params, state)."""
    state = state or dict(t=0, m={}, v={})
    state['t'] += 1
    t = state['t']
    bc1, bc2 = 1 - betas[0] ** t, 1 - betas[1] ** t
    new = {}
    for k, g in grads.items():
        m = state['m'].get(k, torch.zeros_like(g)) * betas[0] + (1 - betas[0]) * g
        v = state['v'].get(k, torch.zeros_like(g)) * betas[1] + (1 - betas[1]) * g * g
        state['m'][k], state['v'][k] = m, v
        new[k] = sd[k].float() - (lr / bc1) * m / (v.sqrt() / (bc2 ** 0.5) + eps)
    return new, state
