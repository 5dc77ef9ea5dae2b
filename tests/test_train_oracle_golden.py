"""The training-step oracle (oracle/train_oracle.py) against fixtures produced by the UNMODIFIED reference loop
(oracle/gen_golden.py train: lib/core/function.py:rsgnet_train / :train, one iteration, torch.optim.Adam) -- CPU only."""
import os

import numpy as np
import pytest
import torch

from oracle import train_oracle
from rsgnet_b200 import presets, synth
from rsgnet_b200.models import _params, pose_hrnet, pose_rsgnet

GOLD = os.path.join(os.path.dirname(__file__), 'golden')


def load_case(key):
    g = np.load(os.path.join(GOLD, f'train_{key}.npz'), allow_pickle=False)
    cfg = presets.preset(key)
    name = cfg.MODEL.NAME
    net = (pose_rsgnet if name == 'pose_rsgnet' else pose_hrnet).get_pose_net(cfg, False)
    sd = _params.synth_state_dict(net, seed=int(g['seed']))
    b = synth.train_batch(int(g['batch']), cfg.MODEL.IMAGE_SIZE, cfg.MODEL.HEATMAP_SIZE, cfg.MODEL.NUM_JOINTS,
                          max(int(cfg.MODEL.NUM_LIMBS), 1), seed=int(g['seed']))
    return g, cfg, net, sd, b


def flat_sub(tensors, names, sub):
    return torch.cat([tensors[k].reshape(-1).float() for k in names])[::sub].numpy()


@pytest.mark.parametrize('key', ['tiny', 'tiny_cp', 'tiny_hrnet', 'w32_coco'])
def test_train_oracle_fp64_equals_reference_loop_fp64(key):
    """Logic pin: in double precision the restatement reproduces the reference loop to rounding."""
    g, cfg, net, sd, b = load_case(key)
    torch.set_num_threads(8)
    batch = {k: torch.from_numpy(v) for k, v in b.items()}
    losses, grads, buffers, _ = train_oracle.forward_backward(sd, cfg, batch, dtype=torch.float64)
    names = [str(n) for n in g['names']]
    assert sorted(names) == sorted(grads), 'trainable parameter set differs from the reference'
    if 'losses64' in g.files:
        ours = np.array([losses['multi_loss'], losses['target_loss'], losses['skeleton_loss'], losses['relation_loss']])
        np.testing.assert_allclose(ours, g['losses64'], rtol=1e-7)          # the reference prints ~8 digits
    norms = np.array([float(grads[k].norm()) for k in names])
    np.testing.assert_allclose(norms, g['grad_norm64'], rtol=1e-6, atol=1e-12)
    gs = torch.cat([grads[k].reshape(-1) for k in names])[::int(g['sub'])].numpy()
    assert np.abs(gs - g['grad_sub64']).max() <= 1e-7 * np.abs(g['grad_sub64']).max()
    for k in g.files:
        if k.startswith('grad64.'):
            assert np.abs(grads[k[7:]].numpy() - g[k]).max() <= 1e-7 * max(np.abs(g[k]).max(), 1e-30), k
    bnames = [str(n) for n in g['buf_names']]
    bs = torch.cat([buffers[k].reshape(-1) for k in bnames])[::7].numpy()
    np.testing.assert_allclose(bs, g['buf_sub64'], rtol=1e-9, atol=1e-12)


@pytest.mark.parametrize('key', ['tiny', 'tiny_cp', 'tiny_hrnet', 'w32_coco'])
def test_train_oracle_fp32_within_reference_fp32_noise(key):
    """In fp32 two correct implementations differ by rounding, amplified by the GroupNorm behind the TRP ('tiny' is the
    ill-conditioned case: the reference's own fp32 gradients are ~1e-3 .. 1e-2 off its fp64 run).  Bar: per parameter
    tensor, distance to the fp64 reference run <= 20 x the reference's own fp32 distance + 1e-5 of the norm (a noise level,
    not a bias: the fp64 test above pins the logic exactly)."""
    g, cfg, net, sd, b = load_case(key)
    torch.set_num_threads(8)
    batch = {k: torch.from_numpy(v) for k, v in b.items()}
    losses, grads, buffers, _ = train_oracle.forward_backward(sd, cfg, batch)
    names = [str(n) for n in g['names']]
    if 'losses' in g.files:
        ours = np.array([losses['multi_loss'], losses['target_loss'], losses['skeleton_loss'], losses['relation_loss']])
        np.testing.assert_allclose(ours, g['losses64'], rtol=5e-4)
    # distance to the fp64 run, measured on the committed sub-sample and on the per-tensor norms
    sub = int(g['sub'])
    gs = flat_sub(grads, names, sub)
    ref_noise = np.abs(g['grad_sub'].astype(np.float64) - g['grad_sub64']).max()
    ours_noise = np.abs(gs.astype(np.float64) - g['grad_sub64']).max()
    assert ours_noise <= 20 * ref_noise + 1e-5 * np.abs(g['grad_sub64']).max(), (ours_noise, ref_noise)
    norms = np.array([float(grads[k].double().norm()) for k in names])
    rho = (g['grad_err32'] / np.maximum(g['grad_norm64'], 1e-30)).max()     # the reference's worst per-tensor fp32 error
    assert (np.abs(norms - g['grad_norm64']) <= (50 * rho + 1e-5) * g['grad_norm64'] + 1e-6 * g['grad_norm64'].max()).all()
    # Adam (lib/utils/utils.py:70-74): the first step moves every weight by -lr * g / (|g| + eps)
    new, _ = train_oracle.adam_step(sd, grads)
    delta = {k: new[k] - sd[k].float() for k in names}
    ds = flat_sub(delta, names, sub)
    big = np.abs(g['grad_sub64']) > 1e-5         # where |g| ~ eps the rounding noise decides the step
    assert np.abs(ds - g['delta_sub'])[big].max() <= 2e-5
    bnames = [str(n) for n in g['buf_names']]
    bs = torch.cat([buffers[k].reshape(-1) for k in bnames])[::7].numpy()
    np.testing.assert_allclose(bs, g['buf_sub'], rtol=1e-3, atol=2e-4)      # BNs behind the TRP's GroupNorm see its fp32 noise
    assert int(g['num_batches_tracked']) == 1
