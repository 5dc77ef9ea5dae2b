"""The training step on the GPU (rsgnet_b200.train.TrainStep and the drop-in `model.train(); model(x, rt); loss.backward()`
route) against fixtures produced by the UNMODIFIED reference loop (lib/core/function.py:rsgnet_train / :train + Adam,
oracle/gen_golden.py train) -- losses, every parameter gradient, the Adam update and the BN running statistics.

Yard-stick: the reference loop run in float64.  The reference's own fp32 run is `ref noise` away from it (the GroupNorm
behind the TRP amplifies rounding; 'tiny' is the ill-conditioned case, 'tiny_cp' / 'tiny_hrnet' are well conditioned), and
the bar for this implementation in its fp32-class mode (`precise`: 3xTF32 products) is a multiple of that noise."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), 'golden')


def setup_case(key):
    from rsgnet_b200 import presets, synth
    from rsgnet_b200.models import _params, pose_hrnet, pose_rsgnet
    g = np.load(os.path.join(GOLD, f'train_{key}.npz'), allow_pickle=False)
    cfg = presets.preset(key)
    name = cfg.MODEL.NAME
    net = (pose_rsgnet if name == 'pose_rsgnet' else pose_hrnet).get_pose_net(cfg, True)
    sd = _params.synth_state_dict(net, seed=int(g['seed']))
    net.load_state_dict(sd)
    torch.cuda.set_device(0)
    net = net.cuda().train()
    b = synth.train_batch(int(g['batch']), cfg.MODEL.IMAGE_SIZE, cfg.MODEL.HEATMAP_SIZE, cfg.MODEL.NUM_JOINTS,
                          max(int(cfg.MODEL.NUM_LIMBS), 1), seed=int(g['seed']))
    batch = {k: torch.from_numpy(v).cuda() for k, v in b.items()}
    return g, cfg, net, sd, batch


def step_args(batch, rsg):
    if rsg:
        return (batch['input'], batch['target'], batch['target_weight'], batch['all_ins_target'],
                batch['all_ins_target_weight'], batch['target_limbs'])
    return (batch['input'], batch['target'], batch['target_weight'])


def grads_of(net, store, names):
    P = dict(net.named_parameters())
    out = {}
    for k in names:
        off, n = store.offsets[id(P[k])]
        out[k] = store.flat_g[off:off + n].detach().cpu().double().reshape(P[k].shape)
    return out


@pytest.mark.parametrize('key', ['tiny', 'tiny_cp', 'tiny_hrnet', 'w32_coco'])
def test_train_step_precise_vs_reference_loop(key):
    from rsgnet_b200.train import TrainStep
    g, cfg, net, sd, batch = setup_case(key)
    rsg = cfg.MODEL.NAME == 'pose_rsgnet'
    ts = TrainStep(net, lr=1e-3, precise=True)
    losses, outs = ts.forward_backward(*step_args(batch, rsg))
    L = losses.read()
    names = [str(n) for n in g['names']]
    if rsg:
        ours = np.array([L['multi_loss'], L['target_loss'], L['skeleton_loss'], L['relation_loss']])
        ref_err = np.abs(g['losses'] - g['losses64']) / g['losses64']
        assert (np.abs(ours - g['losses64']) / g['losses64'] <= 20 * ref_err + 2e-5).all(), (ours, g['losses64'])
    grads = grads_of(net, ts.store, names)
    sub = int(g['sub'])
    gs = torch.cat([grads[k].reshape(-1) for k in names])[::sub].numpy()
    scale = np.abs(g['grad_sub64']).max()
    ref_noise = np.abs(g['grad_sub'].astype(np.float64) - g['grad_sub64']).max()
    ours_noise = np.abs(gs - g['grad_sub64']).max()
    d = np.abs(gs - g['grad_sub64'])
    p90 = np.percentile(d, 90)
    ref_p90 = np.percentile(np.abs(g['grad_sub'].astype(np.float64) - g['grad_sub64']), 90)
    print(f'{key}: gradient distance to the fp64 reference run / max|g|: ours max {ours_noise / scale:.2e} p90 {p90 / scale:.2e}, '
          f'reference fp32 max {ref_noise / scale:.2e} p90 {ref_p90 / scale:.2e}')
    # Bulk (90th percentile): a multiple of the reference's own fp32 noise.  Maximum: 5e-3 -- a ReLU whose pre-activation
    # is below rounding flips its mask in one implementation and not in the other, which moves the gradients upstream of
    # that pixel by ~1e-3 of their maximum (seen on transition1.0 of 'tiny_cp'; every tensor downstream stayed at 1e-6).
    assert p90 <= 20 * ref_p90 + 2e-6 * scale, (p90, ref_p90, scale)
    assert ours_noise <= 20 * ref_noise + 5e-3 * scale, (ours_noise, ref_noise, scale)
    norms = np.array([float(grads[k].norm()) for k in names])
    rho = (g['grad_err32'] / np.maximum(g['grad_norm64'], 1e-30)).max()
    relerr = np.abs(norms - g['grad_norm64']) / np.maximum(g['grad_norm64'], 1e-6 * g['grad_norm64'].max())
    print(f'{key}: per-tensor gradient-norm error: median {np.median(relerr):.2e} max {relerr.max():.2e} (reference fp32 max {rho:.2e})')
    assert np.median(relerr) <= 5 * rho + 1e-5 and relerr.max() <= 50 * rho + 5e-3, (np.median(relerr), relerr.max(), rho)
    for k in g.files:
        if k.startswith('grad64.'):
            ref = g[k]
            err = np.abs(grads[k[7:]].numpy() - ref).max() / max(np.abs(ref).max(), 1e-30)
            noise = np.abs(g['grad.' + k[7:]] - ref).max() / max(np.abs(ref).max(), 1e-30)
            assert err <= 20 * noise + 5e-3, (k, err, noise)
    # Adam: lib/utils/utils.py:70-74; the first step moves every weight by -lr g / (|g| + eps)
    before = {k: p.detach().clone() for k, p in net.named_parameters()}
    ts.all_reduce()
    ts.adam()
    torch.cuda.synchronize()
    after = dict(net.named_parameters())
    ds = torch.cat([(after[k].detach() - before[k]).reshape(-1) for k in names])[::sub].cpu().numpy()
    big = np.abs(g['grad_sub64']) > 10 * (ours_noise + ref_noise) + 1e-5      # below the noise the SIGN of g decides the step
    assert big.sum() > 10 and np.abs(ds - g["delta_sub"])[big].max() <= 3e-5
    for k in ('loc_features', 'kt_machine.real_matrix_limb'):
        if k in after:
            assert torch.equal(after[k], before[k]), 'frozen parameter moved'
    # BN running statistics (momentum 0.1, unbiased variance) and the step counter
    new_sd = net.state_dict()
    bnames = [str(n) for n in g['buf_names']]
    bs = torch.cat([new_sd[k].reshape(-1) for k in bnames])[::7].cpu().numpy()
    np.testing.assert_allclose(bs, g['buf_sub64'], rtol=2e-3, atol=3e-4)
    assert int(new_sd['bn1.num_batches_tracked']) == 1
    assert ts.last_launches > 100


@pytest.mark.parametrize('key', ['tiny_cp', 'tiny_hrnet'])
def test_train_step_tf32_mode(key):
    """The production mode (TF32 products, fp32 accumulation -- what the reference's cuDNN convolutions do on a GPU by
    default) on the two well-conditioned cases: losses within 1e-3, median per-tensor gradient norm within 2 %."""
    from rsgnet_b200.train import TrainStep
    g, cfg, net, sd, batch = setup_case(key)
    rsg = cfg.MODEL.NAME == 'pose_rsgnet'
    ts = TrainStep(net, lr=1e-3, precise=False)
    losses, _ = ts.forward_backward(*step_args(batch, rsg))
    L = losses.read()
    names = [str(n) for n in g['names']]
    if rsg:
        ours = np.array([L['multi_loss'], L['target_loss'], L['skeleton_loss'], L['relation_loss']])
        np.testing.assert_allclose(ours, g['losses64'], rtol=2e-3)
    grads = grads_of(net, ts.store, names)
    sub = int(g['sub'])
    gs = torch.cat([grads[k].reshape(-1) for k in names])[::sub].numpy()
    scale = np.abs(g['grad_sub64']).max()
    err = np.abs(gs - g['grad_sub64']).max() / scale
    norms = np.array([float(grads[k].norm()) for k in names])
    big = g['grad_norm64'] > 1e-4 * g['grad_norm64'].max()
    relerr = (np.abs(norms - g['grad_norm64']) / g['grad_norm64'])[big]
    print(f'{key}: TF32-mode gradient error {err:.2e} of max |g|; per-tensor norm error median {np.median(relerr):.2e} max {relerr.max():.2e}')
    # 10-bit products leave ~3e-4 of noise per layer: pre-activations near zero flip their ReLU masks (hundreds per step),
    # which is what bounds the agreement of any TF32 run -- the reference's own cuDNN path included -- with an fp64 run
    assert err < 0.1 and np.median(relerr) < 0.02 and relerr.max() < 0.25


def test_dropin_autograd_route_equals_train_step():
    """lib/core/function.py:271-319 as the reference writes it: model.train(); outputs = model(input, relation_target);
    torch criteria on the outputs; loss.backward().  The gradients that reach p.grad equal TrainStep's."""
    from oracle import train_oracle
    from rsgnet_b200.train import TrainStep
    g, cfg, net, sd, batch = setup_case('tiny_cp')
    ts = TrainStep(net, lr=1e-3, precise=False)
    losses, _ = ts.forward_backward(*step_args(batch, True))
    L = losses.read()
    names = [str(n) for n in g['names']]
    want = grads_of(net, ts.store, names)
    net.load_state_dict(sd)                     # reset the BN statistics; parameters are untouched (no Adam step ran)
    rt = train_oracle.relation_target(batch['target'])
    multi, kpt, limbs, rel = net(batch['input'], rt)
    assert rel.shape == (batch['input'].shape[0],) and kpt.requires_grad
    target_loss = train_oracle.joints_mse(kpt, batch['target'], batch['target_weight'])
    multi_loss = train_oracle.joints_mse(multi, batch['all_ins_target'], batch['all_ins_target_weight'])
    skel = 0.01 * torch.nn.functional.binary_cross_entropy(limbs, batch['target_limbs'])
    relation = 0.001 * torch.mean(rel)
    loss = multi_loss + target_loss + skel + relation
    assert abs(float(loss) - L['loss']) <= 1e-4 * L['loss']
    loss.backward()
    P = dict(net.named_parameters())
    for k in names:
        assert P[k].grad is not None, k
        ref = want[k].numpy()
        err = np.abs(P[k].grad.detach().cpu().double().numpy() - ref).max()
        # the two routes feed the backward pass with losses / relation targets that differ in the last fp32 bit; the TRP's
        # parameters see that through a ~1e4 cancellation, everything else is far below the bar
        bar = 2e-2 if k.startswith('relation_head.') else 2e-3
        assert err <= bar * max(np.abs(ref).max(), 1e-12) + 1e-9, (k, err, np.abs(ref).max())
    assert P['loc_features'].grad is None
    # the library's Adam through the torch.optim interface (the reference's optimizer.step())
    from rsgnet_b200.train import FusedAdam
    opt = FusedAdam([p for p in net.parameters() if p.requires_grad], lr=1e-3)
    w0 = P['final_layer.weight'].detach().clone()
    opt.step()
    d = (P['final_layer.weight'].detach() - w0).abs()
    assert float(d.max()) <= 1.0001e-3 and float(d.max()) > 5e-4


def test_training_reduces_the_loss_and_eval_follows_the_new_weights():
    from rsgnet_b200.train import TrainStep
    g, cfg, net, sd, batch = setup_case('tiny_cp')
    ts = TrainStep(net, lr=1e-3)
    hist = []
    for _ in range(6):
        L, _ = ts(*step_args(batch, True))
        hist.append(L['loss'])
    assert all(np.isfinite(hist)) and hist[-1] < hist[0], hist
    assert int(net.state_dict()['bn1.num_batches_tracked']) == 6
    net.eval()
    with torch.no_grad():
        out = net(batch['input'])
    assert out[1].shape[1] == cfg.MODEL.NUM_JOINTS and torch.isfinite(out[1]).all()
    # the inference engine was rebuilt from the TRAINED parameters: compare with the CPU oracle on the new state_dict
    from oracle import model_oracle
    new_sd = {k: v.detach().cpu() for k, v in net.state_dict().items()}
    ref = model_oracle.forward(new_sd, cfg, batch['input'].cpu())[1]
    err = float((out[1].cpu() - ref).abs().max() / ref.abs().max())
    assert err < 0.05, err


def test_train_mode_with_trp_sub_sample_vs_oracle_autograd():
    """RELATION_SUB_SAMPLE (association.py:221, 252-266: max-pool before the attention, ConvTranspose + BN + ReLU after) has no
    reference-loop fixture -- the reference's loop cannot train it (its relation target has the full-map size) -- so the
    train-mode forward and backward of that variant are checked against torch autograd over the oracle in float64: the
    drop-in route without a relation target, a random linear functional of all four outputs as the loss."""
    from oracle import model_oracle
    from rsgnet_b200 import presets, synth
    from rsgnet_b200.models import _params, pose_rsgnet
    cfg = presets.preset('tiny_cp_sub')
    net = pose_rsgnet.get_pose_net(cfg, True)
    sd = _params.synth_state_dict(net, seed=1)
    net.load_state_dict(sd)
    torch.cuda.set_device(0)
    net = net.cuda().train()
    x = torch.from_numpy(synth.crops(2, cfg.MODEL.IMAGE_SIZE, seed=12))
    outs = net(x.cuda())
    rs = np.random.RandomState(3)
    R = [torch.from_numpy(rs.standard_normal(tuple(o.shape)).astype(np.float32)) for o in outs]
    loss = sum((o * r.cuda()).sum() for o, r in zip(outs, R))
    loss.backward()
    # oracle: the same graph in float64 with batch-statistics BN
    params = {k: v.clone().double().requires_grad_(v.is_floating_point() and not k.endswith(('running_mean', 'running_var')))
              if v.is_floating_point() else v.clone() for k, v in sd.items()}
    model_oracle.TRAIN, model_oracle.DTYPE = True, torch.float64
    try:
        ref = model_oracle.rsgnet_forward(params, cfg, x.double())
        ref_loss = sum((o * r.double()).sum() for o, r in zip(ref, R))
        ref_loss.backward()
    finally:
        model_oracle.TRAIN, model_oracle.DTYPE = False, torch.float32
    for o, r in zip(outs, ref):
        assert float((o.detach().cpu().double() - r.detach()).abs().max()) <= 5e-3 * float(r.abs().max())
    rel = []
    for k, p in net.named_parameters():
        if not p.requires_grad:
            continue
        g = params[k].grad
        assert p.grad is not None and g is not None, k
        n = float(g.norm())
        if n > 1e-6 * max(float(q.grad.norm()) for q in params.values() if torch.is_tensor(q) and q.requires_grad and q.grad is not None):
            rel.append(abs(float(p.grad.detach().cpu().double().norm()) - n) / n)
    print(f'sub_sample: per-tensor gradient-norm error median {np.median(rel):.2e} max {max(rel):.2e} (TF32 products)')
    assert np.median(rel) < 0.02 and max(rel) < 0.3
