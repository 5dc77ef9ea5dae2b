"""GPU parity of the model path (through the C ABI plan) against the CPU oracle and the
reference-executed golden fixtures.

Tolerance (stated per BASELINE.json): activations are stored in bf16 (8 bits of mantissa) with fp32
accumulation, so each stage is compared with the fp32 oracle as  max|got-ref| <= TOL * max|ref|.
TOL = 0.03 for the taps and outputs of the shallow test models, 0.05 for the outputs of the
full-depth W32/W48 networks (~100 bf16 layers deep; measured 0.006-0.035, printed with -s)."""
import os

import numpy as np
import pytest
import torch

from oracle import model_oracle
from rsgnet_b200 import _engine
from tests.gpu_util import build, crops, rel_err

pytestmark = pytest.mark.gpu
TOL = 0.03


def _limbs_ok(got, ref, logits=None, tol=0.05):
    """limbs_scores = sigmoid(up2(logits)) (pose_rsgnet.py:1005-1013).  The synthetic weights keep the logits O(1)
    (rsgnet_b200/models/_params.py:_norm_gain), so the map is NOT a saturated 0/1 mask and is compared like every other
    output: max|got - ref| <= tol * max|ref| -- and, stricter, on the PRE-sigmoid values against the fp32 oracle
    logits l:  max|logit(got) - l| <= tol * max|l|."""
    got, ref = np.asarray(got, np.float64), np.asarray(ref, np.float64)
    e = float(np.abs(got - ref).max() / np.abs(ref).max())
    sat = float(((ref < 1e-3) | (ref > 1 - 1e-3)).mean())
    print(f'limbs_scores rel err {e:.5f}, saturated fraction {sat:.4f}')
    ok = e <= tol and sat < 0.2
    if logits is not None:
        l = np.asarray(logits, np.float64)
        g = np.clip(got, 1e-7, 1 - 1e-7)
        el = float(np.abs(np.log(g / (1 - g)) - l).max() / np.abs(l).max())
        print(f'limbs logits rel err {el:.5f} (max|l| = {np.abs(l).max():.3f})')
        ok = ok and el <= tol and np.abs(l).max() < 12.0
    return ok


def _heatmap_report(got, ref, tag):
    """What a heat-map error means for the decode: error normalised by each map's own dynamic range, arg-max agreement
    and key-point displacement (in heat-map pixels) of the bf16 pipeline against the fp32 reference maps."""
    n, k, h, w = ref.shape
    g, r = got.reshape(n * k, -1).astype(np.float64), ref.reshape(n * k, -1).astype(np.float64)
    rng = r.max(1) - r.min(1)
    per_map = np.abs(g - r).max(1) / np.maximum(rng, 1e-12)
    ig, ir = g.argmax(1), r.argmax(1)
    disp = np.hypot(ig % w - ir % w, ig // w - ir // w)
    # a differing arg-max must be a near-tie of the REFERENCE map: its two candidates closer than twice the map's error
    rows = np.arange(n * k)
    err_map = np.abs(g - r).max(1)
    unexplained = int(((ig != ir) & (r[rows, ir] - r[rows, ig] > 2.0 * err_map)).sum())
    rep = dict(per_map_norm_err_max=float(per_map.max()), per_map_norm_err_mean=float(per_map.mean()),
               argmax_agree=float((ig == ir).mean()), disp_le1=float((disp <= 1.0).mean()),
               disp_le2=float((disp <= 2.0).mean()), disp_max=float(disp.max()), unexplained_argmax_flips=unexplained)
    print(tag, 'heat-map decode report', {a: round(b, 4) for a, b in rep.items()})
    return rep


def _nchw(t):
    return t.float().permute(0, 3, 1, 2).contiguous().cpu().numpy()


@pytest.mark.parametrize('key,seed', [('tiny', 0), ('tiny_cp_sub', 1), ('tiny_hrnet', 2)])
def test_tiny_models_per_stage(golden_dir, key, seed):
    g = np.load(os.path.join(golden_dir, f'model_{key}.npz'))
    cfg, net, sd = build(key, seed)
    B = int(g['batch'])
    x = crops(cfg, B, seed + 11)
    stages = {}
    ref = model_oracle.forward(sd, cfg, x, stages=stages)
    # debug engine: every buffer keeps its own memory so that taps survive the run
    eng = _engine.Engine(net, 'cuda', chunk=4, reuse=False)
    shapes = eng.out_shapes(B)
    heat = torch.empty(shapes[_engine.EXT_HEAT], device='cuda')
    aux = {s: torch.empty(shapes[s], device='cuda') for s in shapes if s != _engine.EXT_HEAT}
    eng.run(x.cuda(), heat, B, B, aux=aux or None)
    torch.cuda.synchronize()
    report = {}
    for name, view in eng.info['taps'].items():
        if name not in stages:
            continue
        got = _nchw(eng.pb.tensor_of(view)[:B])
        report[name] = rel_err(got, stages[name].numpy())
    if cfg.MODEL.NAME == 'pose_hrnet':
        report['heatmaps'] = rel_err(heat.cpu().numpy(), ref.numpy())
        gold = {'heatmaps': g['out.heatmaps']}
        outs = {'heatmaps': heat}
    else:
        names = ('multi_kpt_scores', 'kpt_scores', 'limbs_scores', 'relation_scores')
        outs = dict(zip(names, (aux[_engine.EXT_MULTI], heat, aux[_engine.EXT_LIMBS], aux[_engine.EXT_REL])))
        for nm, r in zip(names, ref):
            report[nm] = rel_err(outs[nm].cpu().numpy(), r.numpy())
        gold = {nm: g['out.' + nm] for nm in names}
    print(key, {k: f'{v:.4f}' for k, v in report.items()})
    for name, e in report.items():
        if name == 'limbs_scores':
            assert _limbs_ok(outs[name].cpu().numpy(), ref[2].numpy(), stages['limbs_logits'].numpy(), TOL)
        else:
            assert e <= TOL, (name, e)
    for nm, ref_np in gold.items():      # the unmodified reference's outputs
        if nm == 'limbs_scores':
            assert _limbs_ok(outs[nm].cpu().numpy(), ref_np)
        else:
            assert rel_err(outs[nm].cpu().numpy(), ref_np) <= TOL, nm
    # module API: same numbers through nn.Module.forward, and chunking / graph replay change nothing
    out = net(x.cuda())
    main = out if cfg.MODEL.NAME == 'pose_hrnet' else out[1]
    assert torch.equal(main, heat)
    eng2 = _engine.Engine(net, 'cuda', chunk=1)
    h2 = torch.empty_like(heat)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3):                 # eager, capture, replay
            h2.zero_()
            eng2.run(x.cuda(), h2, B, B, use_graph=True)
    s.synchronize()
    assert torch.equal(h2, heat)


TAP_OF = {'layer1': 'layer1', 'stage2.0': 'stage2.0', 'stage2.1': 'stage2.1', 'stage3.0': 'stage3.0', 'stage3.1': 'stage3.1',
          'stage3.2': 'stage3.2', 'stage4.0': 'stage4.0', 'vis_conv': 'vis', 'type_conv': 'type',
          'predict_contact_net': 'final_vis', 'kpt_net': 'kpt_net', 'predict_net': 'kpt_feat'}
FULL_TOL = 0.05


@pytest.mark.parametrize('key,seed', [('w32_coco', 3), ('w32_crowdpose', 4), ('hrnet_w32_coco', 5),
                                      ('w48_coco_384', 6)])
def test_full_models_vs_reference_golden(golden_dir, key, seed):
    """Full-depth networks under the PRODUCTION kernel routing (no environment switches anywhere in the test suite)
    against the outputs AND the per-stage taps of the unmodified reference (oracle/gen_golden.py: forward hooks on the
    reference's top-level modules): max-abs error per stage relative to that stage's max|ref|."""
    g = np.load(os.path.join(golden_dir, f'model_{key}.npz'))
    cfg, net, sd = build(key, seed)
    B = int(g['batch'])
    x = crops(cfg, B, seed + 11).cuda()
    sub = int(g['sub'])
    out = net(x)
    if cfg.MODEL.NAME == 'pose_hrnet':
        outs = {'heatmaps': out}
    else:
        outs = dict(zip(('multi_kpt_scores', 'kpt_scores', 'limbs_scores', 'relation_scores'), out))
    # the same plan with every buffer kept alive (reuse=False changes memory placement only, not the routing)
    eng = _engine.Engine(net, 'cuda', chunk=B, reuse=False)
    heat = torch.empty(eng.out_shapes(B)[_engine.EXT_HEAT], device='cuda')
    eng.run(x, heat, B, B)
    torch.cuda.synchronize()
    main = out if cfg.MODEL.NAME == 'pose_hrnet' else out[1]
    assert torch.equal(heat, main)
    stage_rep = {}
    for ref_name, tap in TAP_OF.items():
        if 'tap.' + ref_name not in g.files or tap not in eng.info['taps']:
            continue
        got = _nchw(eng.pb.tensor_of(eng.info['taps'][tap])[:B]).reshape(-1)
        ref = g['tap.' + ref_name]
        got = got[::sub] if got.size > 65536 else got
        assert got.shape == ref.shape, (ref_name, got.shape, ref.shape)
        stage_rep[ref_name] = float(np.abs(got - ref).max() / float(g['tapabsmax.' + ref_name]))
    print(key, 'per-stage max-abs error / max|ref|:', {k: f'{v:.4f}' for k, v in stage_rep.items()})
    assert len(stage_rep) >= (7 if cfg.MODEL.NAME == 'pose_hrnet' else 11)
    for nm, e in stage_rep.items():
        assert e <= FULL_TOL, (nm, e)
    rep = {}
    logits = None
    if 'limbs_scores' in outs:      # fp32 oracle logits for the pre-sigmoid limbs bar (see _limbs_ok)
        st = {}
        model_oracle.forward(sd, cfg, x.cpu(), stages=st)
        logits = st['limbs_logits'].reshape(-1).numpy()
    for nm, t in outs.items():
        if 'out.' + nm in g:
            ref, got = g['out.' + nm], t.cpu().numpy()
        else:
            ref = g['sub.' + nm]
            flat = t.reshape(-1).cpu().numpy()
            got = flat[::sub] if flat.size > 65536 else flat
        rep[nm] = float(np.abs(got - ref).max() / float(g['absmax.' + nm]))
        if nm == 'limbs_scores':
            lg = logits.reshape(got.shape) if got.size == logits.size else logits[::sub]
            assert _limbs_ok(got, ref, lg, FULL_TOL)
        if nm == 'relation_scores':           # not a saturated mask: the TRP affinity is genuinely exercised
            assert float(((ref > 1e-3) & (ref < 1 - 1e-3)).mean()) > 0.5
    print(key, {k: f'{v:.4f}' for k, v in rep.items()})
    for nm, e in rep.items():
        assert e <= FULL_TOL, (nm, e)
    hm_name = 'heatmaps' if cfg.MODEL.NAME == 'pose_hrnet' else 'kpt_scores'
    dec = _heatmap_report(outs[hm_name].cpu().numpy(), g['out.' + hm_name], key)
    assert dec['per_map_norm_err_max'] <= 0.06 and dec['unexplained_argmax_flips'] == 0, dec


def test_flip_batch_equals_two_forwards():
    """[x ; flip(x)] in one run (stem reads the second half W-reversed) == two separate forwards."""
    cfg, net, sd = build('tiny', 0)
    x = crops(cfg, 3, 5).cuda()
    eng = _engine.engine_for(net, 'cuda')
    shapes = eng.out_shapes(6)
    heat = torch.empty(shapes[_engine.EXT_HEAT], device='cuda')
    eng.run(x, heat, 6, 3)
    a = net(x)[1]
    b = net(x.flip(3))[1]
    assert torch.equal(heat[:3], a) and torch.equal(heat[3:], b)


def test_pipeline_matches_reference_loop_on_oracle_heatmaps():
    """infer_crops == flip-average + decode of this implementation's own heat-maps by the oracle
    (decode is bit-exact given identical heat-maps)."""
    from oracle import decode_oracle
    from rsgnet_b200 import presets, synth
    from rsgnet_b200.pipeline import CropPipeline
    cfg, net, sd = build('tiny', 0)
    B = 5
    x = crops(cfg, B, 9)
    c, s = synth.centers_scales(B, seed=1)
    pipe = CropPipeline(net, cfg, B, use_graph=False)
    preds, maxvals = pipe(x, c, s)
    hm = net(x.cuda())[1].cpu().numpy()
    hf = net(x.cuda().flip(3))[1].cpu().numpy()
    avg = decode_oracle.flip_average(hm, hf, presets.flip_pairs_for(17), shift=True)
    o_preds, o_mv = decode_oracle.get_final_preds(True, avg, c, s)
    assert np.array_equal(preds, o_preds) and np.array_equal(maxvals, o_mv)
    assert pipe.launches_per_step() > 100


def test_overlapped_pipeline_equals_sequential():
    """run_overlapped (double-buffered H2D under compute, CUDA-graph replay) == step-by-step calls."""
    from rsgnet_b200 import synth
    from rsgnet_b200.pipeline import CropPipeline
    cfg, net, sd = build('tiny', 0)
    B = 4
    batches = []
    for i in range(5):
        c, s = synth.centers_scales(B, seed=20 + i)
        batches.append((crops(cfg, B, 30 + i).pin_memory(), torch.from_numpy(c).pin_memory(), torch.from_numpy(s).pin_memory()))
    seq = CropPipeline(net, cfg, B, use_graph=False)
    ref = [seq(x, c, s) for x, c, s in batches]
    pipe = CropPipeline(net, cfg, B, use_graph=True)
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        got = pipe.run_overlapped(batches)
    st.synchronize()
    for (p, m), (rp, rm) in zip(got, ref):
        assert np.array_equal(p.cpu().numpy(), rp) and np.array_equal(m.cpu().numpy(), rm)


def test_module_forward_graph_replay_batch_sizes_and_lazy_aux():
    """nn.Module.forward as the reference's loop calls it (function.py:389-402): fresh output tensors per call (the
    loop holds `output` across the flipped forward), CUDA-graph replay from the third call on, CPU inputs (what
    DataParallel's caller passes), batch sizes on both sides of a chunk boundary, and the lazily materialised outputs."""
    cfg, net, sd = build('tiny', 0)
    x1, x2 = crops(cfg, 3, 5).cuda(), crops(cfg, 3, 6).cuda()
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        outs = [net(x1), net(x2), net(x1), net(x2), net(x1)]          # eager, capture, replay, replay, replay
    st.synchronize()
    for k in range(4):
        assert torch.equal(outs[0][k], outs[2][k]) and torch.equal(outs[0][k], outs[4][k]) and torch.equal(outs[1][k], outs[3][k])
        assert outs[0][k].data_ptr() != outs[2][k].data_ptr()
    assert not torch.equal(outs[0][1], outs[1][1])
    eng = _engine.engine_for(net, 'cuda', 32)
    ref = torch.empty(eng.out_shapes(3)[_engine.EXT_HEAT], device='cuda')
    eng.run(x1, ref, 3, 3)
    assert torch.equal(ref, outs[0][1])
    assert torch.equal(net(x1.cpu())[1], ref)                        # host input: copied into the static buffer
    big = torch.cat([x1, crops(cfg, 34, 7).cuda()])                  # 37 crops: the 64-forward engine
    assert torch.equal(net(big)[1][:3], ref)
    net.lazy_aux = True
    try:
        lz = net(x1)
        assert isinstance(lz[0], _engine.LazyOutput) and tuple(lz[0].shape) == tuple(outs[0][0].shape)
        assert tuple(lz[3].shape) == tuple(outs[0][3].shape) and lz[2].device == ref.device
        assert torch.equal(lz[1], ref)
        assert torch.equal(lz[2].materialize(), outs[0][2])          # second run with the auxiliary ops, same numbers
        assert torch.equal(torch.sigmoid(lz[0]), torch.sigmoid(outs[0][0]))      # torch functions accept it
        assert torch.equal(lz[3][1], outs[0][3][1]) and torch.equal(lz[0].cpu(), outs[0][0].cpu())
    finally:
        net.lazy_aux = False


def test_two_streams_share_an_engine_safely():
    """Two pipelines on two CUDA streams share the cached engine (arena, concat buffers): runs are ordered by an event,
    so interleaving them cannot corrupt either result."""
    from rsgnet_b200 import synth
    from rsgnet_b200.pipeline import CropPipeline
    cfg, net, sd = build('tiny', 0)
    B = 4
    xa, xb = crops(cfg, B, 41), crops(cfg, B, 42)
    ca, sa = synth.centers_scales(B, seed=3)
    seq = CropPipeline(net, cfg, B, use_graph=False)
    ra, rb = seq(xa, ca, sa), seq(xb, ca, sa)
    pa, pb_ = CropPipeline(net, cfg, B, use_graph=False), CropPipeline(net, cfg, B, use_graph=False)
    assert pa.engine is pb_.engine
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    for p, x in ((pa, xa), (pb_, xb)):
        p.x.copy_(x.cuda()); p.center.copy_(torch.from_numpy(ca).cuda()); p.scale.copy_(torch.from_numpy(sa).cuda())
    torch.cuda.synchronize()
    got = []
    for _ in range(6):
        with torch.cuda.stream(s1):
            o1 = pa.run_device()
        with torch.cuda.stream(s2):
            o2 = pb_.run_device()
        got.append((o1, o2))
    torch.cuda.synchronize()
    for o1, o2 in got[-2:]:
        assert np.array_equal(o1[0].cpu().numpy(), ra[0]) and np.array_equal(o2[0].cpu().numpy(), rb[0])


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two CUDA devices')
def test_dataparallel_two_devices():
    """tools/cp_test.py:99 wraps the model in torch.nn.DataParallel unconditionally: replicas (empty _parameters,
    per-forward broadcast copies) take their engine from the source module's per-device cache."""
    cfg, net, sd = build('tiny', 0)
    x = crops(cfg, 6, 5)
    ref = [t.cpu() for t in net(x.cuda())]
    dp = torch.nn.DataParallel(net, device_ids=[0, 1]).cuda()
    dp.eval()
    for _ in range(3):
        out = dp(x)                                              # CPU input, as function.py:389 passes it
        for a, b in zip(out, ref):
            assert a.device.index == 0 and torch.equal(a.cpu(), b)
    assert sorted(k[0] for k in net.__dict__['_rsg_engines']) == ['cuda:0', 'cuda:1']
    hr = __import__('tests.gpu_util', fromlist=['build']).build('tiny_hrnet', 2)[1]
    dph = torch.nn.DataParallel(hr, device_ids=[0, 1]).cuda().eval()
    xh = crops(cfg, 5, 9)
    assert torch.equal(dph(xh).cpu(), hr(xh.cuda()).cpu())
