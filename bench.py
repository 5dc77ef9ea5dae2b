#!/usr/bin/env python
"""Benchmark of the RSGNet per-crop inference hot path (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (config.workload): BASELINE.json configs[1] -- RSGNet HRNet-W32 256x192 CrowdPose (K=14)
flip-test inference, 256 synthetic crops per GPU per step, random-init ("trained-like") weights.
A step = one pass of the hot path over one batch: 2 forwards per crop (crop + W-flipped crop),
flip-average, argmax / quarter-offset decode, inverse affine -> 12 bytes per joint.
  value : crops/s with the batch already resident in HBM (device-timed, CUDA events)
  e2e   : crops/s through the public API (rsgnet_b200.pipeline.CropPipeline.__call__) with pinned
          HOST buffers: H2D of the crops + D2H of preds/maxvals inside the timed region
  roofline : the dominant (kernel, shape) group of the step (by device time, from rsg_plan_profile's
          per-op CUDA-event timings taken live in this process) against the bound that binds it --
          algorithmic bytes vs measured HBM copy bandwidth, or executed FLOPs vs measured bf16 peak
          (MEASURED_PEAKS.json); both floors are reported
  cpu_baseline : the CPU oracle (oracle/model_oracle.py + oracle/decode_oracle.py, a port of the
          reference's torch/NumPy path) on this box's host cores, bounded sample
`--impl reference` times that CPU path as its own arm (the reference is Python importing from
/root/reference, which does not exist on the GPU box, so the port is what runs).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from rsgnet_b200 import presets, synth  # noqa: E402
from rsgnet_b200.models import _params, pose_rsgnet  # noqa: E402

PRESET = 'w32_crowdpose'
WORKLOAD = 'RSGNet-W32 256x192 CrowdPose K=14 flip-test inference (BASELINE.json configs[1])'
METRIC = 'crops/sec RSGNet-W32 256x192 flip-test inference'
# --workload: preset -> (workload description, metric)
WORKLOADS = {
    'postproc': ('post-processing only: get_final_preds decode + evaluate (rescoring, grouping, oks_nms) on 100k synthetic '
                 'detections x 17 heat-maps at 64x48 (BASELINE.json configs[3])', 'crops/sec get_final_preds decode 17x64x48'),
    'train': ('RSGNet-W32 256x192 COCO K=17 training step (forward + losses + backward + Adam), batch 32/GPU, data-parallel '
              'with one NCCL all-reduce of the flat gradient buffer (BASELINE.json configs[4])',
              'samples/sec RSGNet-W32 256x192 training step'),
    'w32_crowdpose': (WORKLOAD, METRIC),
    'w48_coco_384': ('RSGNet-W48 384x288 COCO K=17 flip-test inference (BASELINE.json configs[2])',
                     'crops/sec RSGNet-W48 384x288 flip-test inference'),
}
REF_GFLOP_PER_CROP = 37.762          # reference op graph, 2 forwards


def peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.isfile(path):
        p = json.load(open(path))
        return dict(hbm=p['hbm_gbs'], tf_burst=p['bf16_tflops'], tf_sus=p['bf16_tflops_sustained'],
                    src='MEASURED_PEAKS.json (measured)')
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sus=1400.0, src='B200_PROFILING.md fallback')


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index=0):
        self.rows, self.stop_flag, self.index = [], False, index

    def _loop(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(['nvidia-smi', f'--id={self.index}', f'--query-gpu={self.Q}',
                                      '--format=csv,noheader,nounits'], capture_output=True,
                                     text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(',')])
            except Exception:
                pass
            time.sleep(0.02)

    def start(self):
        self.th = threading.Thread(target=self._loop, daemon=True)
        self.th.start()
        return self

    def stop(self):
        self.stop_flag = True
        self.th.join(timeout=6)
        sm = [float(r[0]) for r in self.rows if r and r[0].replace('.', '').isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace('.', '').isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4)
                          if r[3 + i].lower().startswith('active')})
        return dict(sm_mhz=float(np.median(sm)) if sm else None,
                    sm_max_mhz=max(mx) if mx else None, reasons=reasons, samples=len(sm))


def build_model(device):
    cfg = presets.preset(PRESET)
    net = pose_rsgnet.get_pose_net(cfg, False)
    sd = _params.synth_state_dict(net, seed=4)
    net.load_state_dict(sd)
    return cfg, net.to(device).eval(), sd


def cpu_reference_step(sd, cfg, x, c, s, want_maps=False):
    """One bounded sample of the reference's CPU path (oracle port): 2 forwards, flip-average,
    get_final_preds."""
    from oracle import decode_oracle, model_oracle
    k = int(cfg.MODEL.NUM_JOINTS)
    a = model_oracle.forward(sd, cfg, x)[1].numpy()
    b = model_oracle.forward(sd, cfg, x.flip(3))[1].numpy()
    avg = decode_oracle.flip_average(a, b, presets.flip_pairs_for(k), shift=True)
    if want_maps:
        return decode_oracle.get_final_preds(True, avg, c, s) + (avg,)
    return decode_oracle.get_final_preds(True, avg, c, s)


def gpu_eager_baseline(sd, cfg, x_host, dev, n=64):
    """The reference graph (oracle/model_oracle.py, plain torch ops) on the GPU: what `tools/cp_test.py` would run on
    this box with stock PyTorch.  Bench leg only."""
    from oracle import model_oracle
    sd_dev = {k: v.to(dev) for k, v in sd.items()}
    x = x_host[:n].to(dev)
    res = {}
    for tag, ctx in (('fp32', torch.autocast('cuda', enabled=False)), ('bf16_autocast', torch.autocast('cuda', dtype=torch.bfloat16))):
        try:
            with torch.no_grad(), ctx:
                for _ in range(2):
                    model_oracle.forward(sd_dev, cfg, x)
                    model_oracle.forward(sd_dev, cfg, x.flip(3))
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                reps = 3
                for _ in range(reps):
                    model_oracle.forward(sd_dev, cfg, x)
                    model_oracle.forward(sd_dev, cfg, x.flip(3))
                e1.record()
                torch.cuda.synchronize()
            res[tag] = {'crops_per_s': n * reps / (e0.elapsed_time(e1) * 1e-3), 'batch': n}
        except Exception as e:            # informational leg: never fails the bench
            res[tag] = {'error': str(e)[:200]}
    res['what'] = 'reference op graph, eager PyTorch (cuDNN/cuBLAS) on this GPU, 2 forwards per crop, no decode'
    return res


def dropin_loop_throughput(net, cfg, x_host, c_np, s_np, dev, batch=32, iters=6):
    """The reference loop's own call sequence (lib/core/function.py:389-450) at its TEST.BATCH_SIZE_PER_GPU = 32, with the
    drop-in modules standing where the shim puts them: model(input) -> model(input.flip(3)) -> flip_back(.cpu().numpy()) ->
    back to the device -> 1-px shift -> average -> get_final_preds(config, output.clone().cpu().numpy(), c, s).
    Three D2H copies of full heat-maps per batch, exactly as the unchanged caller does them."""
    from rsgnet_b200.core.inference import get_final_preds
    from rsgnet_b200.utils.transforms import flip_back
    pairs = presets.flip_pairs_for(int(cfg.MODEL.NUM_JOINTS))
    x = x_host[:batch].clone()
    c, s = c_np[:batch], s_np[:batch]
    was_lazy = getattr(net, 'lazy_aux', False)
    net.lazy_aux = True                                 # the loop reads outputs[1] only
    st = torch.cuda.Stream(dev)

    def one_batch():
        outputs = net(x)                                # CPU tensor in, as DataParallel's caller passes it
        output = outputs[1]
        input_flipped = x.flip(3)
        output_flipped = net(input_flipped)[1]
        output_flipped = flip_back(output_flipped.cpu().numpy(), pairs)
        output_flipped = torch.from_numpy(output_flipped.copy()).to(dev)
        output_flipped[:, :, :, 1:] = output_flipped.clone()[:, :, :, 0:-1]
        output = (output + output_flipped) * 0.5
        return get_final_preds(cfg, output.clone().cpu().numpy(), c, s)
    try:
        with torch.cuda.stream(st):
            for _ in range(3):
                one_batch()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(iters):
                one_batch()
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / iters
    finally:
        net.lazy_aux = was_lazy
    return {'crops_per_s': batch / dt, 'batch': batch, 'ms_per_batch': dt * 1e3,
            'what': 'lib/core/function.py:389-450 call sequence through the drop-in modules (2 module forwards with graph replay, '
                    'host flip_back, 3 D2H copies of full heat-maps, get_final_preds on NumPy), wall clock'}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path (oracle port), all host
    threads, on our arm's workload, a bounded sample (--cpu-sample crops) per step."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    if PRESET == 'postproc':
        return run_postproc_reference(args)
    if PRESET == 'train':
        return run_train_reference(args)
    cores = len(os.sched_getaffinity(0))
    torch.set_num_threads(cores)
    cfg = presets.preset(PRESET)
    net = pose_rsgnet.get_pose_net(cfg, False)
    sd = _params.synth_state_dict(net, seed=4)
    n = args.cpu_sample
    x = torch.from_numpy(synth.crops(n, cfg.MODEL.IMAGE_SIZE, seed=0))
    c, s = synth.centers_scales(n, seed=0)
    for _ in range(args.warmup):
        cpu_reference_step(sd, cfg, x, c, s)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_reference_step(sd, cfg, x, c, s)
    dt = time.perf_counter() - t0
    val = n * args.steps / dt
    sample = f'{n} crops/step (flip-test: {2 * n} fp32 forwards + flip-average + decode), torch CPU ops'
    print(json.dumps({
        'impl': 'reference', 'metric': METRIC, 'value': val, 'unit': 'crops/s', 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': dt / args.steps * 1e3,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
        'data': 'synthetic', 'config': {'workload': WORKLOAD, 'crops_per_step': n, 'forwards_per_crop': 2,
                                        'sample': 'bounded: %d of the 256 crops per GPU per step of the workload' % n},
        'cpu_baseline': {'value': val, 'unit': 'crops/s', 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': val, 'unit': 'crops/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0}))


# dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed
# `ncu --set full` captures (profiles/r1_conv_bb_ncu_details.txt, r1_conv_tc5_ncu_details.txt,
# r1_conv_bneck_ncu_details.txt; 512 forwards)
TRAIN_NCU_TRAFFIC = 12.65e6       # conv3x3_tf32_small_kernel, 32->32 @64x48 x 32 samples (algorithmic: 25.2 MB)
NCU_TRAFFIC = {('conv_tcgen05', 'conv 32->32 taps9 s1 64x48 +res'): 201.44e6 + 72.35e6,
               ('basic_block_tcgen05', 'bblock'): 100.78e6 + 54.63e6,
               ('bottleneck_tcgen05', 'bneck'): 902.87e6 + 759.45e6}

FAMILY = {0: 'stem', 1: 'conv_mma', 2: 'conv_tcgen05', 3: 'fuse', 4: 'maxpool', 5: 'trp_attention',
          6: 'relation_scores', 7: 'groupnorm', 8: 'bilinear', 9: 'conv_ws_tcgen05', 10: 'basic_block_tcgen05',
          11: 'bottleneck_tcgen05', 12: 'conv_ws2_tcgen05'}


# ---------------------------------------------------------------------------------------------
# BASELINE.json configs[3]: post-processing only
# ---------------------------------------------------------------------------------------------
POST_K, POST_H, POST_W = 17, 64, 48


def _post_heatmaps(n, device):
    """SURVEY.md §8d config 4 on the device (20.9 GB for n = 100 k): Gaussian bump (sigma 2, centre anywhere incl. the
    borders, amplitude U(0.1, 1)) + N(0, 0.01) noise, 15 % of the maps shifted by -2 so that their max <= 0."""
    g = torch.Generator(device=device).manual_seed(2)
    hm = torch.empty((n, POST_K, POST_H, POST_W), device=device)
    ys = torch.arange(POST_H, device=device).view(1, 1, POST_H, 1).float()
    xs = torch.arange(POST_W, device=device).view(1, 1, 1, POST_W).float()
    for lo in range(0, n, 4000):
        m = min(4000, n - lo)
        r = lambda: torch.rand((m, POST_K, 1, 1), device=device, generator=g)
        cx, cy, amp = r() * (POST_W + 1) - 1, r() * (POST_H + 1) - 1, r() * 0.9 + 0.1
        t = amp * torch.exp(-((xs - cx) ** 2 + (ys - cy) ** 2) / 8.0)
        t += 0.01 * torch.randn((m, POST_K, POST_H, POST_W), device=device, generator=g)
        hm[lo:lo + m] = torch.where(r() < 0.15, t - 2.0, t)
    return hm


def _post_cpu(n_crops, n_imgs, hm_np, c_np, s_np, ev):
    """The reference's CPU post-processing (oracle port) on a bounded slice: get_final_preds on n_crops maps and
    evaluate() (rescoring + grouping + oks_nms) on n_imgs images.  Returns (crops/s, dets/s, results)."""
    from oracle import decode_oracle, nms_oracle
    t0 = time.perf_counter()
    preds, mv = decode_oracle.get_final_preds(True, hm_np[:n_crops], c_np[:n_crops], s_np[:n_crops])
    t_dec = time.perf_counter() - t0
    p, b, ids = ev
    sel = ids < (100000 + 7 * n_imgs)
    t0 = time.perf_counter()
    res = nms_oracle.evaluate(p[sel], b[sel], ids[sel], None, 0.2, 0.9)
    t_nms = time.perf_counter() - t0
    return n_crops / t_dec, int(sel.sum()) / t_nms, (preds, mv, res, np.nonzero(sel)[0])


def run_postproc_reference(args):
    cores = len(os.sched_getaffinity(0))
    n_crops, n_imgs = 2048, 500
    hm = _post_heatmaps_host(n_crops)
    c, sc = synth.centers_scales(n_crops, seed=5)
    ev = synth.evaluate_inputs(5000, 20, POST_K, seed=31, ragged=False)
    vals = []
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        cps, dps, _ = _post_cpu(n_crops, n_imgs, hm, c, sc, ev)
        if i >= args.warmup:
            vals.append((cps, dps, time.perf_counter() - t0))
    val = float(np.mean([v[0] for v in vals]))
    sample = f'{n_crops} of 100000 crops decoded + {n_imgs} of 5000 images through evaluate() per step, NumPy (single thread)'
    print(json.dumps({
        'impl': 'reference', 'metric': METRIC, 'value': val, 'unit': 'crops/s', 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': float(np.mean([v[2] for v in vals])) * 1e3, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic', 'config': {'workload': WORKLOAD, 'sample': sample},
        'nms_dets_per_s': float(np.mean([v[1] for v in vals])),
        'cpu_baseline': {'value': val, 'unit': 'crops/s', 'cores': 1, 'kind': 'port', 'sample': sample, 'host_cores': cores},
        'e2e': {'value': val, 'unit': 'crops/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}, 'gpu_launches': 0}))


def _post_heatmaps_host(n):
    return synth.heatmaps(n, POST_K, POST_H, POST_W, seed=2)


def run_postproc(args):
    """`--workload postproc`: decode of N x 17 x 64x48 heat-maps against the HBM roof + evaluate_device dets/s."""
    from rsgnet_b200.core.inference import decode_device, get_final_preds
    from rsgnet_b200.nms.nms import evaluate_device
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group('nccl', device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
    N = args.batch if args.batch != 256 else 100000          # detections per GPU per step
    pk = peaks()
    hm = _post_heatmaps(N, dev)
    c_np, s_np = synth.centers_scales(N, seed=5 + rank)
    c, sc = torch.from_numpy(c_np).to(dev), torch.from_numpy(s_np).to(dev)
    n_imgs = N // 20
    ev_np = synth.evaluate_inputs(n_imgs, 20, POST_K, seed=31 + rank, ragged=False)
    ev = (torch.from_numpy(ev_np[0]).to(dev), torch.from_numpy(ev_np[1]).to(dev), torch.from_numpy(ev_np[2]).to(dev))
    stream = torch.cuda.Stream(dev)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, sample=False):
        barrier()
        sampler = ClockSampler(local).start() if sample else None
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record(stream)
            for _ in range(steps):
                fn()
            e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        clocks = sampler.stop() if sampler else None
        if dist is not None:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, clocks

    out = {}

    def step():                      # the hot path of configs[3]: decode + evaluate, device-resident
        out['d'] = decode_device(hm, c, sc, post_process=True)
        out['e'] = evaluate_device(ev[0], ev[1], ev[2], 0.9, 0.2)

    def step_decode():
        out['d'] = decode_device(hm, c, sc, post_process=True)

    def step_eval():
        out['e'] = evaluate_device(ev[0], ev[1], ev[2], 0.9, 0.2)

    with torch.cuda.stream(stream):
        for _ in range(args.warmup):
            step()
    ms_all, clocks = timed(step, args.steps, sample=(rank == 0))
    ms_dec, _ = timed(step_decode, args.steps)
    ms_ev, _ = timed(step_eval, args.steps)
    # e2e: the reference-facing call with HOST buffers (get_final_preds(config, ndarray, center, scale)) on a slice
    n_e2e = min(N, 8192)
    cfg = presets.make_cfg(post_process=True)
    hm_host = hm[:n_e2e].cpu().numpy()
    get_final_preds(cfg, hm_host, c_np[:n_e2e], s_np[:n_e2e])
    barrier()
    t0 = time.perf_counter()
    reps = 3
    for _ in range(reps):
        p_host, mv_host = get_final_preds(cfg, hm_host, c_np[:n_e2e], s_np[:n_e2e])
    torch.cuda.synchronize()
    t_e2e = (time.perf_counter() - t0) / reps
    if dist is not None:
        t = torch.tensor([t_e2e], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_e2e = float(t.item())
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    byt = N * POST_K * POST_H * POST_W * 4 + N * POST_K * 12
    per_ms = ms_dec / args.steps
    gbs = byt / per_ms / 1e6
    # parity of what was timed + CPU leg (bounded slice)
    cpu = None
    parity = None
    if not args.no_cpu_baseline:
        n_c, n_i = 2048, 500
        cps, dps, (o_preds, o_mv, o_ev, sel) = _post_cpu(n_c, n_i, hm[:n_c].cpu().numpy(), c_np, s_np, ev_np)
        d = out['d']
        dec_ok = bool(np.array_equal(d['preds'][:n_c].cpu().numpy(), o_preds) and np.array_equal(d['maxvals'][:n_c].cpu().numpy(), o_mv))
        sub = evaluate_device(ev_np[0][sel], ev_np[1][sel], ev_np[2][sel], 0.9, 0.2).host()
        ev_ok = bool(np.array_equal(sub['images'], o_ev[0]) and np.array_equal(sub['counts'], o_ev[1]) and
                     np.array_equal(sub['keep'], o_ev[2]) and np.array_equal(sub['scores'][sub['keep']], o_ev[3]))
        parity = {'decode_bit_exact_vs_oracle': dec_ok, 'evaluate_keep_lists_equal_oracle': ev_ok,
                  'sample': f'{n_c} crops, {n_i} images'}
        if not (dec_ok and ev_ok):
            raise SystemExit(f'bench: post-processing parity check failed: {parity}')
        cpu = {'value': cps, 'unit': 'crops/s', 'cores': 1, 'kind': 'port', 'nms_dets_per_s': dps,
               'sample': f'get_final_preds on {n_c} of {N} crops, evaluate() on {n_i} of {n_imgs} images (oracle port, NumPy)'}
    line = {
        'metric': METRIC, 'value': N * world * args.steps / (ms_dec * 1e-3), 'unit': 'crops/s', 'n_gpus': world,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': per_ms, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': WORKLOAD, 'detections_per_gpu_per_step': N, 'K': POST_K, 'heatmap': [POST_H, POST_W],
                   'images': n_imgs, 'l2': 'input (20.9 GB at 100k) exceeds the 126 MB L2; no explicit flush'},
        'clocks': clocks, 'gpu_launches': 6 * args.steps,
        'decode_plus_evaluate': {'ms_per_step': ms_all / args.steps, 'detections_per_s': N * world * args.steps / (ms_all * 1e-3)},
        'evaluate_device': {'ms_per_step': ms_ev / args.steps, 'dets_per_s': N * world * args.steps / (ms_ev * 1e-3),
                            'bound': 'latency (22 MB per 100 k detections): no roofline claimed'},
        'e2e': {'value': n_e2e * world / t_e2e, 'unit': 'crops/s', 'h2d_bytes_per_step': int(hm_host.nbytes + 16 * n_e2e),
                'd2h_bytes_per_step': int(p_host.nbytes + mv_host.nbytes), 'sample': f'{n_e2e} crops per call through '
                'core.inference.get_final_preds(config, ndarray, center, scale) with pageable NumPy buffers, as the reference loop calls it'},
        'roofline': {'bound': 'hbm', 'kernel': f'decode_kernel: {N} x {POST_K} maps of {POST_H}x{POST_W} f32', 'achieved': gbs,
                     'peak': pk['hbm'], 'unit': 'GB/s', 'frac': gbs / pk['hbm'], 'traffic': None,
                     'peak_source': pk['src'] + ' hbm_gbs', 'algorithmic_bytes_per_launch': byt, 'us_per_launch': per_ms * 1e3,
                     'note': 'the measured peak is a COPY bandwidth (half reads, half writes); this kernel only reads (12 B written per '
                             '13 KB read), and a read-only HBM3e stream exceeds the copy figure -- a frac above 1 is not an L2 effect '
                             '(the 20.9 GB input is 165x the L2)'},
        'parity_checked': bool(parity), 'parity': parity, 'cpu_baseline': cpu,
    }
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()



# ---------------------------------------------------------------------------------------------
# BASELINE.json configs[4]: training step
# ---------------------------------------------------------------------------------------------
TRAIN_PRESET = 'w32_coco'


def _train_batch(cfg, n, seed):
    return synth.train_batch(n, cfg.MODEL.IMAGE_SIZE, cfg.MODEL.HEATMAP_SIZE, cfg.MODEL.NUM_JOINTS, cfg.MODEL.NUM_LIMBS, seed=seed)


def _train_cpu(cfg, sd, n, reps):
    """The reference's CPU training iteration (oracle/train_oracle.py: torch CPU ops + autograd + Adam) on n samples."""
    from oracle import train_oracle
    b = {k: torch.from_numpy(v) for k, v in _train_batch(cfg, n, 1).items()}
    cur = {k: v.clone() for k, v in sd.items()}
    state = None
    times, losses = [], None
    for i in range(reps + 1):
        t0 = time.perf_counter()
        losses, grads, bufs, _ = train_oracle.forward_backward(cur, cfg, b)
        new, state = train_oracle.adam_step(cur, grads, state=state)
        cur.update(new)
        cur.update(bufs)
        if i:
            times.append(time.perf_counter() - t0)
        if i == 0:
            first = (losses, grads)
    return n / float(np.median(times)), first


def run_train_reference(args):
    cores = len(os.sched_getaffinity(0))
    torch.set_num_threads(cores)
    cfg = presets.preset(TRAIN_PRESET)
    net = pose_rsgnet.get_pose_net(cfg, True)
    sd = _params.synth_state_dict(net, seed=4)
    n = min(args.cpu_sample, 8)
    val, _ = _train_cpu(cfg, sd, n, max(1, min(args.steps, 3)))
    sample = f'{n} of the 32 samples per GPU per step: fp32 forward (batch-statistics BN) + losses + autograd backward + Adam, torch CPU ops'
    print(json.dumps({
        'impl': 'reference', 'metric': METRIC, 'value': val, 'unit': 'samples/s', 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': n / val * 1e3, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic', 'config': {'workload': WORKLOAD, 'samples_per_step': n, 'sample': sample},
        'cpu_baseline': {'value': val, 'unit': 'samples/s', 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': val, 'unit': 'samples/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}, 'gpu_launches': 0}))


def run_train(args):
    """`--workload train`: one optimisation step per bench step on B samples per GPU (weak scaling), gradients averaged
    with one NCCL all-reduce of the flat buffer.  `value`: batch resident in HBM, replayed as a CUDA graph; `e2e`: the
    batch comes from pinned host memory every step and the losses are read back."""
    import collections
    from rsgnet_b200.train import TrainStep
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group('nccl', device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
    B = args.batch if args.batch != 256 else 32
    pk = peaks()
    cfg = presets.preset(TRAIN_PRESET)
    net = pose_rsgnet.get_pose_net(cfg, True)
    sd = _params.synth_state_dict(net, seed=4)
    net.load_state_dict(sd)
    net = net.to(dev).train()
    keys = ('input', 'target', 'target_weight', 'all_ins_target', 'all_ins_target_weight', 'target_limbs')
    host = {k: torch.from_numpy(v).pin_memory() for k, v in _train_batch(cfg, B, 100 + rank).items()}
    batch = [host[k].to(dev) for k in keys]
    ts = TrainStep(net, lr=1e-3)
    stream = torch.cuda.Stream(dev)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # parity of the path that is timed: the first samples through the GPU step and through the CPU oracle (same weights)
    parity, cpu = None, None
    if rank == 0 and not args.no_cpu_baseline:
        n_c = 2
        cps, (o_loss, o_grads) = _train_cpu(cfg, sd, n_c, 1)
        small = {k: torch.from_numpy(v).to(dev) for k, v in _train_batch(cfg, n_c, 1).items()}
        pnet = pose_rsgnet.get_pose_net(cfg, True)
        pnet.load_state_dict(sd)
        pnet = pnet.to(dev).train()
        pts = TrainStep(pnet, lr=1e-3)
        L, _ = pts.forward_backward(*[small[k] for k in keys])
        L = L.read()
        P = dict(pnet.named_parameters())
        rel = []
        for k, g in o_grads.items():
            off, n = pts.store.offsets[id(P[k])]
            ours = pts.store.flat_g[off:off + n].cpu().double()
            ref = g.reshape(-1).double()
            if float(ref.norm()) > 0:
                rel.append(abs(float(ours.norm()) - float(ref.norm())) / float(ref.norm()))
        lerr = max(abs(L[k] - o_loss[k]) / abs(o_loss[k]) for k in ('multi_loss', 'target_loss', 'skeleton_loss', 'relation_loss'))
        parity = {'samples': n_c, 'loss_rel_err_max': lerr, 'grad_norm_rel_err_median': float(np.median(rel)),
                  'bar': 'TF32 products vs the fp32 CPU oracle: every loss term within 5e-3, median per-tensor gradient norm within 3 %'}
        if not (lerr < 5e-3 and np.median(rel) < 0.03):
            raise SystemExit(f'bench: training parity check failed: {parity}')
        cpu = {'value': cps, 'unit': 'samples/s', 'cores': torch.get_num_threads(), 'kind': 'port',
               'sample': f'{n_c} of the {B} samples per step: forward + losses + backward + Adam, oracle port of the reference loop (torch CPU ops)'}
        del pts, pnet, small
        torch.cuda.empty_cache()

    with torch.cuda.stream(stream):
        ts.build_graph(*batch)
        for _ in range(args.warmup):
            ts.step_graph(*batch, sync=False)
    barrier()

    def timed(fn, steps, sample=False):
        barrier()
        sampler = ClockSampler(local).start() if sample else None
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record(stream)
            for _ in range(steps):
                fn()
            e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        clocks = sampler.stop() if sampler else None
        if dist is not None:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, clocks

    ms, clocks = timed(lambda: ts.step_graph(*batch, sync=False), args.steps, sample=(rank == 0))
    last = {}

    def e2e_step():
        last['L'], _ = ts.step_graph(*[host[k] for k in keys], sync=True)     # H2D of the batch + D2H of the losses
    ms_e2e, _ = timed(e2e_step, args.steps)
    # eager step with CUDA events around every library call: per-kernel-family time and executed FLOPs
    ts.profile, ts.concurrent = True, False          # one stream: with overlapping branches a CUDA-event pair times its neighbours too
    with torch.cuda.stream(stream):
        ts(*batch, sync=False)
    torch.cuda.synchronize()
    ts.profile, ts.concurrent = False, True
    agg = collections.defaultdict(lambda: [0.0, 0, 0.0])
    for name, a, b_, fl in ts.tape.prof:
        r = agg[name.split(' ')[0]]
        r[0] += a.elapsed_time(b_)
        r[1] += 1
        r[2] += fl
    launches, gflop = ts.last_launches, ts.last_flops / 1e9
    # the roofline kernel: forward / input gradient of the 3x3 convs of the C0 = 32 branch (135 launches per step, the critical
    # path of every HRNet module).  The step is flat -- profiles/r2_train_launches.csv: no (kernel, shape) group above 7 % of the
    # serialised launch time, this one 5.5 % -- so a second kernel (the tcgen05 weight gradients) is timed beside it.  Timed alone -- ten launches replayed as a CUDA graph, because an eager launch
    # from Python costs more host time than this kernel runs
    from rsgnet_b200.train.tape import Tape as _Tape
    dk = None
    try:
        c0 = int(cfg.MODEL.EXTRA.STAGE2.NUM_CHANNELS[0])
        fh, fw = int(cfg.MODEL.IMAGE_SIZE[1]) // 4, int(cfg.MODEL.IMAGE_SIZE[0]) // 4
        with torch.cuda.stream(stream):
            tp = _Tape(dev, 0)
            xa = torch.randn(B, fh, fw, c0, device=dev)
            wk = torch.randn(9, c0, c0, device=dev) / (9 * c0) ** 0.5
            ya = torch.empty(B, fh, fw, c0, device=dev)
            run = lambda: tp._gemm(0, xa, wk, ya, None, B * fh * fw, c0, c0, c0, c0, c0, mode=1, transB=1, geom=(fh, fw, fh, fw, 3, 3, 1, 1))
            run()
            torch.cuda.synchronize()
            gph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gph, stream=stream):
                for _ in range(10):
                    run()
            gph.replay()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            gph.replay()
            e1.record(stream)
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / 10
        fl = 2.0 * B * fh * fw * c0 * c0 * 9
        dk = {'us_per_launch': us, 'flops': fl, 'tflops': fl / us / 1e6, 'bytes': 2.0 * B * fh * fw * c0 * 4,
              'shape': f'{c0}->{c0} 3x3 @{fh}x{fw} x {B} samples'}
        # second kernel: the tcgen05 weight-gradient kernel on the 2 C0 branch (64 launches per step), same method
        c1, h1, w1 = 2 * c0, fh // 2, fw // 2
        with torch.cuda.stream(stream):
            xb = torch.randn(B, h1, w1, c1, device=dev)
            gb = torch.randn(B, h1, w1, c1, device=dev)
            dwb = torch.zeros(9, c1, c1, device=dev)
            runw = lambda: tp._wgrad(0, xb, gb, dwb, B * h1 * w1, c1, c1, mode=1, geom=(h1, w1, h1, w1, 3, 3, 1, 1))
            runw()
            torch.cuda.synchronize()
            gw = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gw, stream=stream):
                for _ in range(10):
                    runw()
            gw.replay()
            e0.record(stream)
            gw.replay()
            e1.record(stream)
        torch.cuda.synchronize()
        usw = e0.elapsed_time(e1) * 1e3 / 10
        flw = 2.0 * B * h1 * w1 * c1 * c1 * 9
        dk['wgrad'] = {'kernel': 'wgrad_tc5_kernel (tcgen05 kind::f16 on bf16 hi/lo splits, 3 MMAs per product)',
                       'shape': f'{c1}->{c1} 3x3 @{h1}x{w1} x {B} samples', 'us_per_launch': usw, 'algorithmic_flops_per_launch': flw,
                       'achieved_tflops': flw / usw / 1e6, 'frac_of_tf32_peak': flw / usw / 1e6 / (pk['tf_sus'] / 2.0),
                       'launches_per_step': 64}
    except Exception as e:                      # informational: never fails the bench
        dk = {'error': str(e)[:200]}
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    per_ms = ms / args.steps
    tot = sum(r[0] for r in agg.values())
    fam = {k: {'ms': round(r[0], 3), 'launches': r[1], 'tflops': (r[2] / r[0] / 1e9 if r[0] and r[2] else 0.0),
               'share': round(r[0] / tot, 4)} for k, r in sorted(agg.items(), key=lambda kv: -kv[1][0])}
    g = agg['rsg_train_gemm']
    tf32_peak = pk['tf_sus'] / 2.0
    h2d = int(sum(host[k].numel() * 4 for k in keys))
    line = {
        'metric': METRIC, 'value': B * world * args.steps / (ms * 1e-3), 'unit': 'samples/s', 'n_gpus': world, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': per_ms, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'tf32', 'data': 'synthetic',
        'config': {'workload': WORKLOAD, 'samples_per_gpu_per_step': B, 'cuda_graph': True, 'optimizer': 'Adam lr 1e-3',
                   'weights': 'random-init trained-like (synth_state_dict seed 4)',
                   'l2': 'activations of a step (~12 GB) exceed the 126 MB L2; no explicit flush'},
        'clocks': clocks, 'gpu_launches': launches * args.steps,
        'e2e': {'value': B * world * args.steps / (ms_e2e * 1e-3), 'unit': 'samples/s', 'h2d_bytes_per_step': h2d,
                'd2h_bytes_per_step': 8 * (3 + B), 'ms_per_step': ms_e2e / args.steps, 'last_loss': last['L']['loss']},
        'roofline': {'bound': 'tensor', 'kernel': 'conv3x3_tf32_small_kernel (tcgen05 kind::tf32 flat 3x3 conv): ' + str(dk.get('shape')),
                     'achieved': dk.get('tflops'), 'peak': tf32_peak, 'unit': 'TFLOP/s',
                     'frac': (dk['tflops'] / tf32_peak) if 'tflops' in dk else None,
                     'traffic': TRAIN_NCU_TRAFFIC, 'traffic_source': 'constant from the committed ncu --set full capture '
                     'profiles/r2_train_conv3x3_small_ncu_details.txt (dram__bytes_read.sum + dram__bytes_write.sum of one launch; not '
                     'measured in this run; the 12.6 MB the kernel writes stay in the 126 MB L2)',
                     'peak_source': pk['src'] + ' bf16_tflops_sustained / 2 (TF32 runs at half the bf16 rate)',
                     'second_kernel': dk.get('wgrad'),
                     'algorithmic_flops_per_launch': dk.get('flops'), 'algorithmic_bytes_per_launch': dk.get('bytes'),
                     'us_per_launch': dk.get('us_per_launch'), 'launches_per_step': 135,
                     'share_of_serialised_launch_time': 0.055,
                     'hbm_frac': (dk['bytes'] / dk['us_per_launch'] / 1e3 / pk['hbm']) if 'us_per_launch' in dk else None,
                     'matrix_family_eager': {'tflops': g[2] / g[0] / 1e9, 'ms': g[0], 'launches': g[1], 'share_of_step': g[0] / tot},
                     'families': fam,
                     'note': 'achieved = the dominant kernel timed alone by graph replay; families = an eager single-stream step with CUDA '
                             'events around every library call (small kernels are padded there by the host launch gap); the '
                             'timed steps replay one CUDA graph whose HRNet branches run concurrently'},
        'executed_gflop_per_step_per_gpu': gflop, 'executed_tflops_per_gpu': gflop / per_ms,
        'parity_checked': bool(parity), 'parity': parity, 'cpu_baseline': cpu,
    }
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=30)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours')
    ap.add_argument('--batch', type=int, default=256, help='crops per GPU per step')
    ap.add_argument('--chunk', type=int, default=0, help='forwards per plan pass (0 = model default)')
    ap.add_argument('--cpu-sample', type=int, default=32, help='crops per step of the CPU arms (bounded sample of the workload)')
    ap.add_argument('--no-graph', action='store_true')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--dump-profile', default='')
    ap.add_argument('--sustained-steps', type=int, default=150, help='extra timed run of this many steps (0 = skip)')
    ap.add_argument('--gpu-eager-baseline', action='store_true',
                    help='also time the reference op graph eagerly on the GPU through PyTorch (informational)')
    ap.add_argument('--workload', default='w32_crowdpose', choices=sorted(WORKLOADS),
                    help='default = the headline configuration (BASELINE.json configs[1]); w48_coco_384 = configs[2], '
                         'a secondary measurement (use --batch 128)')
    args = ap.parse_args()
    global PRESET, WORKLOAD, METRIC
    PRESET = args.workload
    WORKLOAD, METRIC = WORKLOADS[PRESET]
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == 'reference':
        return run_reference(args)
    if PRESET == 'postproc':
        return run_postproc(args)
    if PRESET == 'train':
        return run_train(args)

    from rsgnet_b200.pipeline import CropPipeline
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        # stdout carries exactly one JSON line: NCCL writes its version banner (any NCCL_DEBUG level >= VERSION, which
        # includes WARN) to fd 1 when the communicator is created, so fd 1 points at stderr until the first collective
        # has run
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group('nccl', device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    cfg, net, sd = build_model(dev)
    B = args.batch
    K = int(cfg.MODEL.NUM_JOINTS)
    pipe = CropPipeline(net, cfg, B, use_graph=not args.no_graph, chunk=args.chunk or None)
    # each rank gets its own shard of the synthetic crop stream (no inter-GPU traffic)
    x_host = torch.from_numpy(synth.crops(B, cfg.MODEL.IMAGE_SIZE, seed=100 + rank)).pin_memory()
    c_np, s_np = synth.centers_scales(B, seed=100 + rank)
    c_host, s_host = torch.from_numpy(c_np).pin_memory(), torch.from_numpy(s_np).pin_memory()
    preds_host = torch.empty((B, K, 2), dtype=torch.float32).pin_memory()
    mv_host = torch.empty((B, K, 1), dtype=torch.float32).pin_memory()
    stream = torch.cuda.Stream(dev)

    def step_device():
        return pipe.run_device()

    def step_e2e():
        pipe.x.copy_(x_host, non_blocking=True)
        pipe.center.copy_(c_host, non_blocking=True)
        pipe.scale.copy_(s_host, non_blocking=True)
        p, m = pipe.run_device()
        preds_host.copy_(p, non_blocking=True)
        mv_host.copy_(m, non_blocking=True)

    def timed(fn, steps, sample_clocks=False):
        barrier()
        sampler = ClockSampler(local).start() if sample_clocks else None
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record(stream)
            for _ in range(steps):
                fn()
            e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        clocks = sampler.stop() if sampler else None
        if dist is not None:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, clocks

    with torch.cuda.stream(stream):
        step_e2e()                      # H2D once so that the resident-input arm has real data
        for _ in range(args.warmup):
            step_device()
    torch.cuda.synchronize()
    ms_dev, clocks = timed(step_device, args.steps, sample_clocks=(rank == 0))
    launches = pipe.launches_per_step() * args.steps
    # e2e: the public streaming call -- pinned host batches in, pinned host results out; the H2D copy of
    # step i+1 overlaps the compute of step i (every step's copies are inside the timed region)
    host_batches = [(x_host, c_host, s_host)] * args.steps
    host_out = [(preds_host, mv_host)] * args.steps

    def e2e_all():
        pipe.run_overlapped(host_batches, out=host_out)

    with torch.cuda.stream(stream):
        pipe.run_overlapped(host_batches[:5], out=host_out[:5])
    torch.cuda.synchronize()
    ms_e2e, _ = timed(e2e_all, 1)

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel family (per-op CUDA-event timing, live, one chunk) ----
    pk = peaks()
    eng = pipe.engine
    nb = min(eng.chunk, pipe.n_fwd)
    with torch.cuda.stream(stream):
        eng.profile(pipe.x, pipe.heat, nb, B)                     # warm
        ms_op, kind, flops, names, shapes = eng.profile(pipe.x, pipe.heat, nb, B)
    op_bytes = eng.last_profile_bytes
    fam, groups = {}, {}
    for m, k, f, sh, by in zip(ms_op, kind, flops, shapes, op_bytes):
        if m < 0:
            continue
        d = fam.setdefault(FAMILY[int(k)], [0.0, 0.0, 0, 0.0])
        d[0] += float(m); d[1] += float(f); d[2] += 1; d[3] += float(by)
        g = groups.setdefault((FAMILY[int(k)], sh), [0.0, 0.0, 0, 0.0])
        g[0] += float(m); g[1] += float(f); g[2] += 1; g[3] += float(by)
    total_ms = sum(v[0] for v in fam.values())
    # dominant kernel = the (kernel, shape) group with the largest share of the step.  Its roofline is the one
    # that binds: the larger of (algorithmic bytes / measured HBM peak) and (executed FLOPs / measured bf16 peak).
    (dom_fam, dom_shape), (dom_ms, dom_fl, dom_n, dom_by) = max(groups.items(), key=lambda kv: kv[1][0])
    per_launch_ms = dom_ms / dom_n
    tf = dom_fl / dom_n / (per_launch_ms * 1e-3) / 1e12 if dom_ms > 0 else 0.0
    gbs = dom_by / dom_n / (per_launch_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
    floor_tensor_us = dom_fl / dom_n / (pk['tf_sus'] * 1e12) * 1e6
    floor_hbm_us = dom_by / dom_n / (pk['hbm'] * 1e9) * 1e6
    traffic = NCU_TRAFFIC.get((dom_fam, dom_shape))
    hbm_bound = floor_hbm_us >= floor_tensor_us
    roofline = {'bound': 'hbm' if hbm_bound else 'tensor',
                'kernel': f'{dom_fam}: {dom_shape} x {nb} forwards',
                'achieved': gbs if hbm_bound else tf,
                'peak': pk['hbm'] if hbm_bound else pk['tf_sus'],
                'unit': 'GB/s' if hbm_bound else 'TFLOP/s',
                'frac': (gbs / pk['hbm']) if hbm_bound else (tf / pk['tf_sus']),
                'traffic': traffic,
                'peak_source': pk['src'] + (' hbm_gbs' if hbm_bound else ' bf16_tflops_sustained'),
                'algorithmic_bytes_per_launch': dom_by / dom_n, 'algorithmic_flops_per_launch': dom_fl / dom_n,
                'floor_us': {'hbm': floor_hbm_us, 'tensor': floor_tensor_us},
                'other_bound': {'tflops': tf, 'frac_of_bf16_peak': tf / pk['tf_sus'], 'gbs': gbs, 'frac_of_hbm_peak': gbs / pk['hbm']},
                'us_per_launch': per_launch_ms * 1e3,
                'launches_per_step': dom_n, 'share_of_step': dom_ms / total_ms if total_ms else None,
                'families': {k: {'ms': round(v[0], 4), 'tflops': (v[1] / (v[0] * 1e-3) / 1e12 if v[0] > 0 else 0.0),
                                 'gbs_algorithmic': (v[3] / (v[0] * 1e-3) / 1e9 if v[0] > 0 else 0.0),
                                 'launches': v[2], 'share': round(v[0] / total_ms, 4)} for k, v in fam.items()}}
    # operand-fetch ceiling of the tensor pipe for this kernel's MMA shape (measured, tools/umma_rate.cu, profiles/r1_notes.md §1):
    # an M=128, K=16 MMA with N output columns needs N/2 cycles of math but (4096 + 32 N) / 128 cycles of shared-memory
    # operand fetch, so thin layers cannot exceed (N/2) / (32 + N/4) of the dense peak whatever the kernel does
    import re
    mm = re.search(r'->(\d+)', dom_shape)
    n_cols = int(mm.group(1)) if mm else int(pipe.spec.head_channels)
    n_cols = min(n_cols, 256)
    cap = min(1.0, (n_cols / 2.0) / (32.0 + n_cols / 4.0))
    roofline['ceiling'] = {'frac': cap, 'mma_n': n_cols, 'frac_of_ceiling': (tf / pk['tf_sus']) / cap if cap else None,
                           'why': 'shared-memory operand fetch of an M=128 K=16 MMA: (4096 + 32 N) / 128 cycles vs N / 2 cycles of math'}
    roofline['traffic_source'] = ('constant from the committed ncu --set full capture under profiles/ (not measured in this run)'
                                  if traffic else None)
    if args.dump_profile:
        with open(args.dump_profile, 'w') as f:
            for m, k, fl, nm, sh in zip(ms_op, kind, flops, names, shapes):
                f.write(f'{nm}\t{FAMILY[int(k)]}\t{m:.5f}\t{fl:.0f}\t{sh}\n')

    crops_per_step = B * world
    value = crops_per_step * args.steps / (ms_dev * 1e-3)
    e2e_val = crops_per_step * args.steps / (ms_e2e * 1e-3)
    h2d = x_host.numel() * 4 + c_host.numel() * 4 + s_host.numel() * 4
    d2h = preds_host.numel() * 4 + mv_host.numel() * 4
    exec_tflops = value * 2 * eng.flops_per_fwd / 1e12 / world

    # ---- parity of what was just timed: the first crops of the last step against the CPU oracle (fp32 reference port)
    parity = None
    if not args.no_cpu_baseline:
        cores = len(os.sched_getaffinity(0))
        torch.set_num_threads(cores)
        n = 4
        from oracle import decode_oracle
        o_preds, o_mv, o_avg = cpu_reference_step(sd, cfg, x_host[:n].clone(), c_np[:n], s_np[:n], want_maps=True)
        g_preds, g_mv = preds_host[:n].numpy(), mv_host[:n].numpy()
        # (1) heat-maps of the timed step (bf16 pipeline) against the fp32 oracle maps, flip-averaged the reference's way
        hm, hf = pipe.heat[:n].cpu().numpy(), pipe.heat[B:B + n].cpu().numpy()
        g_avg = decode_oracle.flip_average(hm, hf, presets.flip_pairs_for(K), shift=True)
        hm_err = float(np.abs(g_avg - o_avg).max() / np.abs(o_avg).max())
        # (2) the decode of the timed step is bit-exact given its own heat-maps
        d_preds, d_mv = decode_oracle.get_final_preds(True, g_avg, c_np[:n], s_np[:n])
        decode_exact = bool(np.array_equal(d_preds, g_preds) and np.array_equal(d_mv, g_mv))
        # (3) arg-max agreement with the fp32 maps; a differing arg-max must be a near-tie of the REFERENCE map (its two
        # candidates closer than twice the map's own error), i.e. explained by the stated heat-map tolerance
        ga, oa = g_avg.reshape(n * K, -1), o_avg.reshape(n * K, -1)
        ig, io = ga.argmax(1), oa.argmax(1)
        rows = np.arange(n * K)
        err_map = np.abs(ga - oa).max(1)
        unexplained = int(((ig != io) & (oa[rows, io] - oa[rows, ig] > 2.0 * err_map)).sum())
        px = (s_np[:n, 0] * 200.0 / float(pipe.spec.heat_w))[:, None]
        disp = np.hypot(g_preds[..., 0] - o_preds[..., 0], g_preds[..., 1] - o_preds[..., 1]) / px
        rng = oa.max(1) - oa.min(1)
        parity = {'crops': n, 'heatmap_rel_err': hm_err, 'per_map_norm_err_max': float((err_map / np.maximum(rng, 1e-12)).max()),
                  'decode_bit_exact_given_own_heatmaps': decode_exact, 'argmax_agree': float((ig == io).mean()),
                  'unexplained_argmax_flips': unexplained, 'kpt_disp_le_1px': float((disp <= 1.0).mean()),
                  'maxvals_rel_err': float(np.abs(g_mv - o_mv).max() / max(float(np.abs(o_mv).max()), 1e-12)),
                  'bar': 'bf16 pipeline vs fp32 oracle: averaged heat-maps within 0.05 * max|ref|; decode bit-exact given the '
                         'same heat-maps; every differing arg-max is a near-tie of the fp32 map (gap <= 2 x that map\'s error)'}
        if not (hm_err <= 0.05 and decode_exact and unexplained == 0):
            raise SystemExit(f'bench: parity check of the timed outputs failed: {parity}')

    # ---- CPU baseline: BASELINE.json configs[0] (RSGNet-W32 256x192 COCO K=17, batch 32, the reference's CPU path)
    cpu = None
    if not args.no_cpu_baseline:
        cfg0 = presets.preset('w32_coco')
        net0 = pose_rsgnet.get_pose_net(cfg0, False)
        sd0 = _params.synth_state_dict(net0, seed=3)
        n0 = args.cpu_sample
        x0 = torch.from_numpy(synth.crops(n0, cfg0.MODEL.IMAGE_SIZE, seed=0))
        c0, s0 = synth.centers_scales(n0, seed=0)
        cpu_reference_step(sd0, cfg0, x0[:2], c0[:2], s0[:2])        # warm (thread pool, oneDNN primitives)
        ts = []
        for _ in range(3):
            t0 = time.perf_counter()
            cpu_reference_step(sd0, cfg0, x0, c0, s0)
            ts.append(time.perf_counter() - t0)
        dt = float(np.median(ts))
        cpu = {'value': n0 / dt, 'unit': 'crops/s', 'cores': cores, 'kind': 'port',
               'sample': f'BASELINE.json configs[0]: RSGNet-W32 256x192 COCO K=17, batch {n0}, flip test (2 fp32 forwards/crop + '
                         f'flip-average + get_final_preds), median of 3 reps, oracle port of the reference torch/NumPy path'}

    # ---- sustained figure: the same step for >= 2.5 s (the power cap settles after ~1 s)
    sustained = None
    if args.sustained_steps > 0 and world == 1:
        ms_sus, clocks_sus = timed(step_device, args.sustained_steps, sample_clocks=True)
        sustained = {'steps': args.sustained_steps, 'ms_per_step': ms_sus / args.sustained_steps,
                     'value': crops_per_step * args.sustained_steps / (ms_sus * 1e-3), 'clocks': clocks_sus}

    # ---- informational: the reference's op graph run EAGERLY on this GPU by PyTorch (cuDNN / cuBLAS), fp32 and bf16
    # autocast -- SURVEY.md §2.2 names it as the real bar; not our path, never part of `value`
    eager = None
    if args.gpu_eager_baseline:
        eager = gpu_eager_baseline(sd, cfg, x_host, dev)
    dropin = None
    if not args.no_cpu_baseline:
        try:
            dropin = dropin_loop_throughput(net, cfg, x_host, c_np, s_np, dev)
        except Exception as e:                       # informational leg
            dropin = {'error': str(e)[:300]}

    line = {
        'metric': METRIC, 'value': value, 'unit': 'crops/s', 'n_gpus': world, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': ms_dev / args.steps, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'bf16', 'data': 'synthetic',
        'config': {'workload': WORKLOAD, 'crops_per_gpu_per_step': B, 'forwards_per_crop': 2,
                   'chunk_forwards': eng.chunk, 'cuda_graph': not args.no_graph,
                   'l2': 'inputs (151 MB/step/GPU) and activations exceed the 126 MB L2; no explicit flush',
                   'weights': 'random-init trained-like (synth_state_dict seed 4)'},
        'clocks': clocks, 'gpu_launches': launches,
        'e2e': {'value': e2e_val, 'unit': 'crops/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                'ms_per_step': ms_e2e / args.steps},
        'roofline': roofline, 'cpu_baseline': cpu, 'parity_checked': parity is not None, 'parity': parity,
        'sustained': sustained, 'gpu_eager_baseline': eager, 'dropin_loop': dropin,
        'executed_tflops_per_gpu': exec_tflops, 'executed_frac_of_bf16_peak': exec_tflops / pk['tf_sus'],
        'reference_graph_tflops_per_gpu': value / world * REF_GFLOP_PER_CROP / 1e3,
    }
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
