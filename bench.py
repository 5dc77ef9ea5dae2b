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


def cpu_reference_step(sd, cfg, x, c, s):
    """One bounded sample of the reference's CPU path (oracle port): 2 forwards, flip-average,
    get_final_preds."""
    from oracle import decode_oracle, model_oracle
    k = int(cfg.MODEL.NUM_JOINTS)
    a = model_oracle.forward(sd, cfg, x)[1].numpy()
    b = model_oracle.forward(sd, cfg, x.flip(3))[1].numpy()
    avg = decode_oracle.flip_average(a, b, presets.flip_pairs_for(k), shift=True)
    return decode_oracle.get_final_preds(True, avg, c, s)


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path (oracle port), all host
    threads, bounded sample per step."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
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
        'data': 'synthetic', 'config': {'workload': WORKLOAD, 'crops_per_step': n},
        'cpu_baseline': {'value': val, 'unit': 'crops/s', 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': val, 'unit': 'crops/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0}))


# dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed
# `ncu --set full` captures (profiles/r1_conv_bb_ncu_details.txt, r1_conv_tc5_ncu_details.txt,
# r1_conv_bneck_ncu_details.txt; 512 forwards)
NCU_TRAFFIC = {('conv_tcgen05', 'conv 32->32 taps9 s1 64x48 +res'): 201.44e6 + 72.35e6,
               ('basic_block_tcgen05', 'bblock'): 100.78e6 + 54.63e6,
               ('bottleneck_tcgen05', 'bneck'): 902.87e6 + 759.45e6}

FAMILY = {0: 'stem', 1: 'conv_mma', 2: 'conv_tcgen05', 3: 'fuse', 4: 'maxpool', 5: 'trp_attention',
          6: 'relation_scores', 7: 'groupnorm', 8: 'bilinear', 9: 'conv_ws_tcgen05', 10: 'basic_block_tcgen05',
          11: 'bottleneck_tcgen05'}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=30)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours')
    ap.add_argument('--batch', type=int, default=256, help='crops per GPU per step')
    ap.add_argument('--chunk', type=int, default=0, help='forwards per plan pass (0 = model default)')
    ap.add_argument('--cpu-sample', type=int, default=4, help='crops per CPU-baseline step')
    ap.add_argument('--no-graph', action='store_true')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--dump-profile', default='')
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

    cpu = None
    if not args.no_cpu_baseline:
        cores = len(os.sched_getaffinity(0))
        torch.set_num_threads(cores)
        n = args.cpu_sample
        xs = x_host[:n].clone()
        cpu_reference_step(sd, cfg, xs, c_np[:n], s_np[:n])        # warm
        t0 = time.perf_counter()
        reps = 2
        for _ in range(reps):
            cpu_reference_step(sd, cfg, xs, c_np[:n], s_np[:n])
        dt = (time.perf_counter() - t0) / reps
        cpu = {'value': n / dt, 'unit': 'crops/s', 'cores': cores, 'kind': 'port',
               'sample': f'{n} crops x {reps} reps (flip-test: 2 fp32 forwards/crop + flip-average + decode), '
                         'oracle port of the reference torch/NumPy path'}

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
        'roofline': roofline, 'cpu_baseline': cpu,
        'executed_tflops_per_gpu': exec_tflops,
        'reference_graph_tflops_per_gpu': value / world * REF_GFLOP_PER_CROP / 1e3,
    }
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
