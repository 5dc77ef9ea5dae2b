"""Helpers for the GPU parity tests."""
import numpy as np
import torch

from rsgnet_b200 import presets, synth
from rsgnet_b200.models import _params, pose_hrnet, pose_rsgnet


def build(key, seed, device='cuda'):
    cfg = presets.preset(key)
    mod = pose_rsgnet if cfg.MODEL.NAME == 'pose_rsgnet' else pose_hrnet
    net = mod.get_pose_net(cfg, False)
    sd = _params.synth_state_dict(net, seed=seed)
    net.load_state_dict(sd, strict=True)
    net = net.to(device).eval()
    return cfg, net, sd


def crops(cfg, n, seed):
    return torch.from_numpy(synth.crops(n, cfg.MODEL.IMAGE_SIZE, seed=seed))


def rel_err(got, ref):
    got = np.asarray(got, np.float64)
    ref = np.asarray(ref, np.float64)
    return float(np.abs(got - ref).max() / max(np.abs(ref).max(), 1e-12))
