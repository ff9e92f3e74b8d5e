"""Shared helpers for the tests (inputs of BASELINE.md's measurement recipe)."""
import json
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def make_image(i, H, W):
    g = torch.Generator().manual_seed(1000 + i)
    return (torch.rand(3, H, W, generator=g) * 255).round().to(torch.uint8)


def golden_summary():
    with open(os.path.join(GOLDEN, 'summary.json')) as f:
        return json.load(f)


def golden_batch_bytes():
    """container sizes the unmodified reference writes for every image of the benchmark batches
    (oracle/gen_golden_batch.py): {'l3c_512x512': [...16], 'rgbs_256x256': [...32]}"""
    with open(os.path.join(GOLDEN, 'batch_bytes.json')) as f:
        return json.load(f)


def bpsp_deltas(sizes, ref_sizes, subpixels):
    """signed per-image bpsp differences to the reference's containers"""
    return (np.asarray(sizes, np.float64) - np.asarray(ref_sizes, np.float64)) * 8.0 / subpixels


def golden_npz(name):
    return np.load(os.path.join(GOLDEN, name + '.npz'))


_BP = {}


def blueprint(cfg_name='cr', device=None):
    """seed-0 default-init blueprint (identical weights to the reference's, see test_host_logic)."""
    from l3c_pytorch_b200 import config
    from l3c_pytorch_b200.blueprint import MultiscaleBlueprint
    key = (cfg_name, str(device))
    if key not in _BP:
        torch.manual_seed(0)
        bp = MultiscaleBlueprint(config.ms_config(cfg_name), device=device)
        bp.set_eval()
        _BP[key] = bp
    return _BP[key]


def cpu_state_dict(bp):
    return {k: v.detach().cpu().clone() for k, v in bp.net.state_dict().items()}


def oracle_cfg(cfg_name):
    from oracle import model as om
    return om.CFG_L3C if cfg_name == 'cr' else om.CFG_RGB_SHARED
