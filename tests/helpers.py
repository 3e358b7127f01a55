"""Shared test helpers: seeded model construction, deterministic weight perturbation, error metrics."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


from planerecnet_b200.utils.synth import make_input, perturb_  # noqa: E402,F401


def rel_l2(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / (b.norm() + 1e-30))


def max_rel(a, b):
    return float((a.double() - b.double()).abs().max() / (b.double().abs().max() + 1e-30))


def build_ours(preset, seed=0):
    from planerecnet_b200.config import cfg, set_cfg
    from planerecnet_b200.planerecnet import PlaneRecNet
    set_cfg(preset)
    torch.manual_seed(seed)
    net = PlaneRecNet(cfg)
    return net
