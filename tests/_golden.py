"""Loader for tests/golden/*.npz (written by tests/golden/make_golden.py)."""
import json
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


class Golden:
    def __init__(self, name):
        raw = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        self.cfg = json.loads(bytes(raw["cfg"]).decode())
        self.groups = {}
        for key in raw.files:
            if key == "cfg":
                continue
            grp, sub = key.split("/", 1)
            self.groups.setdefault(grp, {})[sub] = torch.from_numpy(np.array(raw[key]))

    def __getitem__(self, grp):
        return self.groups[grp]

    def sd(self, grp="sd", device=None, dtype=None):
        out = {}
        for k, v in self.groups[grp].items():
            if dtype is not None and v.is_floating_point() and v.dtype == torch.float32:
                v = v.to(dtype)
            out[k] = v.to(device) if device is not None else v
        return out


def rel_err(a, b):
    """max|a-b| / max|b| -- the 'rel' of BASELINE.json's tolerances, on the scale of the tensor."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    denom = b.abs().max().item()
    if denom == 0.0:
        return (a - b).abs().max().item()
    return (a - b).abs().max().item() / denom


def rel_l2(a, b):
    """||a-b||_2 / ||b||_2 -- robust to the isolated ReLU-mask flips a 1e-6 perturbation of h can cause."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    n = b.norm().item()
    return (a - b).norm().item() / (n if n else 1.0)
