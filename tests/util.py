"""Shared helpers for the tests: golden loading, case configs, error metrics."""
import importlib
import os

import numpy as np

from make_golden import CASES, net_kwargs  # tests/golden/make_golden.py (no reference import at load)

synth = importlib.import_module("pl-nerf_b200.synth")
GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    with np.load(os.path.join(GOLDEN_DIR, name + ".npz")) as z:
        return {k: z[k] for k in z.files}


def case_params(name):
    cfg = CASES[name]
    kw = net_kwargs(cfg)
    return cfg, kw, synth.nerf_params(cfg["seeds"][0], **kw), synth.nerf_params(cfg["seeds"][1], **kw)


def oracle_net_kw(kw):
    return dict(D=kw["D"], skips=kw["skips"], input_ch=kw["input_ch"],
                input_ch_views=kw["input_ch_views"], use_viewdirs=kw["use_viewdirs"])


def max_rel(a, b, floor=1e-3):
    """max |a-b| / max(|b|, floor): relative error with an absolute floor for values near 0."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor))) if a.size else 0.0
