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


# torch stand-ins for the two small kernels of the training step (plnerf_mse_loss_grad / plnerf_adam_step), for the host-logic
# tests that run without a GPU: written with the ops of img2mse and of torch's single-tensor Adam, so that a TrainStep driven
# by them stays bit-identical to the reference sequence with stock optimisers
def fake_mse_loss_grad(rgb, rgb0, target, scale, sqerr, pix=None):
    t = target if pix is None else target[pix]
    d = rgb - t
    sqerr[0] += (d * d).sum()
    if rgb0 is None:
        return d * scale, None
    d0 = rgb0 - t
    sqerr[1] += (d0 * d0).sum()
    return d * scale, d0 * scale


def fake_adam_step(params, grads, exp_avg, exp_avg_sq, lr, step, betas=(0.9, 0.999), eps=1e-8, zero_grads=False):
    b1, b2 = betas
    exp_avg.lerp_(grads, 1 - b1)
    exp_avg_sq.mul_(b2).addcmul_(grads, grads, value=1 - b2)
    denom = (exp_avg_sq.sqrt() / ((1 - b2 ** step) ** 0.5)).add_(eps)
    params.addcdiv_(exp_avg, denom, value=-(lr / (1 - b1 ** step)))
    if zero_grads:
        grads.zero_()
