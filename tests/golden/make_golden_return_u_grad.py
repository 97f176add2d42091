"""Gradient goldens of the depth-experiment sampler variants (differentiable form): torch autograd through
the UNMODIFIED reference functions run_nerf_helpers.sample_pdf_reformulation_return_u (:448-533) and sample_pdf_return_u
(:286-337), fed with the coarse-pass tensors of the existing render goldens and fixed cotangents (build container only):

    python tests/golden/make_golden_return_u_grad.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import refimport  # noqa: E402


def load(name):
    with np.load(os.path.join(HERE, name + ".npz")) as z:
        return {k: z[k] for k in z.files}


def main():
    H, R = refimport.load()
    out = {}
    rs = np.random.RandomState(11)
    for name in ("lego_linear_mid", "llff_ndc_linear"):
        g = load(name)
        leaf = lambda a: torch.from_numpy(np.ascontiguousarray(a)).clone().requires_grad_(True)
        rb = g["ray_batch"]
        z, w, tau, T = leaf(g["z_vals0"]), leaf(g["weights0"]), leaf(g["tau0"]), leaf(g["T0"])
        near, far = leaf(rb[:, 6:7]), leaf(rb[:, 7:8])
        u = torch.from_numpy(g["u"])
        outs = H.sample_pdf_reformulation_return_u(z, w, tau, T, near, far, u.shape[1], load_u=u)
        cot = [torch.from_numpy(rs.randn(*u.shape).astype(np.float32)) for _ in range(4)]
        loss = sum((o * c).sum() for o, c in zip(outs[:4], cot))
        loss.backward()
        assert w.grad is None or not w.grad.any()           # the weights only choose the bracket
        for key, c in zip(("samples", "T_below", "tau_below", "bin_below"), cot):
            out[f"{name}.pl.cot.{key}"] = c.numpy()
        for key, t in (("z", z), ("tau", tau), ("T", T), ("near", near), ("far", far)):
            out[f"{name}.pl.grad.{key}"] = t.grad.numpy()
        # samples only (what a depth loss on the samples alone sends back)
        for t in (z, w, tau, T, near, far):
            t.grad = None
        outs = H.sample_pdf_reformulation_return_u(z, w, tau, T, near, far, u.shape[1], load_u=u)
        (outs[0] * cot[0]).sum().backward()
        for key, t in (("z", z), ("tau", tau), ("T", T), ("near", near), ("far", far)):
            out[f"{name}.pl.grad_samples_only.{key}"] = t.grad.numpy()
    g = load("llff_ndc_constant")
    z = torch.from_numpy(g["z_vals0"])
    bins = (.5 * (z[..., 1:] + z[..., :-1])).clone().requires_grad_(True)
    w = torch.from_numpy(g["weights0"])[..., 1:-1].clone().requires_grad_(True)
    u = torch.from_numpy(g["u"])
    s, _ = H.sample_pdf_return_u(bins, w, u.shape[1], load_u=u)
    cot = torch.from_numpy(rs.randn(*u.shape).astype(np.float32))
    (s * cot).sum().backward()
    out["llff_ndc_constant.const.cot.samples"] = cot.numpy()
    out["llff_ndc_constant.const.grad.bins"] = bins.grad.numpy()
    out["llff_ndc_constant.const.grad.weights"] = w.grad.numpy()
    path = os.path.join(HERE, "return_u_grad.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path} ({os.path.getsize(path) / 1024:.0f} KiB), {len(out)} arrays")
    for k, v in out.items():
        if ".grad" in k:
            print(f"  {k:50s} max|g| = {np.abs(v).max():.4g}  nonzero = {(v != 0).mean():.3f}")


if __name__ == "__main__":
    main()
