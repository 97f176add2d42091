"""Goldens of the depth-experiment sampler variants from the UNMODIFIED reference functions
run_nerf_helpers.sample_pdf_reformulation_return_u (:448-533) and sample_pdf_return_u (:286-337), fed with the coarse-pass
tensors of the existing render goldens (build container only):

    python tests/golden/make_golden_return_u.py
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
    # piecewise-linear variant on two linear-mode cases; explicit u (load_u) and the pytest draw
    for name in ("lego_linear_mid", "llff_ndc_linear"):
        g = load(name)
        t = lambda k: torch.from_numpy(g[k])
        rb = t("ray_batch")
        near, far = rb[:, 6:7], rb[:, 7:8]
        Ni = g["u"].shape[1]
        for tag, kw in (("load", dict(load_u=t("u"))), ("pytest", dict(pytest=True))):
            r = H.sample_pdf_reformulation_return_u(t("z_vals0"), t("weights0"), t("tau0"), t("T0"), near, far, Ni, **kw)
            for key, v in zip(("samples", "T_below", "tau_below", "bin_below", "u"), r):
                out[f"{name}.pl.{tag}.{key}"] = v.numpy()
    # piecewise-constant variant on the constant-mode case (bins = z_mid, weights[..., 1:-1], run_plnerf.py:726)
    g = load("llff_ndc_constant")
    z = torch.from_numpy(g["z_vals0"])
    w = torch.from_numpy(g["weights0"])
    z_mid = .5 * (z[..., 1:] + z[..., :-1])
    Ni = g["u"].shape[1]
    for tag, kw in (("load", dict(load_u=torch.from_numpy(g["u"]))), ("pytest", dict(pytest=True))):
        s, u = H.sample_pdf_return_u(z_mid, w[..., 1:-1], Ni, **kw)
        out[f"llff_ndc_constant.const.{tag}.samples"] = s.numpy()
        out[f"llff_ndc_constant.const.{tag}.u"] = u.numpy()
    path = os.path.join(HERE, "return_u.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path} ({os.path.getsize(path) / 1024:.0f} KiB), {len(out)} arrays")


if __name__ == "__main__":
    main()
