"""Golden vectors for the rarely-used options of the path, from the UNMODIFIED reference (build container only):

    python tests/golden/make_golden_flags.py

* ``raw2outputs(..., farcolorfix=True)`` (run_plnerf.py:583-587: the colour appended at the far end is zero) --
  maps + d(loss)/d(raw) through torch autograd, midpoint colour mode;
* ``sample_pdf_reformulation(..., zero_threshold, epsilon_)`` with non-default thresholds (run_nerf_helpers.py:364-445;
  the CLI never forwards these flags -- SURVEY.md Appendix B.5 -- but the functions accept them).

Inputs are the coarse-pass tensors already stored in lego_linear_mid.npz (raw0, z_vals0, ray_batch, weights0, tau0, T0,
u, up_*), so flags_lego.npz only adds outputs.

* ``render(..., perturb=0, mode="constant")``: the deterministic path (no jitter, linspace u) end to end -> det_constant.npz.
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

BASE = "lego_linear_mid"
SAMPLER_FLAGS = dict(zero_threshold=5e-2, epsilon_=1e-2)


def main():
    H, R = refimport.load()
    with np.load(os.path.join(HERE, BASE + ".npz")) as z:
        g = {k: z[k] for k in z.files}
    t = lambda k: torch.from_numpy(g[k].copy())
    rb = t("ray_batch")
    near, far, rays_d = rb[:, 6:7], rb[:, 7:8], rb[:, 3:6]
    out = {}
    # ---- farcolorfix: forward maps and the gradient w.r.t. raw
    raw = t("raw0").requires_grad_(True)
    r = R.raw2outputs(raw, t("z_vals0"), near, far, rays_d, "linear", "midpoint", 0., pytest=True, white_bkgd=True,
                      farcolorfix=True)
    names = ("rgb_map", "disp_map", "acc_map", "weights", "depth_map", "tau", "T")
    for k, v in zip(names, r):
        out["fcf_" + k] = v.detach().numpy()
    loss = (r[0] * t("up_rgb")).sum() + (r[4] * t("up_depth")).sum() + (r[2] * t("up_acc")).sum() + (r[1] * t("up_disp")).sum()
    loss.backward()
    out["fcf_g_raw"] = raw.grad.numpy()
    plain = R.raw2outputs(t("raw0"), t("z_vals0"), near, far, rays_d, "linear", "midpoint", 0., pytest=True, white_bkgd=True)
    assert not np.allclose(plain[0].numpy(), out["fcf_rgb_map"]), "farcolorfix changed nothing on these rays"
    np.testing.assert_array_equal(plain[3].numpy(), out["fcf_weights"])      # only the colours change
    # ---- sampler with non-default thresholds (pytest=True: u = head of np.random.seed(0), = the stored u)
    Ni = g["u"].shape[1]
    with torch.no_grad():
        zs, _, _, _ = H.sample_pdf_reformulation(t("z_vals0"), t("weights0"), t("tau0"), t("T0"), near, far, Ni, det=False,
                                                 pytest=True, **SAMPLER_FLAGS)
        zs_default, _, _, _ = H.sample_pdf_reformulation(t("z_vals0"), t("weights0"), t("tau0"), t("T0"), near, far, Ni,
                                                         det=False, pytest=True)
    out["flags_z_samples"] = zs.numpy()
    frac = float((zs != zs_default).float().mean())
    print(f"sampler: {100 * frac:.1f}% of the samples move with zero_threshold={SAMPLER_FLAGS['zero_threshold']}, "
          f"epsilon={SAMPLER_FLAGS['epsilon_']}")
    assert frac > 0.01
    out["zero_threshold"] = np.float32(SAMPLER_FLAGS["zero_threshold"])
    out["epsilon"] = np.float32(SAMPLER_FLAGS["epsilon_"])
    path = os.path.join(HERE, "flags_lego.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path} ({os.path.getsize(path) / 1024:.0f} KiB), keys={sorted(out)}")


DET = dict(n=16, Ns=64, Ni=64, seeds=(61, 62))


def det_net_kwargs():
    return dict(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=(4,), use_viewdirs=True)


def main_det():
    """perturb = 0 (no stratified jitter, det=True in the sampler: u = linspace(0, 1, N_importance),
    run_nerf_helpers.py:248-250) in constant mode -- in linear mode the reference itself raises on u == 1
    (SURVEY.md 8c caveat 3).  Fully deterministic: no pytest hook needed."""
    import importlib
    synth = importlib.import_module("pl-nerf_b200.synth")
    H, R = refimport.load()
    kw = det_net_kwargs()
    pc, pf = synth.nerf_params(DET["seeds"][0], **kw), synth.nerf_params(DET["seeds"][1], **kw)

    def mk(p):
        net = H.NeRF(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=[4], use_viewdirs=True)
        net.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in p.items()})
        return net
    ro, rd, K, (Hh, Ww, focal) = synth.lego_rays(DET["n"], seed=9)
    embed_fn, _ = H.get_embedder(10, 0)
    embeddirs_fn, _ = H.get_embedder(4, 0)
    q = lambda p, v, fn: R.run_network(p, v, fn, embed_fn=embed_fn, embeddirs_fn=embeddirs_fn, netchunk=1024 * 64)
    with torch.no_grad():
        rgb, disp, acc, extras = R.render(Hh, Ww, K, chunk=32768, rays=torch.stack([torch.from_numpy(ro), torch.from_numpy(rd)]),
                                          ndc=False, near=2., far=6., use_viewdirs=True, network_query_fn=q, network_fn=mk(pc),
                                          network_fine=mk(pf), N_samples=DET["Ns"], N_importance=DET["Ni"], perturb=0.,
                                          raw_noise_std=0., white_bkgd=True, mode="constant", color_mode="midpoint")
    out = {"rays_o": ro, "rays_d": rd, "K": K.astype(np.float32), "hwf": np.array([Hh, Ww, focal], np.float64),
           "rgb_map": rgb.numpy(), "disp_map": disp.numpy(), "acc_map": acc.numpy()}
    out.update({k: v.numpy() for k, v in extras.items()})
    path = os.path.join(HERE, "det_constant.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path} ({os.path.getsize(path) / 1024:.0f} KiB), keys={sorted(out)}")


if __name__ == "__main__":
    main()
    main_det()
