"""Goldens at the sizes of BASELINE.json's configs, from the UNMODIFIED reference (build container only).

    python tests/golden/make_golden_sized.py

  c1_coarse_4096   config 1: one 4096-ray batch, coarse only (N_samples=64), 8x256 MLP without view directions,
                   linear / midpoint, perturb=1 with the reference's pytest draws.
  c2_lego_1024     config 2 shape: 1024 lego-shaped rays, N_samples=64 + N_importance=128, PL quadrature, view
                   directions, white background, density-boosted coarse + fine nets (the bench's nets).
  train_c3_1024    config 3 shape: 1024 rays at 128 + 64 samples; loss = img2mse(rgb, t) + img2mse(rgb0, t) against a seeded
                   random target, back-propagated through the unmodified reference: the loss, and per parameter of both
                   networks the gradient's L2 norm and its first 2048 entries.
  selfcheck_orders the same network function evaluated by the unmodified reference in two fp32 summation orders
                   (hidden units of every trunk layer permuted consistently: mathematically the identical network),
                   on 256 density-boosted rays at 64/128.  How far the reference disagrees with ITSELF bounds what any
                   other fp32-faithful implementation can be held to; the fine depth is the sensitive output (the
                   inverse-CDF sampler divides by the pdf).

Only outputs are stored (the inputs are regenerated from pl-nerf_b200/synth.py seeds and the pytest draws are the head of
np.random.seed(0)); `inds` (what the reference's torch.searchsorted call returned, recorded by wrapping the call) as int16.
"""
import importlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import refimport  # noqa: E402

synth = importlib.import_module("pl-nerf_b200.synth")

SIZED = {
    "c1_coarse_4096": dict(n=4096, ray_seed=9, Ns=64, Ni=0, use_viewdirs=False, white_bkgd=True, seeds=(41, 41),
                           density_boost=True),
    "c2_lego_1024": dict(n=1024, ray_seed=7, Ns=64, Ni=128, use_viewdirs=True, white_bkgd=True, seeds=(1, 2),
                         density_boost=True),
    "selfcheck_orders": dict(n=256, ray_seed=3, Ns=64, Ni=128, use_viewdirs=True, white_bkgd=True, seeds=(1, 2),
                             density_boost=True),
    # config 3 shape (configs/blender_linear.txt: N_rand=1024, N_samples=128, N_importance=64): the training loss and every
    # parameter gradient of the unmodified reference (loss.backward() through its own render)
    "train_c3_1024": dict(n=1024, ray_seed=13, Ns=128, Ni=64, use_viewdirs=True, white_bkgd=True, seeds=(1, 2),
                          density_boost=True, train=True),
    # the same loss through networks WITHOUT view directions (output_linear head, run_nerf_helpers.py:100-103, :126): the
    # unused views_linears of such a network gets no gradient from autograd (stored as zeros)
    "train_noviews_256": dict(n=256, ray_seed=17, Ns=64, Ni=32, use_viewdirs=False, white_bkgd=True, seeds=(1, 2),
                              density_boost=True, train=True),
}
GRAD_SLICE = 2048      # leading entries of every parameter gradient stored next to its norm


def net_kwargs(cfg):
    return dict(D=8, W=256, input_ch=63, input_ch_views=27 if cfg["use_viewdirs"] else 0,
                output_ch=5 if cfg["Ni"] > 0 else 4, skips=(4,), use_viewdirs=cfg["use_viewdirs"])


def sized_inputs(cfg):
    """(rays_o, rays_d, K, (H, W, focal), params_coarse, params_fine) of a sized case -- shared with the tests."""
    kw = net_kwargs(cfg)
    ro, rd, K, hwf = synth.lego_rays(cfg["n"], seed=cfg["ray_seed"])
    pc = synth.nerf_params(cfg["seeds"][0], density_boost=cfg["density_boost"], **kw)
    pf = synth.nerf_params(cfg["seeds"][1], density_boost=cfg["density_boost"], **kw)
    return ro, rd, K, hwf, pc, pf


def pytest_draws(n, Ns, Ni):
    """What the reference's pytest=True hook draws: every tensor is the head of np.random.seed(0)."""
    def head(shape):
        np.random.seed(0)
        return np.random.rand(*shape).astype(np.float32)
    return head((n, Ns)), (head((n, Ni)) if Ni > 0 else None)


def permute_hidden_units(params, D, skips, seed):
    """The same network with the 256 hidden units of every trunk layer (and the feature / views layers) renumbered:
    rows of layer l and the matching input columns of its consumers are permuted together."""
    rs = np.random.RandomState(seed)
    p = {k: v.copy() for k, v in params.items()}
    W = p["pts_linears.0.weight"].shape[0]
    in_ch = p["pts_linears.0.weight"].shape[1]
    for l in range(D):
        perm = rs.permutation(W)
        p[f"pts_linears.{l}.weight"] = p[f"pts_linears.{l}.weight"][perm]
        p[f"pts_linears.{l}.bias"] = p[f"pts_linears.{l}.bias"][perm]
        consumers = [f"pts_linears.{l + 1}.weight"] if l + 1 < D else (
            ["feature_linear.weight", "alpha_linear.weight"] if "feature_linear.weight" in p else ["output_linear.weight"])
        for c in consumers:
            w = p[c]
            off = in_ch if (c.startswith("pts_linears") and l in skips) else 0      # [input_pts, h] after a skip
            w[:, off:off + W] = w[:, off:off + W][:, perm]
    if "feature_linear.weight" in p:
        perm = rs.permutation(W)
        p["feature_linear.weight"] = p["feature_linear.weight"][perm]
        p["feature_linear.bias"] = p["feature_linear.bias"][perm]
        p["views_linears.0.weight"][:, :W] = p["views_linears.0.weight"][:, :W][:, perm]
        perm2 = rs.permutation(W // 2)
        p["views_linears.0.weight"] = p["views_linears.0.weight"][perm2]
        p["views_linears.0.bias"] = p["views_linears.0.bias"][perm2]
        p["rgb_linear.weight"] = p["rgb_linear.weight"][:, perm2]
    return p


def train_target(cfg):
    return np.random.RandomState(99).rand(cfg["n"], 3).astype(np.float32)


def ref_render(H, R, cfg, ro, rd, K, hwf, pc, pf):
    kw = net_kwargs(cfg)

    def mk(prm):
        net = H.NeRF(D=kw["D"], W=kw["W"], input_ch=kw["input_ch"], input_ch_views=kw["input_ch_views"],
                     output_ch=kw["output_ch"], skips=list(kw["skips"]), use_viewdirs=kw["use_viewdirs"])
        net.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in prm.items()})
        return net
    net_c = mk(pc)
    net_f = mk(pf) if cfg["Ni"] > 0 else None
    embed_fn, _ = H.get_embedder(10, 0)
    embeddirs_fn = H.get_embedder(4, 0)[0] if cfg["use_viewdirs"] else None
    q = lambda p, v, fn: R.run_network(p, v, fn, embed_fn=embed_fn, embeddirs_fn=embeddirs_fn, netchunk=1024 * 64)
    Hh, Ww, focal = hwf
    rays_t = torch.stack([torch.from_numpy(ro), torch.from_numpy(rd)])
    # instrument (not modify) the reference: record what its torch.searchsorted call returns -- the "integer sample
    # indices" of the importance sampler (run_nerf_helpers.py:397)
    captured = []
    real_searchsorted = torch.searchsorted

    def recording_searchsorted(*a, **k):
        r = real_searchsorted(*a, **k)
        captured.append(r.clone())
        return r
    torch.searchsorted = recording_searchsorted
    try:
        with torch.set_grad_enabled(bool(cfg.get("train"))):
            rgb, disp, acc, ex = _render(R, Hh, Ww, K, rays_t, cfg, q, net_c, net_f)
    finally:
        torch.searchsorted = real_searchsorted
    out = {"rgb_map": rgb.detach().numpy(), "disp_map": disp.detach().numpy(), "acc_map": acc.detach().numpy()}
    out.update({k: v.detach().numpy() for k, v in ex.items()})
    if cfg.get("train"):
        tgt = torch.from_numpy(train_target(cfg))
        loss = H.img2mse(rgb, tgt) + H.img2mse(ex["rgb0"], tgt)           # run_plnerf.py:1290-1297
        loss.backward()
        out["train_loss"] = np.float64(loss.item())
        for tag, net in (("c", net_c), ("f", net_f)):
            for k_, p_ in net.named_parameters():
                g_ = (p_.grad if p_.grad is not None else torch.zeros_like(p_)).detach().numpy().ravel()
                out[f"gnorm_{tag}.{k_}"] = np.float64(np.linalg.norm(g_.astype(np.float64)))
                out[f"ghead_{tag}.{k_}"] = g_[:GRAD_SLICE].copy()
    if captured:
        assert len(captured) == 1
        out["inds"] = captured[0].numpy().astype(np.int16)
    return out


def _render(R, Hh, Ww, K, rays_t, cfg, q, net_c, net_f):
    if True:
        rgb, disp, acc, ex = R.render(Hh, Ww, K, chunk=32768, rays=rays_t, ndc=False, near=2., far=6.,
                                      use_viewdirs=cfg["use_viewdirs"], network_query_fn=q, network_fn=net_c,
                                      network_fine=net_f, N_samples=cfg["Ns"], N_importance=cfg["Ni"], perturb=1.0,
                                      raw_noise_std=0., white_bkgd=cfg["white_bkgd"], mode="linear",
                                      color_mode="midpoint", lindisp=False, pytest=True, retraw=False)
    return rgb, disp, acc, ex


def main():
    H, R = refimport.load()
    only = sys.argv[1:]
    for name, cfg in SIZED.items():
        if only and name not in only:
            continue
        ro, rd, K, hwf, pc, pf = sized_inputs(cfg)
        out = ref_render(H, R, cfg, ro, rd, K, hwf, pc, pf)
        if name == "selfcheck_orders":
            kw = net_kwargs(cfg)
            out_b = ref_render(H, R, cfg, ro, rd, K, hwf, permute_hidden_units(pc, kw["D"], kw["skips"], 101),
                               permute_hidden_units(pf, kw["D"], kw["skips"], 102))
            out = {**{k + "_a": v for k, v in out.items()}, **{k + "_b": v for k, v in out_b.items()}}
            d = np.abs(out["depth_map_a"] - out["depth_map_b"]) / 6.0
            print(f"  self-disagreement of the reference (two fp32 orders): fine depth / far max {d.max():.3e} "
                  f"p99 {np.percentile(d, 99):.3e}; rgb max {np.abs(out['rgb_map_a'] - out['rgb_map_b']).max():.3e}; "
                  f"coarse depth / far max {(np.abs(out['depth0_a'] - out['depth0_b']) / 6.0).max():.3e}")
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **{k: v for k, v in out.items()})
        print(f"{name}: wrote {path} ({os.path.getsize(path) / 1024:.0f} KiB), keys={sorted(out)}")


if __name__ == "__main__":
    main()
