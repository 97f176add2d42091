"""Generate golden vectors from the UNMODIFIED reference (run in the build container only).

    python tests/golden/make_golden.py

Imports /root/reference/run_plnerf.py + run_nerf_helpers.py through oracle/refimport.py (stub
modules for the uninstalled imageio/lpips/...), runs the reference's own
``render`` / ``render_rays`` / ``raw2outputs`` / ``sample_pdf*`` on small seeded inputs on CPU
(torch %s) and stores inputs + every intermediate + outputs as .npz next to this script.

Determinism: the reference's ``pytest=True`` hook replaces every random draw with
``np.random.seed(0); np.random.rand(...)`` (run_plnerf.py:699-703, run_nerf_helpers.py:255-264,
383-392, run_plnerf.py:572-576), so t_rand / u / noise are all the head of the same numpy stream;
we store them explicitly so the CUDA path and the oracle can be fed identical draws.
Network parameters come from pl-nerf_b200/synth.py (numpy RandomState), not from torch's RNG.
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

CASES = {
    # name: dict(config)
    "lego_linear_mid": dict(rays="lego", n=24, Ns=64, Ni=128, mode="linear", color_mode="midpoint",
                            use_viewdirs=True, white_bkgd=True, lindisp=False, raw_noise_std=0.0,
                            near=2.0, far=6.0, ndc=False, constant_init=False, seeds=(11, 12)),
    "lego_constant": dict(rays="lego", n=24, Ns=64, Ni=128, mode="linear", color_mode="midpoint",
                          use_viewdirs=True, white_bkgd=True, lindisp=False, raw_noise_std=0.0,
                          near=2.0, far=6.0, ndc=False, constant_init=True, seeds=(11, 12)),
    "lego_left_noise_lindisp": dict(rays="lego", n=16, Ns=128, Ni=64, mode="linear", color_mode="left",
                                    use_viewdirs=False, white_bkgd=False, lindisp=True, raw_noise_std=1.0,
                                    near=2.0, far=6.0, ndc=False, constant_init=False, seeds=(21, 22)),
    "llff_ndc_constant": dict(rays="llff", n=16, Ns=64, Ni=64, mode="constant", color_mode="midpoint",
                              use_viewdirs=True, white_bkgd=False, lindisp=False, raw_noise_std=1.0,
                              near=0.0, far=1.0, ndc=True, constant_init=False, seeds=(31, 32)),
    "llff_ndc_linear": dict(rays="llff", n=16, Ns=128, Ni=64, mode="linear", color_mode="midpoint",
                            use_viewdirs=True, white_bkgd=False, lindisp=False, raw_noise_std=0.0,
                            near=0.0, far=1.0, ndc=True, constant_init=False, seeds=(31, 32)),
    "coarse_only": dict(rays="lego", n=32, Ns=64, Ni=0, mode="linear", color_mode="midpoint",
                        use_viewdirs=False, white_bkgd=True, lindisp=False, raw_noise_std=0.0,
                        near=2.0, far=6.0, ndc=False, constant_init=False, seeds=(41, 41)),
}


def net_kwargs(cfg):
    if cfg["use_viewdirs"]:
        return dict(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5 if cfg["Ni"] > 0 else 4,
                    skips=(4,), use_viewdirs=True)
    return dict(D=8, W=256, input_ch=63, input_ch_views=0, output_ch=5 if cfg["Ni"] > 0 else 4,
                skips=(4,), use_viewdirs=False)


def build_ref_net(H, params, kw):
    net = H.NeRF(D=kw["D"], W=kw["W"], input_ch=kw["input_ch"], input_ch_views=kw["input_ch_views"],
                 output_ch=kw["output_ch"], skips=list(kw["skips"]), use_viewdirs=kw["use_viewdirs"])
    sd = {k: torch.from_numpy(v.copy()) for k, v in params.items()}
    net.load_state_dict(sd)
    return net


def run_case(H, R, name, cfg):
    kw = net_kwargs(cfg)
    pc = synth.nerf_params(cfg["seeds"][0], **kw)
    pf = synth.nerf_params(cfg["seeds"][1], **kw)
    net_c, net_f = build_ref_net(H, pc, kw), build_ref_net(H, pf, kw)
    if cfg["rays"] == "lego":
        ro, rd, K, (Hh, Ww, focal) = synth.lego_rays(cfg["n"], seed=5)
    else:
        ro, rd, K, (Hh, Ww, focal) = synth.llff_rays(cfg["n"], seed=5)
    n, Ns, Ni = cfg["n"], cfg["Ns"], cfg["Ni"]
    embed_fn, ic = H.get_embedder(10, 0)
    embeddirs_fn, icv = H.get_embedder(4, 0)
    if not cfg["use_viewdirs"]:
        embeddirs_fn = None
    q = lambda p, v, fn: R.run_network(p, v, fn, embed_fn=embed_fn, embeddirs_fn=embeddirs_fn,
                                       netchunk=1024 * 64)
    common = dict(network_query_fn=q, network_fn=net_c, network_fine=net_f if Ni > 0 else None,
                  N_samples=Ns, N_importance=Ni, perturb=1.0, raw_noise_std=cfg["raw_noise_std"],
                  white_bkgd=cfg["white_bkgd"], mode=cfg["mode"], color_mode=cfg["color_mode"],
                  lindisp=cfg["lindisp"], pytest=True, retraw=True, constant_init=cfg["constant_init"])
    rays_t = torch.stack([torch.from_numpy(ro), torch.from_numpy(rd)])
    with torch.no_grad():
        rgb, disp, acc, extras = R.render(Hh, Ww, K, chunk=32768, rays=rays_t, ndc=cfg["ndc"],
                                          near=cfg["near"], far=cfg["far"],
                                          use_viewdirs=cfg["use_viewdirs"], **common)
    out = {"rays_o": ro, "rays_d": rd, "K": K.astype(np.float32),
           "hwf": np.array([Hh, Ww, focal], np.float64),
           "rgb_map": rgb.numpy(), "disp_map": disp.numpy(), "acc_map": acc.numpy()}
    for k, v in extras.items():
        out[k] = v.numpy()

    # ---- the explicit random draws the pytest hook produced
    def head(shape):
        np.random.seed(0)
        return np.random.rand(*shape)
    out["t_rand"] = head((n, Ns)).astype(np.float32)           # torch.Tensor(float64) -> float32
    if Ni > 0:
        out["u"] = head((n, Ni)).astype(np.float32)
    std = cfg["raw_noise_std"]
    if std > 0:
        out["noise0"] = torch.Tensor(head((n, Ns)) * std).numpy()
        if Ni > 0:
            out["noise1"] = torch.Tensor(head((n, Ns + Ni)) * std).numpy()

    # ---- intermediates: replay render_rays' own steps with the reference's own functions
    with torch.no_grad():
        ro_t, rd_t = torch.from_numpy(ro), torch.from_numpy(rd)
        viewdirs = None
        if cfg["use_viewdirs"]:
            viewdirs = rd_t / torch.norm(rd_t, dim=-1, keepdim=True)
        if cfg["ndc"]:
            ro_t, rd_t = H.ndc_rays(Hh, Ww, K[0][0], 1., ro_t, rd_t)
        near = cfg["near"] * torch.ones_like(rd_t[..., :1])
        far = cfg["far"] * torch.ones_like(rd_t[..., :1])
        cols = [ro_t, rd_t, near, far] + ([viewdirs] if viewdirs is not None else [])
        ray_batch = torch.cat(cols, -1).float()
        out["ray_batch"] = ray_batch.numpy()
        t_vals = torch.linspace(0., 1., steps=Ns)
        if not cfg["lindisp"]:
            z = near * (1. - t_vals) + far * t_vals
        else:
            z = 1. / (1. / near * (1. - t_vals) + 1. / far * t_vals)
        z = z.expand([n, Ns])
        mids = .5 * (z[..., 1:] + z[..., :-1])
        upper = torch.cat([mids, z[..., -1:]], -1)
        lower = torch.cat([z[..., :1], mids], -1)
        z = lower + (upper - lower) * torch.from_numpy(out["t_rand"])
        out["z_vals0"] = z.numpy()
        pts = ro_t[..., None, :] + rd_t[..., None, :] * z[..., :, None]
        out["pts0"] = pts.numpy()
        # PE golden for the first ray's points + all viewdirs
        out["embed_pts0"] = embed_fn(pts[0]).numpy()
        if viewdirs is not None:
            out["embed_dirs"] = H.get_embedder(4, 0)[0](viewdirs).numpy()
        mode = "constant" if cfg["constant_init"] else cfg["mode"]
        raw0 = q(pts, viewdirs, net_c)
        out["raw0"] = raw0.numpy()
        r0 = R.raw2outputs(raw0, z, near, far, rd_t, mode, cfg["color_mode"], cfg["raw_noise_std"],
                           pytest=True, white_bkgd=cfg["white_bkgd"])
        rgb0, disp0, acc0, w0, depth0, tau0, T0 = r0
        out["weights0"] = w0.numpy()
        if tau0 is not None:
            out["tau0"], out["T0"] = tau0.numpy(), T0.numpy()
        if Ni > 0:
            u = torch.from_numpy(out["u"])
            if mode == "linear":
                zs, _, _, _ = H.sample_pdf_reformulation(z, w0, tau0, T0, near, far, Ni, det=False, pytest=True)
                cdf = torch.cat([torch.zeros_like(w0[..., :1]), torch.cumsum(w0, -1)], -1)
                cdf[:, -1] = 1.0
            else:
                z_mid = .5 * (z[..., 1:] + z[..., :-1])
                zs = H.sample_pdf(z_mid, w0[..., 1:-1], Ni, det=False, pytest=True)
                ww = w0[..., 1:-1] + 1e-5
                pdf = ww / torch.sum(ww, -1, keepdim=True)
                cdf = torch.cat([torch.zeros_like(pdf[..., :1]), torch.cumsum(pdf, -1)], -1)
            out["cdf0"] = cdf.numpy()
            out["inds"] = torch.searchsorted(cdf, u.contiguous(), right=True).numpy()
            out["z_samples_raw"] = zs.numpy()
            zs = torch.clamp(zs, near, far)
            z1, _ = torch.sort(torch.cat([z, zs], -1), -1)
            out["z_vals"] = z1.numpy()
            # cross-check the replay against the end-to-end render
            np.testing.assert_array_equal(out["rgb0"], rgb0.numpy())
            np.testing.assert_array_equal(out["z_std"], torch.std(zs, dim=-1, unbiased=False).numpy())
            pts1 = ro_t[..., None, :] + rd_t[..., None, :] * z1[..., :, None]
            raw1 = q(pts1, viewdirs, net_f)
            np.testing.assert_array_equal(out["raw"], raw1.numpy())
            r1 = R.raw2outputs(raw1, z1, near, far, rd_t, mode, cfg["color_mode"], cfg["raw_noise_std"],
                               pytest=True, white_bkgd=cfg["white_bkgd"])
            np.testing.assert_array_equal(out["rgb_map"], r1[0].numpy())
            out["weights1"] = r1[3].numpy()
        else:
            np.testing.assert_array_equal(out["rgb_map"], rgb0.numpy())
    # ---- gradient golden: d(loss)/d(raw0) through the reference's own raw2outputs (torch autograd),
    # loss = <rgb,Wr> + <depth,Wd> + <acc,Wa> + <disp,Wp> with seeded upstream weights
    rs = np.random.RandomState(77)
    up = {"up_rgb": rs.randn(n, 3).astype(np.float32), "up_depth": rs.randn(n).astype(np.float32),
          "up_acc": rs.randn(n).astype(np.float32), "up_disp": rs.randn(n).astype(np.float32)}
    raw_req = torch.from_numpy(out["raw0"].copy()).requires_grad_(True)
    mode_g = "constant" if cfg["constant_init"] else cfg["mode"]
    rb = torch.from_numpy(out["ray_batch"])
    rg = R.raw2outputs(raw_req, torch.from_numpy(out["z_vals0"]), rb[:, 6:7], rb[:, 7:8], rb[:, 3:6], mode_g,
                       cfg["color_mode"], cfg["raw_noise_std"], pytest=True, white_bkgd=cfg["white_bkgd"])
    loss = (rg[0] * torch.from_numpy(up["up_rgb"])).sum() + (rg[4] * torch.from_numpy(up["up_depth"])).sum() \
        + (rg[2] * torch.from_numpy(up["up_acc"])).sum() + (rg[1] * torch.from_numpy(up["up_disp"])).sum()
    loss.backward()
    out.update(up)
    out["g_raw0"] = raw_req.grad.numpy()
    # ---- training-loss gradient golden (viewdirs + fine cases): the reference's own loss
    # img2mse(rgb, target) + img2mse(rgb0, target) (run_plnerf.py:1290-1297) back-propagated through the
    # unmodified reference; stored per parameter as L2 norm + first 256 entries (full tensors are 4.8 MB)
    if cfg["use_viewdirs"] and Ni > 0:
        tgt = torch.from_numpy(np.random.RandomState(99).rand(n, 3).astype(np.float32))
        for p_ in list(net_c.parameters()) + list(net_f.parameters()):
            p_.grad = None
        rgb_t, disp_t, acc_t, ex_t = R.render(Hh, Ww, K, chunk=32768, rays=rays_t, ndc=cfg["ndc"], near=cfg["near"],
                                               far=cfg["far"], use_viewdirs=True, **common)
        loss_t = H.img2mse(rgb_t, tgt) + H.img2mse(ex_t["rgb0"], tgt)
        loss_t.backward()
        out["train_target"] = tgt.numpy()
        out["train_loss"] = np.float32(loss_t.item())
        for tag, net in (("c", net_c), ("f", net_f)):
            for k_, p_ in net.named_parameters():
                g_ = p_.grad.detach().numpy().ravel()
                out[f"gnorm_{tag}.{k_}"] = np.float32(np.linalg.norm(g_))
                out[f"ghead_{tag}.{k_}"] = g_[:256].copy()
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"{name}: wrote {path} ({os.path.getsize(path) / 1024:.0f} KiB), keys={sorted(out)}")


def main():
    H, R = refimport.load()
    torch.manual_seed(0)
    only = sys.argv[1:]
    for name, cfg in CASES.items():
        if only and name not in only:
            continue
        run_case(H, R, name, cfg)


if __name__ == "__main__":
    main()
