"""Rarely-used options of the path -- ``farcolorfix`` in raw2outputs (run_plnerf.py:583-587) and non-default
``zero_threshold`` / ``epsilon_`` in sample_pdf_reformulation (run_nerf_helpers.py:364-445) -- against outputs of the
unmodified reference (tests/golden/make_golden_flags.py; inputs = the coarse-pass tensors of lego_linear_mid.npz).
CPU: the oracle.  GPU (-m gpu): the CUDA operators through the C ABI, same tolerances as their default-option tests."""
import numpy as np
import pytest
import torch

import plnerf_oracle as O
from util import load_golden, max_rel

BASE = "lego_linear_mid"


def _inputs():
    g, f = load_golden(BASE), load_golden("flags_lego")
    rb = g["ray_batch"]
    return g, f, rb


def test_oracle_farcolorfix_vs_reference():
    g, f, rb = _inputs()
    rgb, disp, acc, w, depth, tau, T = O.raw2outputs(g["raw0"], g["z_vals0"], rb[:, 6:7], rb[:, 7:8], rb[:, 3:6], "linear",
                                                     "midpoint", 0.0, white_bkgd=True, farcolorfix=True)
    assert max_rel(rgb, f["fcf_rgb_map"]) < 2e-5 and max_rel(depth, f["fcf_depth_map"]) < 2e-5
    assert max_rel(acc, f["fcf_acc_map"]) < 2e-5 and max_rel(disp, f["fcf_disp_map"]) < 2e-5
    assert max_rel(w, f["fcf_weights"], 1e-2) < 2e-5
    np.testing.assert_array_equal(tau, f["fcf_tau"])
    # the option matters on these rays (otherwise this test would pin nothing)
    plain = O.raw2outputs(g["raw0"], g["z_vals0"], rb[:, 6:7], rb[:, 7:8], rb[:, 3:6], "linear", "midpoint", 0.0, white_bkgd=True)
    assert np.abs(plain[0] - rgb).max() > 1e-3


def test_oracle_sampler_thresholds_vs_reference():
    g, f, rb = _inputs()
    zs, inds = O.sample_pdf_reformulation(g["z_vals0"], g["weights0"], g["tau0"], g["T0"], rb[:, 6:7], rb[:, 7:8], g["u"],
                                          zero_threshold=float(f["zero_threshold"]), epsilon_=float(f["epsilon"]))
    np.testing.assert_array_equal(inds, g["inds"])                  # the thresholds do not touch the search
    assert max_rel(zs, f["flags_z_samples"], 1e-2) < 5e-6
    assert np.mean(zs != g["z_samples_raw"]) > 0.01                 # but they move the samples


# ------------------------------------------------------------------------------------------------ GPU
def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def host(t):
    return t.detach().cpu().numpy()


@pytest.mark.gpu
def test_farcolorfix_forward_and_backward_vs_reference():
    from plnerf_b200 import ops as P
    g, f, rb = _inputs()
    rgb, disp, acc, w, depth, tau, T = P.raw2outputs(dev(g["raw0"]), dev(g["z_vals0"]), dev(rb), "linear", "midpoint",
                                                     white_bkgd=True, farcolorfix=True)
    assert max_rel(host(rgb), f["fcf_rgb_map"]) < 2e-5 and max_rel(host(depth), f["fcf_depth_map"]) < 2e-5
    assert max_rel(host(acc), f["fcf_acc_map"]) < 2e-5 and max_rel(host(disp), f["fcf_disp_map"]) < 2e-5
    assert max_rel(host(w), f["fcf_weights"], 1e-2) < 2e-5
    np.testing.assert_array_equal(host(tau), f["fcf_tau"])
    graw = host(P.raw2outputs_bwd(dev(g["raw0"]), dev(g["z_vals0"]), dev(rb), "linear", "midpoint", g_rgb=dev(g["up_rgb"]),
                                  g_depth=dev(g["up_depth"]), g_acc=dev(g["up_acc"]), g_disp=dev(g["up_disp"]),
                                  white_bkgd=True, farcolorfix=True))
    ref = f["fcf_g_raw"]
    for c in range(4):
        scale = np.abs(ref[..., c]).max() + 1e-12
        assert np.abs(graw[..., c] - ref[..., c]).max() / scale < 1e-4, c


@pytest.mark.gpu
def test_sampler_thresholds_vs_reference():
    from plnerf_b200 import ops as P
    g, f, rb = _inputs()
    zs, inds = P.sample_pdf_pl(dev(g["z_vals0"]), dev(g["weights0"]), dev(g["tau0"]), dev(g["T0"]), dev(rb), g["u"].shape[1],
                               u=dev(g["u"]), zero_tol=float(f["zero_threshold"]), epsilon=float(f["epsilon"]),
                               return_inds=True)
    np.testing.assert_array_equal(host(inds), g["inds"])
    assert max_rel(host(zs), f["flags_z_samples"], 1e-2) < 1e-5


# ------------------------------------------------------------------------------------------------
# perturb = 0: no stratified jitter, det=True sampler (u = linspace), constant mode (the reference raises in linear mode)
# ------------------------------------------------------------------------------------------------
def _det_case():
    from make_golden_flags import DET, det_net_kwargs
    from util import synth
    kw = det_net_kwargs()
    return DET, kw, synth.nerf_params(DET["seeds"][0], **kw), synth.nerf_params(DET["seeds"][1], **kw), load_golden("det_constant")


def test_oracle_deterministic_render_vs_reference():
    DET, kw, pc, pf, g = _det_case()
    Hh, Ww, focal = g["hwf"]
    n, Ni = DET["n"], DET["Ni"]
    u = np.broadcast_to(torch.linspace(0., 1., steps=Ni).numpy(), (n, Ni)).copy()      # run_nerf_helpers.py:248-250
    ref = O.render(int(Hh), int(Ww), g["K"], g["rays_o"], g["rays_d"], ndc=False, near=2., far=6., use_viewdirs=True,
                   t_rand=None, u=u, params_coarse=pc, params_fine=pf, N_samples=DET["Ns"], mode="constant",
                   color_mode="midpoint", N_importance=Ni, white_bkgd=True,
                   net_kw=dict(D=8, skips=(4,), input_ch=63, input_ch_views=27, use_viewdirs=True))
    for k in ("rgb_map", "rgb0", "acc_map", "acc0", "depth_map", "depth0", "disp_map", "disp0", "z_std"):
        assert max_rel(ref[k], g[k]) < 2e-5, k


@pytest.mark.gpu
def test_deterministic_render_vs_reference():
    """render(perturb=0) through the public API (bf16x3): the device makes its own linspace u; maps within 1e-4."""
    from plnerf_b200 import run_plnerf as RP
    from plnerf_b200.run_nerf_helpers import NeRF
    DET, kw, pc, pf, g = _det_case()

    def mk(p):
        net = NeRF(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=[4], use_viewdirs=True)
        net.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in p.items()})
        return net.cuda()
    Hh, Ww, focal = g["hwf"]
    rays = torch.stack([dev(g["rays_o"]), dev(g["rays_d"])])
    with torch.no_grad():
        rgb, disp, acc, extras = RP.render(int(Hh), int(Ww), g["K"], chunk=1024 * 32, rays=rays, ndc=False, near=2., far=6.,
                                           use_viewdirs=True, network_query_fn=None, network_fn=mk(pc), network_fine=mk(pf),
                                           N_samples=DET["Ns"], N_importance=DET["Ni"], perturb=0., raw_noise_std=0.,
                                           white_bkgd=True, mode="constant", color_mode="midpoint", precision="bf16x3")
    assert max_rel(host(rgb), g["rgb_map"]) < 1e-4 and max_rel(host(extras["rgb0"]), g["rgb0"]) < 1e-4
    assert max_rel(host(acc), g["acc_map"]) < 1e-4 and max_rel(host(extras["depth0"]), g["depth0"]) < 1e-4
    assert max_rel(host(extras["depth_map"]), g["depth_map"]) < 1e-3      # sampler conditioning, see smoke()
