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
