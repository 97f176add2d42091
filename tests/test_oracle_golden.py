"""Pin the numpy oracle against golden vectors produced by the unmodified reference
(tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

import plnerf_oracle as O
from util import CASES, case_params, load_golden, max_rel, oracle_net_kw

ALL = list(CASES)
FINE = [c for c in ALL if CASES[c]["Ni"] > 0]


@pytest.mark.parametrize("name", ALL)
def test_embed_and_pack_rays(name):
    g = load_golden(name)
    cfg = CASES[name]
    Hh, Ww, focal = g["hwf"]
    rays = O.pack_rays(int(Hh), int(Ww), g["K"], g["rays_o"], g["rays_d"], cfg["near"], cfg["far"],
                       cfg["use_viewdirs"], cfg["ndc"])
    assert max_rel(rays, g["ray_batch"], 1e-2) < 2e-6
    assert max_rel(O.embed(g["pts0"][0], 10), g["embed_pts0"], 1.0) < 2e-6
    if "embed_dirs" in g:
        assert max_rel(O.embed(g["ray_batch"][:, -3:], 4), g["embed_dirs"], 1.0) < 1e-6


@pytest.mark.parametrize("name", ALL)
def test_stratified_z_bit_exact(name):
    g = load_golden(name)
    cfg = CASES[name]
    rb = g["ray_batch"]
    z = O.stratified_z(rb[:, 6:7], rb[:, 7:8], cfg["Ns"], g["t_rand"], cfg["lindisp"])
    np.testing.assert_array_equal(z, g["z_vals0"])


@pytest.mark.parametrize("name", ALL)
def test_mlp_forward(name):
    g = load_golden(name)
    cfg, kw, pc, pf = case_params(name)
    rb = g["ray_batch"]
    vd = rb[:, -3:] if cfg["use_viewdirs"] else None
    raw0 = O.run_network(g["pts0"], vd, pc, **oracle_net_kw(kw))
    # raw spans O(1..100); compare against its own scale
    assert max_rel(raw0[..., :4], g["raw0"][..., :4], 1.0) < 2e-5


@pytest.mark.parametrize("name", ALL)
def test_quadrature_coarse(name):
    g = load_golden(name)
    cfg = CASES[name]
    rb = g["ray_batch"]
    mode = "constant" if cfg["constant_init"] else cfg["mode"]
    rgb, disp, acc, w, depth, tau, T = O.raw2outputs(
        g["raw0"], g["z_vals0"], rb[:, 6:7], rb[:, 7:8], rb[:, 3:6], mode, cfg["color_mode"],
        g.get("noise0", 0.0), cfg["white_bkgd"])
    # w = (1-e)T cancels: 1 ulp of e is ~6e-8 absolute, so compare on an absolute floor
    assert max_rel(w, g["weights0"], 1e-2) < 2e-5
    sfx = "0" if cfg["Ni"] > 0 else "_map"
    rk = "rgb0" if cfg["Ni"] > 0 else "rgb_map"
    assert max_rel(rgb, g[rk]) < 1e-5
    assert max_rel(depth, g["depth" + sfx]) < 1e-5
    assert max_rel(acc, g["acc" + sfx]) < 1e-5
    assert max_rel(disp, g["disp" + sfx]) < 1e-5
    if mode == "linear":
        assert max_rel(tau, g["tau0"], 1e-3) < 1e-6
        assert max_rel(T, g["T0"], 1e-6) < 2e-5


@pytest.mark.parametrize("name", FINE)
def test_sampler_indices_bit_exact_and_samples(name):
    """Given the reference's own coarse outputs, cdf and inds must be bit-identical."""
    g = load_golden(name)
    cfg = CASES[name]
    rb = g["ray_batch"]
    mode = "constant" if cfg["constant_init"] else cfg["mode"]
    if mode == "linear":
        zs, inds = O.sample_pdf_reformulation(g["z_vals0"], g["weights0"], g["tau0"], g["T0"],
                                              rb[:, 6:7], rb[:, 7:8], g["u"])
    else:
        z = g["z_vals0"]
        zs, inds = O.sample_pdf(0.5 * (z[..., 1:] + z[..., :-1]), g["weights0"][..., 1:-1], g["u"])
    np.testing.assert_array_equal(inds, g["inds"])
    assert max_rel(zs, g["z_samples_raw"], 1e-2) < 5e-6


@pytest.mark.parametrize("name", ALL)
def test_render_end_to_end(name):
    g = load_golden(name)
    cfg, kw, pc, pf = case_params(name)
    Hh, Ww, focal = g["hwf"]
    out = O.render(int(Hh), int(Ww), g["K"], g["rays_o"], g["rays_d"], chunk=1024 * 32, ndc=cfg["ndc"],
                   near=cfg["near"], far=cfg["far"], use_viewdirs=cfg["use_viewdirs"],
                   t_rand=g["t_rand"], u=g.get("u"), noise0=g.get("noise0"), noise1=g.get("noise1"),
                   params_coarse=pc, params_fine=pf if cfg["Ni"] > 0 else None, N_samples=cfg["Ns"],
                   mode=cfg["mode"], color_mode=cfg["color_mode"], N_importance=cfg["Ni"],
                   lindisp=cfg["lindisp"], white_bkgd=cfg["white_bkgd"],
                   constant_init=cfg["constant_init"], retraw=True, net_kw=oracle_net_kw(kw))
    tol = 1e-4
    for k in ["rgb_map", "depth_map", "acc_map", "disp_map"]:
        assert max_rel(out[k], g[k]) < tol, k
    if cfg["Ni"] > 0:
        for k in ["rgb0", "depth0", "acc0", "disp0"]:
            assert max_rel(out[k], g[k]) < tol, k
        assert max_rel(out["z_std"], g["z_std"], 1e-3) < 1e-3
        # merged depths: identical unless a coarse-MLP ulp moved a sample across a knot
        frac_diff = np.mean(np.abs(out["z_vals"] - g["z_vals"]) > 1e-4)
        assert frac_diff < 0.01
