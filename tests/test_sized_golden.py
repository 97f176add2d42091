"""Parity at the sizes of BASELINE.json's configs (goldens from the unmodified reference, tests/golden/make_golden_sized.py):
config 1 = one 4096-ray coarse-only batch, config 2 shape = 1024 lego rays at 64 + 128 samples with the bench's
density-boosted nets.  CPU: the oracle against the goldens.  GPU: render() through the C ABI against the goldens, 1e-4
relative in the parity mode (bf16x3), the measured bounds in the bf16 mode.

The fine-pass DEPTH is a conditioned quantity: the inverse-CDF sampler divides by the pdf, so in near-empty space a 1e-7
change of a coarse weight moves a sample by ~1e-5.  `selfcheck_orders` measures that on the reference itself: the same
network evaluated in two fp32 summation orders (hidden units renumbered) agrees to 1.6e-7 of `far` in the coarse depth
and only to 5.9e-6 in the fine depth -- an amplification of ~37x that any implementation inherits.  The fine-depth gate
below is therefore stated as  max(1e-4, SAFETY x amplification x own largest coarse-output error), capped at 1e-3, with the
99th percentile still held to 1e-4.  SAFETY = 4: the amplification is a maximum over rays with a heavy tail and was
measured on 256 rays, the gated cases have up to 1024 (measured on a B200, bf16x3: coarse outputs within 1.2e-6 of the
reference, fine depth max 4e-5 of far, fine disparity max 2.1e-4 / p99 4.5e-5 of its scale).
"""
import numpy as np
import pytest
import torch

import plnerf_oracle as O
from make_golden_sized import SIZED, net_kwargs, pytest_draws, sized_inputs
from util import load_golden, max_rel

FAR = 6.0
SAFETY = 4


def oracle_render(name):
    cfg = SIZED[name]
    ro, rd, K, hwf, pc, pf = sized_inputs(cfg)
    t_rand, u = pytest_draws(cfg["n"], cfg["Ns"], cfg["Ni"])
    kw = net_kwargs(cfg)
    return O.render(hwf[0], hwf[1], K, ro, rd, ndc=False, near=2., far=FAR, use_viewdirs=cfg["use_viewdirs"], t_rand=t_rand,
                    u=u, params_coarse=pc, params_fine=pf if cfg["Ni"] > 0 else None, N_samples=cfg["Ns"], mode="linear",
                    color_mode="midpoint", N_importance=cfg["Ni"], white_bkgd=cfg["white_bkgd"],
                    net_kw=dict(D=8, skips=(4,), input_ch=63, input_ch_views=kw["input_ch_views"],
                                use_viewdirs=cfg["use_viewdirs"]))


def amplification(key="depth_map"):
    """(coarse, fine): how far the reference's two fp32 orders disagree in the coarse depth and in fine output `key`,
    each relative to the output's scale (depth: far; disparity: its maximum)."""
    g = load_golden("selfcheck_orders")
    coarse = np.abs(g["depth0_a"] - g["depth0_b"]).max() / FAR
    scale = FAR if key == "depth_map" else float(np.abs(g[key + "_a"]).max())
    fine = np.abs(g[key + "_a"] - g[key + "_b"]).max() / scale
    return coarse, fine


def coarse_error(r, g):
    """Largest error of the coarse-pass outputs, each relative to its scale: what perturbs the sampler's cdf."""
    errs = {"rgb0": np.abs(r["rgb0"] - g["rgb0"]).max(), "acc0": np.abs(r["acc0"] - g["acc0"]).max(),
            "depth0": np.abs(r["depth0"] - g["depth0"]).max() / FAR,
            "disp0": np.abs(r["disp0"] - g["disp0"]).max() / float(np.abs(g["disp0"]).max())}
    return max(errs.values()), errs


def test_reference_disagrees_with_itself_in_the_fine_depth():
    """Two fp32 evaluation orders of the UNMODIFIED reference: ~1 ulp in every coarse output, 10-100x more in the fine
    depth.  Pins the numbers the fine-depth gates below are derived from."""
    g = load_golden("selfcheck_orders")
    coarse, fine = amplification()
    assert coarse < 5e-7                                  # fp32 rounding only
    assert np.abs(g["rgb0_a"] - g["rgb0_b"]).max() < 1e-6
    assert 2e-6 < fine < 2e-5                             # measured 5.9e-6
    assert fine / coarse > 10                             # the sampler's conditioning, measured ~37x
    c2, fine_disp = amplification("disp_map")
    assert fine_disp / c2 > 10                            # fine disparity = acc / depth: measured ~53x
    # the integer indices themselves are NOT reproducible across fp32 orders of the reference's own coarse pass ...
    mism = np.mean(g["inds_a"] != g["inds_b"])
    assert mism < 1e-3                                    # ... a handful of u's sit within an ulp of a cdf knot
    # ... which is why "indices bit-exact" is gated at the operator level (same weights in -> same indices out)


@pytest.mark.parametrize("name", ["c1_coarse_4096", "c2_lego_1024"])
def test_oracle_vs_sized_golden(name):
    g = load_golden(name)
    r = oracle_render(name)
    for k in ("rgb_map", "acc_map", "disp_map"):
        assert max_rel(r[k], g[k]) < 2e-5, k
    if SIZED[name]["Ni"] > 0:
        for k in ("rgb0", "acc0", "disp0", "depth0"):
            assert max_rel(r[k], g[k]) < 2e-5, k
        coarse, fine = amplification()
        e0 = np.abs(r["depth0"] - g["depth0"]).max() / FAR
        e1 = np.abs(r["depth_map"] - g["depth_map"]).max() / FAR
        assert e1 < max(1e-4, SAFETY * (fine / coarse) * e0)
        assert np.mean(r["inds"].astype(np.int64) != g["inds"].astype(np.int64)) < 1e-3
    else:
        assert max_rel(r["depth_map"], g["depth_map"]) < 2e-5


# ------------------------------------------------------------------------------------------------------------------
def gpu_render(name, precision):
    from plnerf_b200 import run_plnerf as RP
    from plnerf_b200.run_nerf_helpers import NeRF
    cfg = SIZED[name]
    kw = net_kwargs(cfg)
    ro, rd, K, hwf, pc, pf = sized_inputs(cfg)

    def mk(prm):
        net = NeRF(D=8, W=256, input_ch=63, input_ch_views=kw["input_ch_views"], output_ch=kw["output_ch"], skips=[4],
                   use_viewdirs=cfg["use_viewdirs"])
        net.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in prm.items()})
        return net.cuda()
    rays = torch.stack([torch.from_numpy(ro), torch.from_numpy(rd)]).cuda()
    with torch.no_grad():
        rgb, disp, acc, ex = RP.render(hwf[0], hwf[1], K, chunk=1024 * 32, rays=rays, ndc=False, near=2., far=FAR,
                                       use_viewdirs=cfg["use_viewdirs"], network_query_fn=None, network_fn=mk(pc),
                                       network_fine=mk(pf) if cfg["Ni"] > 0 else None, N_samples=cfg["Ns"],
                                       N_importance=cfg["Ni"], perturb=1.0, raw_noise_std=0., white_bkgd=cfg["white_bkgd"],
                                       mode="linear", color_mode="midpoint", pytest=True, precision=precision)
    out = {"rgb_map": rgb, "disp_map": disp, "acc_map": acc, **ex}
    return {k: v.detach().cpu().numpy() for k, v in out.items()}


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["c1_coarse_4096", "c2_lego_1024", "selfcheck_orders"])
def test_render_sized_parity_mode(name):
    """bf16x3 (the parity mode): every output within 1e-4 relative of the unmodified reference at BASELINE sizes; the fine
    depth within the conditioning-aware gate (module docstring)."""
    g = load_golden(name)
    if name == "selfcheck_orders":
        g = {k[:-2]: v for k, v in g.items() if k.endswith("_a")}
    r = gpu_render(name, "bf16x3")
    for k in ("rgb_map", "acc_map"):
        assert max_rel(r[k], g[k]) < 1e-4, k
    if SIZED[name]["Ni"] > 0:
        for k in ("rgb0", "acc0", "disp0", "depth0"):
            assert max_rel(r[k], g[k]) < 1e-4, k
        e0, e0_all = coarse_error(r, g)
        # fine depth, and the fine disparity = acc / depth which inherits its error
        for k, scale in (("depth_map", FAR), ("disp_map", float(np.abs(g["disp_map"]).max()))):
            coarse, fine = amplification(k)
            gate = min(1e-3, max(1e-4, SAFETY * (fine / coarse) * e0))
            e1 = np.abs(r[k] - g[k]) / scale
            print(f"{name} {k}: max {e1.max():.3e} p99 {np.percentile(e1, 99):.3e} gate {gate:.3e} coarse errors {e0_all}")
            assert e1.max() < gate, (k, e1.max(), gate, e0_all)
            assert np.percentile(e1, 99) < 1e-4, k
        assert max_rel(r["z_std"], g["z_std"], 1e-3) < 1e-3
    else:
        assert max_rel(r["disp_map"], g["disp_map"]) < 1e-4
        assert max_rel(r["depth_map"], g["depth_map"]) < 1e-4


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["c1_coarse_4096", "c2_lego_1024"])
def test_render_sized_bf16_mode(name):
    """bf16 (the benchmarked mode: bf16 operands, fp32 accumulate) against the fp32 reference at BASELINE sizes, gated at
    ~2x what was measured on a B200 (rgb 3e-3, coarse depth 2e-3 of far, fine depth 1e-2 of far; PSNR of the image against
    the reference's >= 70 dB)."""
    g = load_golden(name)
    r = gpu_render(name, "bf16")
    assert np.abs(r["rgb_map"] - g["rgb_map"]).max() < 6e-3
    mse = np.mean((r["rgb_map"].astype(np.float64) - g["rgb_map"]) ** 2)
    assert -10 * np.log10(mse) > 70.0
    if SIZED[name]["Ni"] > 0:
        assert np.abs(r["depth0"] - g["depth0"]).max() / FAR < 4e-3
        assert np.abs(r["depth_map"] - g["depth_map"]).max() / FAR < 2e-2
        assert np.percentile(np.abs(r["depth_map"] - g["depth_map"]) / FAR, 99) < 4e-3
    else:
        assert np.abs(r["depth_map"] - g["depth_map"]).max() / FAR < 4e-3
