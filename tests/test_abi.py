"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads and exports every
symbol include/plnerf_b200.h declares; host-only entry points behave (no GPU compute here)."""
import ctypes as C
import os
import re

import pytest

import plnerf_b200
from plnerf_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    plnerf_b200.build()
    return L.lib()


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "plnerf_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(plnerf_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(lib):
    syms = declared_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/plnerf_b200.h but not exported"
    assert set(syms) == set(L.PUBLIC_SYMBOLS)


def test_abi_version_and_packed_size(lib):
    assert lib.plnerf_abi_version() == L.ABI_VERSION == 2
    d = L.NetDesc()
    d.D, d.W, d.input_ch, d.input_ch_views, d.output_ch, d.use_viewdirs, d.n_skips = 8, 256, 63, 27, 5, 1, 1
    d.skips[0] = 4
    # bf16 stream: (4 + 7*16 + 4 + 16) K-steps * 2 halves * 4 KiB + views 16 * 4 KiB, + fp32 tail
    n_fast = lib.plnerf_packed_bytes(C.byref(d), L.PREC_BF16)
    n_x3 = lib.plnerf_packed_bytes(C.byref(d), L.PREC_BF16X3)
    ks = (4 + 7 * 16 + 4) * 2 + 16 * 2 + 16
    tail = n_fast - ks * 4096
    assert 0 < tail < 64 * 1024
    assert n_x3 == 2 * ks * 4096 + tail


def test_unsupported_shapes_fail_loudly(lib):
    d = L.NetDesc()
    d.D, d.W, d.input_ch, d.input_ch_views, d.output_ch, d.use_viewdirs, d.n_skips = 8, 128, 63, 27, 5, 1, 0
    assert lib.plnerf_packed_bytes(C.byref(d), L.PREC_BF16) == 0
    assert b"W=256" in lib.plnerf_last_error()


def test_bad_arguments_return_codes(lib):
    rc = lib.plnerf_encode(None, 4, 10, None, None)
    assert rc == -1 and b"null" in lib.plnerf_last_error()
    rc = lib.plnerf_stratified_z(None, 0, 4, 64, 0, 1, None, 0, 0, None, None)
    assert rc == -1  # stride < 8
    # pixel-subset ray packing: null pixel list / pose with n > 0, bad image size, row stride too small for a viewdir
    import ctypes
    buf = (ctypes.c_float * 64)()
    pixel_rays = lambda H, W, c2w, ld, pix, n, use_viewdirs, out, stride: lib.plnerf_pack_pixel_rays(
        H, W, 50.0, 50.0, 4.0, 4.0, c2w, ld, pix, n, 0, -1.0, -1.0, 1.0, 2.0, 6.0, use_viewdirs, out, stride, None)
    assert pixel_rays(8, 8, None, 4, None, 4, 1, buf, 11) == -1 and b"null" in lib.plnerf_last_error()
    assert pixel_rays(0, 8, buf, 4, buf, 4, 1, buf, 11) == -1
    assert pixel_rays(8, 8, buf, 3, buf, 4, 1, buf, 11) == -1          # c2w_ld < 4
    assert pixel_rays(8, 8, buf, 4, buf, 4, 1, buf, 8) == -1 and b"stride" in lib.plnerf_last_error()
    assert pixel_rays(8, 8, None, 4, None, 0, 1, None, 11) == 0        # empty batch: nothing to do, nothing launched


def test_reference_module_surface_mesh_and_train():
    """The f-2 / f-3 mirrors keep the reference's names and argument order."""
    import inspect
    import plnerf_b200.nerf_extract_mesh as M
    import plnerf_b200.train as T
    assert list(inspect.signature(M.extract_fields).parameters)[:5] == ["bound_min", "bound_max", "resolution",
                                                                        "query_func", "model"]
    assert list(inspect.signature(M.extract_iso_level).parameters) == ["density", "threshold"]
    assert inspect.signature(M.extract_iso_level).parameters["threshold"].default == 25
    # TrainStep takes the reference's argparse names (run_plnerf.py:config_parser) for everything it consumes
    for nme in ("N_rand", "chunk", "lrate", "coarse_lrate", "lrate_decay", "precrop_iters", "precrop_frac", "constant_init"):
        assert nme in inspect.signature(T.TrainStep.__init__).parameters, nme


def test_reference_module_surface():
    """Names the reference's run_plnerf.py star-imports / looks up must exist with the same signatures."""
    import inspect
    import plnerf_b200.run_nerf_helpers as H
    import plnerf_b200.run_plnerf as R
    for n in ("img2mse", "mse2psnr", "to8b", "to16b", "Embedder", "get_embedder", "NeRF", "get_rays", "get_rays_np",
              "ndc_rays", "sample_pdf", "sample_pdf_reformulation", "pw_linear_sample_increasing",
              "pw_linear_sample_decreasing"):
        assert hasattr(H, n), n
    ref_args = ["ray_batch", "network_fn", "network_query_fn", "N_samples", "mode", "color_mode", "retraw", "lindisp",
                "perturb", "N_importance", "network_fine", "white_bkgd", "raw_noise_std", "verbose", "pytest",
                "quad_solution_v2", "zero_tol", "epsilon", "farcolorfix", "constant_init"]
    assert list(inspect.signature(R.render_rays).parameters)[:len(ref_args)] == ref_args
    assert list(inspect.signature(R.render).parameters)[:10] == ["H", "W", "K", "chunk", "rays", "c2w", "ndc", "near",
                                                                 "far", "use_viewdirs"]
    net = H.NeRF(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=[4], use_viewdirs=True)
    shapes = {k: tuple(v.shape) for k, v in net.state_dict().items()}
    want = dict(plnerf_b200.synth.nerf_param_shapes(output_ch=5))
    assert shapes == want
    assert sum(v.numel() for v in net.parameters()) == 595844


def test_no_cpu_fallback():
    import torch
    import plnerf_b200.run_nerf_helpers as H
    net = H.NeRF(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=[4], use_viewdirs=True)
    with torch.no_grad(), pytest.raises(RuntimeError, match="no CPU fallback"):
        net(torch.zeros(4, 90))


def test_training_entries_reject_bad_arguments(lib):
    """plnerf_train_rays_mse / plnerf_pack_weights_train validate before they touch the device (no GPU needed)."""
    import ctypes
    cfg = L.RenderCfg()
    cfg.N_samples, cfg.N_importance, cfg.mode, cfg.precision = 64, 64, L.MODE_LINEAR, L.PREC_BF16
    d = L.NetDesc()
    d.D, d.W, d.input_ch, d.input_ch_views, d.output_ch, d.use_viewdirs, d.n_skips = 8, 256, 63, 27, 5, 1, 1
    d.skips[0] = 4
    buf = (ctypes.c_float * 64)()
    g = L.NetGrads()
    call = lambda cpacked, cbwd, grads, sqerr, target, n, stride: lib.plnerf_train_rays_mse(
        C.byref(cfg), C.byref(d), cpacked, cbwd, None, None, None, buf, n, stride, None, None, None, None, target, None, 1.0,
        sqerr, None, grads, None, None, 0, None)
    assert call(None, buf, C.byref(g), buf, buf, 4, 11) == -1 and b"null" in lib.plnerf_last_error()       # no packed weights
    assert call(buf, None, C.byref(g), buf, buf, 4, 11) == -1                                               # no transposed weights
    assert call(buf, buf, None, buf, buf, 4, 11) == -1                                                      # no gradient buffers
    assert call(buf, buf, C.byref(g), None, buf, 4, 11) == -1                                               # no loss accumulators
    assert call(buf, buf, C.byref(g), buf, None, 4, 11) == -1                                               # rays but no target
    assert call(buf, buf, C.byref(g), buf, buf, 4, 8) == -1 and b"stride" in lib.plnerf_last_error()        # no room for a viewdir
    assert call(buf, buf, C.byref(g), buf, buf, 4, 11) < 0 and b"workspace" in lib.plnerf_last_error()      # no workspace
    cfg.precision = L.PREC_BF16X3
    assert call(buf, buf, C.byref(g), buf, buf, 4, 11) < 0 and b"BF16" in lib.plnerf_last_error()           # gradients are bf16 only
    # the one-launch repack: 1 or 2 networks, no null entries
    descs, prms = (C.POINTER(L.NetDesc) * 2)(), (C.POINTER(L.NetParams) * 2)()
    bufs, bwds = (C.c_void_p * 2)(), (C.c_void_p * 2)()
    assert lib.plnerf_pack_weights_train(0, descs, prms, bufs, bwds, None) == -1
    assert lib.plnerf_pack_weights_train(3, descs, prms, bufs, bwds, None) == -1
    assert lib.plnerf_pack_weights_train(1, descs, prms, bufs, bwds, None) == -1 and b"null" in lib.plnerf_last_error()
