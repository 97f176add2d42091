"""The drop-in boundary against the REAL reference module: the unmodified run_plnerf.py / run_nerf_helpers.py
(from /root/reference in the build container, from the byte-identical staged copies under oracle/_ref/ on the GPU box)
with this package's callables rebound into it by ``plnerf_b200.run_plnerf.install``.

CPU: every reference callable on the path and its mirror agree in ``inspect.signature`` (same parameters, order and
defaults; the mirrors only append keyword extras), and the star-import surface of run_nerf_helpers is complete.
GPU: ``R.create_nerf(args)`` builds the networks (our NeRF class, the reference's optimizers), ``R.render(**render_kwargs_test)``
reproduces the golden of the same unmodified module run on CPU, five iterations of the reference's training-loop body
(run_plnerf.py:1283-1315) run through its own ``render`` / ``img2mse`` / Adam objects, and the launcher's
``torch.set_default_tensor_type('torch.cuda.FloatTensor')`` (:1582) is in force throughout."""
import inspect
import os
import sys
import types
from argparse import Namespace

import numpy as np
import pytest
import torch

import refimport
from make_golden_sized import SIZED, sized_inputs
from util import load_golden, max_rel

pytestmark = pytest.mark.skipif(not refimport.available(), reason="reference modules not reachable (run oracle/stage_ref.py)")


def _mods():
    import plnerf_b200
    from plnerf_b200 import run_plnerf as RP, run_nerf_helpers as HP
    H, R = refimport.load()
    return H, R, HP, RP


def _params(fn):
    return [(p.name, p.default, p.kind) for p in inspect.signature(fn).parameters.values()]


@pytest.mark.parametrize("name", ["render", "batchify_rays", "render_rays", "raw2outputs", "run_network", "batchify",
                                  "compute_weights", "compute_weights_piecewise_linear"])
def test_run_plnerf_signatures(name):
    H, R, HP, RP = _mods()
    ref, ours = _params(getattr(R, name)), _params(getattr(RP, name))
    assert ours[:len(ref)] == ref, (name, ref, ours)
    for extra in ours[len(ref):]:      # appended knobs must be optional
        assert extra[1] is not inspect.Parameter.empty or extra[2] == inspect.Parameter.VAR_KEYWORD, (name, extra)


@pytest.mark.parametrize("name", ["get_embedder", "sample_pdf", "sample_pdf_reformulation", "sample_pdf_return_u",
                                  "sample_pdf_reformulation_return_u", "pw_linear_sample_increasing",
                                  "pw_linear_sample_decreasing", "get_rays", "get_rays_np", "ndc_rays", "compute_rmse"])
def test_helper_signatures(name):
    H, R, HP, RP = _mods()
    ref, ours = _params(getattr(H, name)), _params(getattr(HP, name))
    assert ours[:len(ref)] == ref, (name, ref, ours)
    for extra in ours[len(ref):]:
        assert extra[1] is not inspect.Parameter.empty, (name, extra)


def test_nerf_class_surface():
    H, R, HP, RP = _mods()
    assert _params(HP.NeRF.__init__) == _params(H.NeRF.__init__)
    kw = dict(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=[4], use_viewdirs=True)
    a, b = H.NeRF(**kw), HP.NeRF(**kw)
    assert [(k, tuple(v.shape)) for k, v in a.state_dict().items()] == [(k, tuple(v.shape)) for k, v in b.state_dict().items()]
    assert [k for k, _ in a.named_parameters()] == [k for k, _ in b.named_parameters()]     # order feeds Adam (:431-447)
    for attr in ("D", "W", "input_ch", "input_ch_views", "skips", "use_viewdirs"):
        assert getattr(a, attr) == getattr(b, attr)
    for m in ("add", "has", "get", "as_dict", "reset", "print"):
        assert _params(getattr(HP.MeanTracker, m)) == _params(getattr(H.MeanTracker, m))
    t1, t2 = H.MeanTracker(), HP.MeanTracker()
    for t in (t1, t2):
        t.add({"a": 1.0, "b": 4.0}); t.add({"a": 3.0}, weight=2.)
    assert t1.as_dict() == t2.as_dict() and t1.total_weight == t2.total_weight


def test_star_import_surface_complete():
    """run_plnerf.py does `from run_nerf_helpers import *`: every public name the reference module defines itself must
    exist in the mirror (INTEGRATION.md's "shadow the module on sys.path" route)."""
    H, R, HP, RP = _mods()
    public = [n for n, v in vars(H).items() if not n.startswith("_") and not isinstance(v, types.ModuleType)
              and getattr(v, "__module__", H.__name__) in (H.__name__, None)]
    missing = [n for n in public if not hasattr(HP, n)]
    assert not missing, missing


# ---------------------------------------------------------------------------------------------------------------------
def _args(tmp):
    os.makedirs(os.path.join(tmp, "exp"), exist_ok=True)
    return Namespace(multires=10, i_embed=0, use_viewdirs=True, multires_views=4, N_importance=128, N_samples=64, netdepth=8,
                     netwidth=256, netdepth_fine=8, netwidth_fine=256, netchunk=1024 * 64, lrate=5e-4, coarse_lrate=5e-4,
                     ft_path=None, ckpt_dir=tmp, expname="exp", no_reload=True, perturb=1., white_bkgd=True, raw_noise_std=0.,
                     mode="linear", color_mode="midpoint", dataset="blender", no_ndc=False, lindisp=False, lrate_decay=500,
                     chunk=1024 * 32, constant_init=0)


@pytest.mark.gpu
def test_install_into_the_real_module(tmp_path):
    H, R, HP, RP = _mods()
    from plnerf_b200 import ops
    cfg = SIZED["c2_lego_1024"]
    ro, rd, K, (Hh, Ww, focal), pc, pf = sized_inputs(cfg)
    g = load_golden("c2_lego_1024")
    R.device = torch.device("cuda")
    torch.set_default_tensor_type('torch.cuda.FloatTensor')          # what the reference's __main__ does (:1582)
    saved = RP.install(R)
    try:
        assert R.render is RP.render and R.render_rays is RP.render_rays and R.NeRF is HP.NeRF
        args = _args(str(tmp_path))
        render_kwargs_train, render_kwargs_test, start, grad_vars, optimizer, optimizer_coarse = R.create_nerf(args)
        net_c, net_f = render_kwargs_train["network_fn"], render_kwargs_train["network_fine"]
        assert isinstance(net_c, HP.NeRF) and next(net_c.parameters()).is_cuda and start == 0
        net_c.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in pc.items()})
        net_f.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in pf.items()})
        bds = {"near": 2., "far": 6.}                                 # train() does the same (:1155-1160)
        render_kwargs_train.update(bds); render_kwargs_test.update(bds)
        rays = torch.stack([torch.from_numpy(ro), torch.from_numpy(rd)]).cuda()
        # ---- eval path: render(**render_kwargs_test) as render_images_with_metrics calls it (:325), parity mode
        ops.set_precision("bf16x3")
        try:
            with torch.no_grad():
                rgb, disp, acc, extras = R.render(Hh, Ww, K, chunk=args.chunk, rays=rays, pytest=True, **render_kwargs_test)
        finally:
            ops.set_precision("bf16")
        assert rgb.shape == (cfg["n"], 3) and disp.shape == (cfg["n"],) and rgb.is_cuda
        assert max_rel(rgb.cpu().numpy(), g["rgb_map"]) < 1e-4
        assert max_rel(acc.cpu().numpy(), g["acc_map"]) < 1e-4
        assert max_rel(extras["rgb0"].cpu().numpy(), g["rgb0"]) < 1e-4
        assert max_rel(extras["depth0"].cpu().numpy(), g["depth0"]) < 1e-4
        assert np.percentile(np.abs(extras["depth_map"].cpu().numpy() - g["depth_map"]) / 6.0, 99) < 1e-4
        for k in ("rgb0", "disp0", "depth0", "acc0", "z_std", "depth_map"):
            assert k in extras
        # ---- five iterations of the reference's loop body (:1283-1315), its own names throughout
        tgt_img = torch.rand(cfg["n"], 3)
        assert tgt_img.is_cuda                                        # the default tensor type is in force
        global_step = start
        losses = []
        before = [p.detach().clone() for p in list(net_c.parameters()) + list(net_f.parameters())]
        for i in range(start + 1, start + 6):
            sel = torch.randperm(cfg["n"])[:256]
            batch_rays, target_s = rays[:, sel], tgt_img[sel]
            rgb, disp, acc, extras = R.render(Hh, Ww, K, chunk=args.chunk, rays=batch_rays, verbose=i < 10, retraw=True,
                                              constant_init=i < args.constant_init, **render_kwargs_train)
            optimizer.zero_grad()
            optimizer_coarse.zero_grad()
            img_loss = R.img2mse(rgb, target_s)
            trans = extras['raw'][..., -1]
            assert trans.shape == (256, 192)
            loss = img_loss
            psnr = R.mse2psnr(img_loss)
            if 'rgb0' in extras:
                img_loss0 = R.img2mse(extras['rgb0'], target_s)
                loss = loss + img_loss0
            loss.backward()
            for p in grad_vars:
                assert p.grad is not None and torch.isfinite(p.grad).all()
            optimizer.step()
            optimizer_coarse.step()
            new_lrate = args.lrate * (0.1 ** (global_step / (args.lrate_decay * 1000)))
            for param_group in optimizer.param_groups:
                param_group['lr'] = new_lrate
            for param_group in optimizer_coarse.param_groups:
                param_group['lr'] = new_lrate
            global_step += 1
            losses.append(float(loss))
            assert np.isfinite(losses[-1]) and np.isfinite(float(psnr))
        after = list(net_c.parameters()) + list(net_f.parameters())
        assert all(not torch.equal(a, b) for a, b in zip(before, after))          # both networks were updated
        # the packed copies follow the optimizer (version counters): a fresh render sees the new weights
        with torch.no_grad():
            rgb2, _, _, _ = R.render(Hh, Ww, K, chunk=args.chunk, rays=rays[:, :64], pytest=True, **render_kwargs_test)
        assert not torch.allclose(rgb2, torch.from_numpy(g["rgb_map"][:64]).cuda(), atol=1e-6)
    finally:
        RP.uninstall(R, saved)
        torch.set_default_tensor_type('torch.FloatTensor')
    assert R.render is not RP.render and R.NeRF is H.NeRF


@pytest.mark.gpu
def test_extract_fields_under_cuda_default_tensor_type():
    """the mesh extractor's launcher sets the CUDA default tensor type (nerf_extract_mesh.py:1213); the pinned
    host grid must still be allocated on the host."""
    from plnerf_b200 import nerf_extract_mesh as EM, synth
    from plnerf_b200.run_nerf_helpers import NeRF
    kw = dict(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=(4,), use_viewdirs=True)
    net = NeRF(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=[4], use_viewdirs=True)
    net.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in synth.nerf_params(1, **kw).items()})
    net = net.cuda()
    ref = EM.extract_fields([-1., -1., -1.], [1., 1., 1.], 48, None, net)
    torch.set_default_tensor_type('torch.cuda.FloatTensor')
    try:
        got = EM.extract_fields([-1., -1., -1.], [1., 1., 1.], 48, None, net)
    finally:
        torch.set_default_tensor_type('torch.FloatTensor')
    assert isinstance(got, np.ndarray) and got.shape == (48, 48, 48)
    np.testing.assert_array_equal(got, ref)
