"""Training parity for BASELINE config 3 (configs/blender_linear.txt shape: N_rand=1024, N_samples=128, N_importance=64),
against the UNMODIFIED reference -- not against this package's own autograd route:

* gradients: TrainStep's flat gradient buffer (direct path: fused forward-with-stash + backward entries) per parameter of
  both networks against the reference's loss.backward() golden (tests/golden/train_c3_1024.npz: norm + first 2048
  entries), every layer gated including the trunk;
* convergence: the same initial weights, the same pixel batches and the same draws for 300 iterations of (a) TrainStep
  (bf16 tensor-core operands) and (b) the reference's own loop body in fp32 PyTorch autograd on the GPU (its render,
  img2mse, two Adam optimizers), target = an image rendered by a teacher network; loss curves and final PSNR compared.
"""
import numpy as np
import pytest
import torch

import refimport
from make_golden_sized import SIZED, net_kwargs, pytest_draws, sized_inputs, train_target
from util import load_golden, synth

pytestmark = pytest.mark.gpu

NET_KW = dict(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=(4,), use_viewdirs=True)


def mk_net(params):
    from plnerf_b200.run_nerf_helpers import NeRF
    net = NeRF(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=[4], use_viewdirs=True)
    net.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in params.items()})
    return net.cuda()


def test_train_step_gradients_vs_reference_golden():
    from plnerf_b200 import train as T
    cfg = SIZED["train_c3_1024"]
    g = load_golden("train_c3_1024")
    ro, rd, K, (H, W, focal), pc, pf = sized_inputs(cfg)
    t_rand, u = pytest_draws(cfg["n"], cfg["Ns"], cfg["Ni"])
    net_c, net_f = mk_net(pc), mk_net(pf)
    kw = dict(network_query_fn=None, network_fn=net_c, network_fine=net_f, N_samples=cfg["Ns"], N_importance=cfg["Ni"],
              perturb=1.0, white_bkgd=True, raw_noise_std=0., mode="linear", color_mode="midpoint", use_viewdirs=True,
              ndc=False, near=2., far=6., t_rand=torch.from_numpy(t_rand).cuda(), u=torch.from_numpy(u).cuda())
    step = T.TrainStep(H, W, K, kw, N_rand=cfg["n"], lrate=5e-4, coarse_lrate=5e-4, lrate_decay=500)
    assert step._direct                                   # the fused forward / backward entries, no autograd graph
    batch_rays = torch.stack([torch.from_numpy(ro), torch.from_numpy(rd)]).cuda()
    out = step.step_rays(batch_rays, torch.from_numpy(train_target(cfg)).cuda(), 1)
    # forward: the loss of the bf16 forward against the fp32 reference's
    assert abs(out["loss"].item() - float(g["train_loss"])) / float(g["train_loss"]) < 5e-3
    flat = step.bucket.flat
    off = 0
    worst = {}
    for tag, net in (("f", net_f), ("c", net_c)):        # the bucket holds the fine network first
        for k, p in net.named_parameters():
            got = flat[off:off + p.numel()].double()
            off += p.numel()
            ref_norm = float(g[f"gnorm_{tag}.{k}"])
            head = torch.from_numpy(g[f"ghead_{tag}.{k}"]).cuda().double()
            a = got[:head.numel()]
            rel = (a - head).norm().item() / head.norm().item()
            cos = torch.dot(a, head).item() / (a.norm().item() * head.norm().item())
            nrm = got.norm().item() / ref_norm
            worst[f"{tag}.{k}"] = (round(nrm, 4), round(rel, 4), round(cos, 5))
            trunk = k.startswith("pts_linears")
            # bf16 operand rounding accumulates along the chained gradient GEMMs (heads -> layer 0).  Measured on a B200
            # against the fp32 reference: norms within 0.9% (trunk) / 1.9% (heads: alpha_linear.bias), slices within 9.7%
            # (fine layer 0; 6% at layer 1, <= 2.6% from layer 3 up, <= 1.9% outside the trunk), cosines >= 0.9953.
            # Gates at ~1.5x that, EVERY layer:
            assert abs(nrm - 1) < (0.02 if trunk else 0.03), (tag, k, worst[f"{tag}.{k}"])
            assert cos > (0.993 if trunk else 0.9998), (tag, k, worst[f"{tag}.{k}"])
            assert rel < (0.14 if trunk else 0.03), (tag, k, worst[f"{tag}.{k}"])
    print("norm ratio / slice rel err / slice cos per parameter:", worst)


def test_train_step_gradients_without_viewdirs_vs_reference_golden():
    """Networks WITHOUT view directions (output_linear head, run_nerf_helpers.py:100-103, :126; output_ch = 5 like
    create_nerf builds them when N_importance > 0): TrainStep's gradients against the unmodified reference's loss.backward()
    (tests/golden/train_noviews_256.npz).  The unused views_linears of such a network gets no gradient (autograd: None, here:
    zeros); the fifth output channel neither."""
    from plnerf_b200 import train as T
    from plnerf_b200.run_nerf_helpers import NeRF
    cfg = SIZED["train_noviews_256"]
    g = load_golden("train_noviews_256")
    ro, rd, K, (H, W, focal), pc, pf = sized_inputs(cfg)
    t_rand, u = pytest_draws(cfg["n"], cfg["Ns"], cfg["Ni"])
    kwn = net_kwargs(cfg)

    def mk(params):
        net = NeRF(D=kwn["D"], W=kwn["W"], input_ch=kwn["input_ch"], input_ch_views=kwn["input_ch_views"],
                   output_ch=kwn["output_ch"], skips=list(kwn["skips"]), use_viewdirs=False)
        net.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in params.items()})
        return net.cuda()
    net_c, net_f = mk(pc), mk(pf)
    kw = dict(network_query_fn=None, network_fn=net_c, network_fine=net_f, N_samples=cfg["Ns"], N_importance=cfg["Ni"],
              perturb=1.0, white_bkgd=True, raw_noise_std=0., mode="linear", color_mode="midpoint", use_viewdirs=False,
              ndc=False, near=2., far=6., t_rand=torch.from_numpy(t_rand).cuda(), u=torch.from_numpy(u).cuda())
    step = T.TrainStep(H, W, K, kw, N_rand=cfg["n"], lrate=5e-4, coarse_lrate=5e-4, lrate_decay=500)
    assert step._direct
    before = step.flat_params.clone()
    batch_rays = torch.stack([torch.from_numpy(ro), torch.from_numpy(rd)]).cuda()
    out = step.step_rays(batch_rays, torch.from_numpy(train_target(cfg)).cuda(), 1)
    assert abs(out["loss"].item() - float(g["train_loss"])) / float(g["train_loss"]) < 5e-3
    flat = step.bucket.flat
    off = 0
    worst = {}
    for tag, net in (("f", net_f), ("c", net_c)):
        for k, p in net.named_parameters():
            got = flat[off:off + p.numel()].double()
            moved = (step.flat_params[off:off + p.numel()] != before[off:off + p.numel()]).any().item()
            off += p.numel()
            ref_norm = float(g[f"gnorm_{tag}.{k}"])
            if k.startswith("views_linears"):
                assert ref_norm == 0.0 and not got.any().item() and not moved      # unused: no gradient, no update
                continue
            head = torch.from_numpy(g[f"ghead_{tag}.{k}"]).cuda().double()
            a = got[:head.numel()]
            rel = (a - head).norm().item() / head.norm().item()
            cos = torch.dot(a, head).item() / (a.norm().item() * head.norm().item())
            nrm = got.norm().item() / ref_norm
            worst[f"{tag}.{k}"] = (round(nrm, 4), round(rel, 4), round(cos, 5))
            trunk = k.startswith("pts_linears")
            # measured on a B200 (256 rays: a quarter of the config-3 case's samples, so the bf16 operand noise averages
            # less, and the bf16 forward moves the importance samples of the fine pass): norms within 1.3%; cosines 1.00000 at
            # output_linear, >= 0.9977 in the coarse trunk, >= 0.9768 in the fine trunk (its layer 0; >= 0.9888 above it)
            assert abs(nrm - 1) < (0.03 if trunk else 0.01), (tag, k, worst[f"{tag}.{k}"])
            assert cos > ((0.965 if tag == "f" else 0.995) if trunk else 0.9998), (tag, k, worst[f"{tag}.{k}"])
            assert rel < ((0.27 if tag == "f" else 0.10) if trunk else 0.02), (tag, k, worst[f"{tag}.{k}"])
            assert moved
            if k == "output_linear.weight":          # row 4 (the unused fifth channel): exactly zero
                assert not got.view(p.shape)[4].any().item()
    print("norm ratio / slice rel err / slice cos per parameter (no view directions):", worst)


@pytest.mark.skipif(not refimport.available(), reason="reference modules not reachable (run oracle/stage_ref.py)")
def test_training_convergence_vs_reference_fp32_autograd():
    from plnerf_b200 import ops, run_plnerf as RP, train as T
    Hh = Ww = 64
    focal = 0.5 * Ww / np.tan(0.5 * synth.LEGO_CAMERA_ANGLE_X)
    K = synth.intrinsics(Hh, Ww, focal)
    Ns, Ni, B, iters = 128, 64, 1024, 300
    pose_np = synth.pose_spherical(-60.0, -30.0, 4.0)[:3, :4].astype(np.float32).copy()
    pose = torch.from_numpy(pose_np).cuda()
    teacher_c, teacher_f = synth.nerf_params(1, **NET_KW), synth.nerf_params(2, **NET_KW)     # density-boosted: a scene with structure
    init_c = synth.nerf_params(21, density_boost=False, **NET_KW)
    init_f = synth.nerf_params(22, density_boost=False, **NET_KW)
    t_rand, u = pytest_draws(B, Ns, Ni)
    t_rand_d, u_d = torch.from_numpy(t_rand).cuda(), torch.from_numpy(u).cuda()
    full, _ = ops.pack_rays(Hh, Ww, K, c2w=pose, ndc=False, near=2., far=6., use_viewdirs=True)
    with torch.no_grad():       # the target image: the teacher rendered deterministically (parity mode)
        tr = RP.batchify_rays(full, 1024 * 32, network_fn=mk_net(teacher_c), network_query_fn=None, network_fine=mk_net(teacher_f),
                              N_samples=Ns, N_importance=Ni, perturb=0., white_bkgd=True, mode="linear", color_mode="midpoint",
                              precision="bf16x3")
    target = tr["rgb_map"].reshape(Hh, Ww, 3).contiguous()
    gen = torch.Generator(device="cuda")
    gen.manual_seed(3)
    pix_seq = [T.sample_pixels(Hh, Ww, B, "cuda", gen) for _ in range(iters)]

    def psnr_of(net_c, net_f):
        with torch.no_grad():
            r = RP.batchify_rays(full, 1024 * 32, network_fn=net_c, network_query_fn=None, network_fine=net_f, N_samples=Ns,
                                 N_importance=Ni, perturb=0., white_bkgd=True, mode="linear", color_mode="midpoint",
                                 precision="bf16x3")
        return float(-10. * torch.log10(torch.mean((r["rgb_map"] - target.reshape(-1, 3)) ** 2)))

    # ---- (a) this package: TrainStep, bf16
    a_c, a_f = mk_net(init_c), mk_net(init_f)
    kw = dict(network_query_fn=None, network_fn=a_c, network_fine=a_f, N_samples=Ns, N_importance=Ni, perturb=1.0,
              white_bkgd=True, raw_noise_std=0., mode="linear", color_mode="midpoint", use_viewdirs=True, ndc=False, near=2.,
              far=6., t_rand=t_rand_d, u=u_d)
    step = T.TrainStep(Hh, Ww, K, kw, N_rand=B, lrate=5e-4, coarse_lrate=5e-4, lrate_decay=500)
    loss_a = []
    for i in range(1, iters + 1):
        loss_a.append(step(target, pose, i, pix=pix_seq[i - 1])["loss"])
    loss_a = torch.stack(loss_a).cpu().numpy()
    psnr_a = psnr_of(a_c, a_f)

    # ---- (b) the unmodified reference in fp32 autograd on the GPU: its own render / img2mse / Adam, same batches and draws
    H, R = refimport.load()
    R.device = torch.device("cuda")
    torch.set_default_tensor_type('torch.cuda.FloatTensor')
    try:
        def mk_ref(prm):
            net = H.NeRF(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=[4], use_viewdirs=True)
            net.load_state_dict({k: torch.from_numpy(v.copy()).cuda() for k, v in prm.items()})
            return net.cuda()
        b_c, b_f = mk_ref(init_c), mk_ref(init_f)
        embed_fn, _ = H.get_embedder(10, 0)
        embeddirs_fn, _ = H.get_embedder(4, 0)
        q = lambda p, v, fn: R.run_network(p, v, fn, embed_fn=embed_fn, embeddirs_fn=embeddirs_fn, netchunk=1024 * 64)
        opt = torch.optim.Adam(params=list(b_f.parameters()), lr=5e-4, betas=(0.9, 0.999))
        opt_c = torch.optim.Adam(params=list(b_c.parameters()), lr=5e-4, betas=(0.9, 0.999))
        rays_o, rays_d = full[:, 0:3], full[:, 3:6]
        tflat = target.reshape(-1, 3)
        loss_b = []
        global_step = 0
        for i in range(1, iters + 1):
            pix = pix_seq[i - 1]
            batch_rays = torch.stack([rays_o[pix], rays_d[pix]], 0)
            target_s = tflat[pix]
            rgb, disp, acc, extras = R.render(Hh, Ww, K, chunk=1024 * 32, rays=batch_rays, verbose=False, retraw=True,
                                              constant_init=False, network_query_fn=q, perturb=1.0, N_importance=Ni,
                                              network_fine=b_f, N_samples=Ns, network_fn=b_c, use_viewdirs=True,
                                              white_bkgd=True, raw_noise_std=0., mode="linear", color_mode="midpoint",
                                              ndc=False, lindisp=False, near=2., far=6., pytest=True)
            opt.zero_grad(); opt_c.zero_grad()
            loss = H.img2mse(rgb, target_s) + H.img2mse(extras['rgb0'], target_s)
            loss.backward()
            opt.step(); opt_c.step()
            new_lrate = 5e-4 * (0.1 ** (global_step / (500 * 1000)))
            for pg in opt.param_groups + opt_c.param_groups:
                pg['lr'] = new_lrate
            global_step += 1
            loss_b.append(loss.detach())
        loss_b = torch.stack(loss_b).cpu().numpy()
    finally:
        torch.set_default_tensor_type('torch.FloatTensor')
    # evaluate the reference-trained weights with the same evaluator
    c_c, c_f = mk_net({k: v.detach().cpu().numpy() for k, v in b_c.state_dict().items()}), \
        mk_net({k: v.detach().cpu().numpy() for k, v in b_f.state_dict().items()})
    psnr_b = psnr_of(c_c, c_f)

    assert np.isfinite(loss_a).all() and np.isfinite(loss_b).all()
    assert abs(loss_a[0] - loss_b[0]) / loss_b[0] < 5e-3                     # same weights, same batch, same draws
    win = 25
    sm = lambda x: np.convolve(x, np.ones(win) / win, mode="valid")
    dev_curve = np.abs(sm(loss_a) - sm(loss_b)) / sm(loss_b)
    print(f"loss first/last a: {loss_a[0]:.5f}/{loss_a[-win:].mean():.5f}  b: {loss_b[0]:.5f}/{loss_b[-win:].mean():.5f}; "
          f"smoothed curve deviation max {dev_curve.max():.4f} mean {dev_curve.mean():.4f}; PSNR a {psnr_a:.3f} dB, b {psnr_b:.3f} dB")
    assert loss_b[-win:].mean() < 0.5 * loss_b[:win].mean()                   # the job actually optimises
    # measured on a B200: curve deviation max 0.22% / mean 0.06%, PSNR 49.909 vs 49.921 dB
    assert dev_curve.max() < 0.02 and dev_curve.mean() < 0.005               # loss curves (25-iteration means) within 2%
    assert abs(psnr_a - psnr_b) < 0.1                                         # final PSNR of the full image within 0.1 dB
