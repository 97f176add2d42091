"""Host logic of the training step (pl-nerf_b200/train.py) that needs no GPU: the pixel draw
(the law of run_plnerf.py:1259-1280), the precrop window and the learning-rate decay (:1307-1315)."""
import numpy as np
import pytest
import torch

import plnerf_b200.train as T


def test_crop_window_matches_reference_linspace():
    for H, W, frac in ((800, 800, 0.5), (378, 504, 0.5), (401, 301, 0.3)):
        r0, c0, rows, cols = T.crop_window(H, W, frac)
        dH, dW = int(H // 2 * frac), int(W // 2 * frac)
        # the reference's coords: linspace(H//2 - dH, H//2 + dH - 1, 2*dH) x linspace(W//2 - dW, W//2 + dW - 1, 2*dW)
        rr = torch.linspace(H // 2 - dH, H // 2 + dH - 1, 2 * dH).long()
        cc = torch.linspace(W // 2 - dW, W // 2 + dW - 1, 2 * dW).long()
        assert (r0, rows) == (int(rr[0]), rr.numel()) and int(rr[-1]) == r0 + rows - 1
        assert (c0, cols) == (int(cc[0]), cc.numel()) and int(cc[-1]) == c0 + cols - 1


@pytest.mark.parametrize("frac", [None, 0.5])
def test_sample_pixels_distinct_in_window_and_reproducible(frac):
    H, W, N = 60, 80, 512
    gen = torch.Generator(device="cpu")
    gen.manual_seed(3)
    pix = T.sample_pixels(H, W, N, "cpu", gen, frac)
    assert pix.dtype == torch.int64 and pix.shape == (N,)
    assert torch.unique(pix).numel() == N                                   # replace=False
    r, c = pix // W, pix % W
    r0, c0, rows, cols = (0, 0, H, W) if frac is None else T.crop_window(H, W, frac)
    assert int(r.min()) >= r0 and int(r.max()) < r0 + rows and int(c.min()) >= c0 and int(c.max()) < c0 + cols
    gen.manual_seed(3)
    assert torch.equal(pix, T.sample_pixels(H, W, N, "cpu", gen, frac))
    assert not torch.equal(pix, T.sample_pixels(H, W, N, "cpu", gen, frac))    # the stream advances
    with pytest.raises(ValueError):
        T.sample_pixels(4, 4, 17, "cpu", gen)


def test_sample_pixels_is_uniform():
    """Every pixel of the window is equally likely: chi-square of 4000 draws of 8 pixels from a 6 x 8 window."""
    H, W, N, trials = 12, 16, 8, 4000
    gen = torch.Generator(device="cpu")
    gen.manual_seed(0)
    r0, c0, rows, cols = T.crop_window(H, W, 0.5)
    counts = np.zeros((H, W))
    for _ in range(trials):
        p = T.sample_pixels(H, W, N, "cpu", gen, 0.5).numpy()
        np.add.at(counts, (p // W, p % W), 1)
    win = counts[r0:r0 + rows, c0:c0 + cols]
    assert counts.sum() == win.sum() == trials * N
    exp = trials * N / win.size
    chi2 = ((win - exp) ** 2 / exp).sum()
    assert chi2 < 2.0 * win.size, chi2          # dof = 47; 2x dof is > 6 sigma


def test_decayed_lrate_formula():
    # run_plnerf.py:1307-1309 with blender_linear.txt's lrate_decay = 500
    for step in (0, 1, 1000, 250000, 500000):
        want = 5e-4 * (0.1 ** (step / (500 * 1000)))
        assert T.decayed_lrate(5e-4, 500, step) == want
    assert T.decayed_lrate(5e-4, 500, 500000) == pytest.approx(5e-5)


def test_train_step_needs_cuda_models():
    from plnerf_b200.run_nerf_helpers import NeRF
    net = NeRF(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=[4], use_viewdirs=True)
    with pytest.raises(RuntimeError, match="CUDA"):
        T.TrainStep(8, 8, np.eye(3), dict(network_fn=net, network_fine=net, N_samples=8, N_importance=8))


@pytest.mark.parametrize("n_items,n_draws", [(100, 37), (64 * 16, 16), (640000, 1024)])
def test_distinct_draws_rows_are_distinct_and_in_range(n_items, n_draws):
    """Both branches (batched permutation for small windows, first-occurrence-distinct draws for large ones)."""
    gen = torch.Generator(device="cpu")
    gen.manual_seed(11)
    k = T.distinct_draws(n_items, n_draws, 5, "cpu", gen)
    assert k.shape == (5, n_draws) and k.dtype == torch.int64
    assert int(k.min()) >= 0 and int(k.max()) < n_items
    for row in k:
        assert torch.unique(row).numel() == n_draws
    assert not torch.equal(k[0], k[1])
    with pytest.raises(ValueError):
        T.distinct_draws(10, 11, 1, "cpu", gen)


def test_distinct_draws_large_window_law():
    """Large-window branch: uniform over items (chi-square) and the FIRST-occurrence order of the underlying draws
    is kept (a row is the de-duplicated prefix of its own i.i.d. stream)."""
    n_items, n_draws, rows = 64 * 8, 8, 6000
    gen = torch.Generator(device="cpu")
    gen.manual_seed(2)
    k = T.distinct_draws(n_items, n_draws, rows, "cpu", gen)
    counts = np.bincount(k.numpy().ravel(), minlength=n_items)
    exp = rows * n_draws / n_items
    chi2 = ((counts - exp) ** 2 / exp).sum()
    assert chi2 < 1.3 * n_items, chi2                    # dof = 511, sigma = 32: 1.3x is ~5 sigma
    # replay: the same generator state gives the same i.i.d. draws; de-duplicate them on the host
    gen.manual_seed(2)
    c = torch.randint(0, n_items, (rows, 2 * n_draws), device="cpu", generator=gen).numpy()
    for r in (0, 1, 17, rows - 1):
        seen, want = set(), []
        for v in c[r]:
            if v not in seen:
                seen.add(v)
                want.append(v)
        assert k[r].tolist() == want[:n_draws]


def test_pixel_sampler_blocks_and_window_change():
    H, W, N = 40, 50, 16
    gen = torch.Generator(device="cpu")
    gen.manual_seed(9)
    s = T.PixelSampler(H, W, N, "cpu", gen, block=4)
    r0, c0, rows, cols = T.crop_window(H, W, 0.5)
    seen = []
    for i in range(11):
        frac = 0.5 if i < 3 else None                    # the precrop phase ends inside the first block
        pix = s.next(frac)
        assert pix.shape == (N,) and torch.unique(pix).numel() == N
        r, c = pix // W, pix % W
        if frac is not None:
            assert int(r.min()) >= r0 and int(r.max()) < r0 + rows and int(c.min()) >= c0 and int(c.max()) < c0 + cols
        seen.append(pix)
    assert any(int((p // W).min()) < r0 or int((p // W).max()) >= r0 + rows for p in seen[3:])   # full image afterwards
    assert len({tuple(p.tolist()) for p in seen}) == len(seen)
    gen.manual_seed(9)
    s2 = T.PixelSampler(H, W, N, "cpu", gen, block=4)
    assert all(torch.equal(s2.next(0.5 if i < 3 else None), seen[i]) for i in range(11))         # reproducible


def test_alias_parameters_flat_keeps_values_and_state_dict():
    from plnerf_b200.run_nerf_helpers import NeRF
    net = NeRF(D=2, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=[], use_viewdirs=True)
    before = {k: v.clone() for k, v in net.state_dict().items()}
    params = list(net.parameters())
    flat = T.alias_parameters_flat(params)
    assert flat.numel() == sum(p.numel() for p in params)
    for k, v in net.state_dict().items():
        assert torch.equal(v, before[k])
    flat.mul_(2.0)                                       # an update through the alias is seen by every parameter
    for k, v in net.state_dict().items():
        assert torch.equal(v, before[k] * 2.0)
    net.load_state_dict(before)                          # and loading writes through the views into the flat buffer
    off = 0
    for p in params:
        assert torch.equal(flat[off:off + p.numel()].view_as(p), p.data)
        off += p.numel()
    assert torch.equal(flat[:params[0].numel()].view_as(params[0]), before["pts_linears.0.weight"])


def test_train_step_host_logic_equals_two_stock_adams(monkeypatch):
    """The step's own bookkeeping -- pixel batch -> targets, direct loss gradient, flat gradient buffer, ONE fused Adam
    over flat parameter segments, learning-rate decay with the reference's coarse-gets-fine-rate quirk, packed-copy
    invalidation -- against the reference loop's sequence (img2mse x2 -> backward -> optimizer.step x2 -> lr loop,
    run_plnerf.py:1286-1315) with two stock Adam optimisers.  The renderer and the ray kernel are replaced by small
    differentiable torch stand-ins (this test has no GPU; the real kernels are held to the same sequence in
    tests/test_gpu_train_step.py), so the two runs must agree bit for bit."""
    from plnerf_b200 import ops, run_plnerf as RP
    from plnerf_b200.run_nerf_helpers import NeRF
    H = W = 16

    def fake_pack(H, W, K, pose, pix, ndc, near, far, use_viewdirs):
        r = torch.zeros(pix.shape[0], 11)
        r[:, 0], r[:, 1] = (pix % W).float() / W, (pix // W).float() / H
        return r

    def fake_batchify(rays, chunk, ray_id_offset=0, retraw=False, constant_init=False, network_query_fn="unset",
                      pytest=False, **kw):
        assert network_query_fn is None and retraw is False
        x = torch.cat([rays[:, :3]] * 21, -1)
        return {"rgb_map": torch.sigmoid(kw["network_fine"].pts_linears[0](x)[:, :3]),
                "rgb0": torch.sigmoid(kw["network_fn"].pts_linears[0](x)[:, :3])}
    # stand-ins for the kernel pair of the direct path (autograd.forward_stashed / backward_stashed): same outs layout,
    # parameter gradients ACCUMULATED into the dicts the step hands over
    def fake_forward(cfg, rays):
        with torch.enable_grad():
            r = fake_batchify(rays, 1, network_query_fn=None, network_fn=cfg["net_c"], network_fine=cfg["net_f"])
        z = torch.zeros(rays.shape[0])
        outs = (r["rgb_map"].detach(), z, z, z, None, r["rgb0"].detach(), z, z, z, z)
        return outs, (r["rgb_map"], r["rgb0"]), (None, None)

    def fake_backward(cfg, saved, stashes, g_fine, g_coarse, grads_c, grads_f):
        assert g_fine[1:] == (None, None, None) and g_coarse[1:] == (None, None, None)
        for out, g, net, grads in ((saved[0], g_fine[0], cfg["net_f"], grads_f), (saved[1], g_coarse[0], cfg["net_c"], grads_c)):
            names = [k for k, _ in net.named_parameters()]
            got = torch.autograd.grad(out, list(net.parameters()), g, allow_unused=True)
            for k, gk in zip(names, got):
                if gk is not None:
                    grads[k].add_(gk)
    from plnerf_b200 import autograd as AG
    invalidated = []
    def fake_pack_rays(H, W, K, rays=None, ndc=True, near=0., far=1., use_viewdirs=False, **kw):
        r = torch.zeros(rays[0].shape[0], 11)
        r[:, 0:3], r[:, 3:6] = rays[0], rays[1]
        return r, tuple(rays[1].shape)
    monkeypatch.setattr(ops, "pack_pixel_rays", fake_pack)
    monkeypatch.setattr(ops, "pack_rays", fake_pack_rays)
    monkeypatch.setattr(RP, "batchify_rays", fake_batchify)
    monkeypatch.setattr(AG, "forward_stashed", fake_forward)
    monkeypatch.setattr(AG, "backward_stashed", fake_backward)

    def fake_train_rays_mse(cfg, rays, target, pix, scale, sqerr, grads_c, grads_f):
        # the fused entry (plnerf_train_rays_mse) = forward -> the two MSE terms -> backward of the same ray batch
        outs, saved, stashes = fake_forward(cfg, rays)
        g, g0 = ops.mse_loss_grad(outs[0], outs[5], target, scale, sqerr, pix=pix)
        fake_backward(cfg, saved, stashes, (g, None, None, None), (g0, None, None, None), grads_c, grads_f)
    monkeypatch.setattr(AG, "train_rays_mse", fake_train_rays_mse)
    monkeypatch.setattr(ops, "invalidate_packed", lambda net: invalidated.append(net))

    # torch stand-ins for the two small kernels of the step (tests/util.py)
    from util import fake_mse_loss_grad as fake_mse, fake_adam_step as fake_adam
    monkeypatch.setattr(ops, "mse_loss_grad", fake_mse)
    monkeypatch.setattr(ops, "adam_step", fake_adam)

    class HostOnlyStep(T.TrainStep):
        def _check_device(self):       # the stand-ins above run on the host
            pass

    mk = lambda: NeRF(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=[4], use_viewdirs=True)
    for coarse_lrate, n_groups, extra in ((5e-4, 1, {}), (1e-4, 2, {}), (5e-4, 1, {"pytest": True})):
        torch.manual_seed(0)
        net_c, net_f, ref_c, ref_f = mk(), mk(), mk(), mk()
        ref_c.load_state_dict(net_c.state_dict())
        ref_f.load_state_dict(net_f.state_dict())
        kw = dict(network_fn=net_c, network_fine=net_f, N_samples=8, N_importance=8, perturb=1., white_bkgd=True,
                  raw_noise_std=0., mode="linear", color_mode="midpoint", use_viewdirs=True, ndc=False, near=2., far=6.,
                  **extra)
        step = HostOnlyStep(H, W, np.eye(3), kw, N_rand=64, precrop_iters=2, constant_init=1, lrate=5e-4,
                            coarse_lrate=coarse_lrate, lrate_decay=500)
        assert len(step.optimizer.param_groups) == n_groups
        assert step._direct == (not extra)        # explicit draws / the pytest hook go through render_rays' autograd.Function
        opt = torch.optim.Adam(ref_f.parameters(), lr=5e-4, betas=(0.9, 0.999))
        opt_c = torch.optim.Adam(ref_c.parameters(), lr=coarse_lrate, betas=(0.9, 0.999))
        target = torch.rand(H, W, 3)
        for i in range(4):
            if i < 3:
                out = step(target, torch.eye(4)[:3], i)
                pix = out["pix"]
                tgt = target.reshape(-1, 3)[pix]
                packed = fake_pack(H, W, None, None, pix, False, 2., 6., True)
            else:       # the use_batching branch: the caller hands over [2, B, 3] rays + their targets
                gen = torch.Generator().manual_seed(40 + n_groups)
                batch_rays, tgt = torch.rand(2, 48, 3, generator=gen), torch.rand(48, 3, generator=gen)
                out = step.step_rays(batch_rays, tgt, i)
                assert "pix" not in out
                packed = fake_pack_rays(H, W, None, rays=(batch_rays[0], batch_rays[1]))[0]
            r = fake_batchify(packed, 1, network_query_fn=None, network_fn=ref_c, network_fine=ref_f)
            opt.zero_grad(); opt_c.zero_grad()
            loss = torch.mean((r["rgb_map"] - tgt) ** 2) + torch.mean((r["rgb0"] - tgt) ** 2)
            loss.backward()
            opt.step(); opt_c.step()
            new_lrate = 5e-4 * (0.1 ** (max(i - 1, 0) / (500 * 1000)))     # global_step = i - 1 (run_plnerf.py:1153, 1235, 1400)
            for g in opt.param_groups + opt_c.param_groups:     # run_plnerf.py:1310-1315 (both get the fine rate)
                g["lr"] = new_lrate
            assert float(out["loss"]) == pytest.approx(float(loss.detach()), rel=1e-6)
            for a, b in zip(list(net_f.parameters()) + list(net_c.parameters()),
                            list(ref_f.parameters()) + list(ref_c.parameters())):
                assert torch.equal(a.data, b.data)
        assert len(invalidated) >= 8
    assert len(invalidated) == 3 * 4 * 2


def test_flat_adam_state_round_trips_through_the_reference_optimizer_state(monkeypatch):
    """FlatAdam's flat moments <-> the per-tensor ``optimizer.state_dict()`` the reference checkpoints
    (run_plnerf.py:1326-1331 saves it, :466 restores it): export after two steps loads into a stock Adam that then takes
    the same third step; import of the stock state reproduces the flat state."""
    from plnerf_b200 import ops

    from util import fake_adam_step as fake_adam
    monkeypatch.setattr(ops, "adam_step", fake_adam)
    torch.manual_seed(0)
    fine = [torch.nn.Parameter(torch.randn(5, 3)), torch.nn.Parameter(torch.randn(7))]
    coarse = [torch.nn.Parameter(torch.randn(4, 4))]
    params = fine + coarse
    grads = torch.zeros(sum(p.numel() for p in params))
    flat = T.alias_parameters_flat(params)
    opt = T.FlatAdam(flat, grads, [("fine", 22, 5e-4), ("coarse", 16, 1e-4)])
    assert len(opt.param_groups) == 2 and opt.param_groups[1]["lr"] == 1e-4
    for _ in range(2):
        grads.copy_(torch.randn_like(grads))
        opt.step()
    sd = opt.export_reference_state("fine", fine)
    twins = [torch.nn.Parameter(p.detach().clone()) for p in fine]
    stock = torch.optim.Adam(twins, lr=5e-4, betas=(0.9, 0.999))
    stock.load_state_dict(sd)
    grads.copy_(torch.randn_like(grads))
    off = 0
    for t in twins:
        t.grad = grads[off:off + t.numel()].view_as(t).clone()
        off += t.numel()
    opt.step()
    stock.step()
    for a, b in zip(fine, twins):
        assert torch.equal(a.data, b.data)
    # and back: a fresh flat optimiser that imports the stock state continues identically
    opt2 = T.FlatAdam(flat.clone(), grads, [("fine", 22, 5e-4), ("coarse", 16, 1e-4)])
    opt2.import_reference_state("fine", fine, stock.state_dict())
    assert opt2.step_count == 3
    assert torch.equal(opt2.exp_avg[:22], opt.exp_avg[:22]) and torch.equal(opt2.exp_avg_sq[:22], opt.exp_avg_sq[:22])
    rt = T.FlatAdam(flat.clone(), grads, [("fine", 22, 5e-4), ("coarse", 16, 1e-4)])
    rt.load_state_dict(opt.state_dict())
    assert rt.step_count == 3 and torch.equal(rt.exp_avg, opt.exp_avg)
    with pytest.raises(ValueError):
        opt.export_reference_state("fine", coarse)
