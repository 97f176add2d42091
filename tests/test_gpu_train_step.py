"""GPU tests of the device-side training step (pl-nerf_b200/train.py): the pixel-subset ray
kernel against the full-image packing (bit-exact), TrainStep against the plain "render -> img2mse -> backward ->
two Adam steps" sequence of the reference loop (run_plnerf.py:1283-1303) built from the same public API, and a
short optimisation run."""
import numpy as np
import pytest
import torch

from util import synth

pytestmark = pytest.mark.gpu

NET_KW = dict(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=(4,), use_viewdirs=True)


def make_net(seed):
    from plnerf_b200.run_nerf_helpers import NeRF
    net = NeRF(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=[4], use_viewdirs=True)
    net.load_state_dict({k: torch.from_numpy(v.copy())
                         for k, v in synth.nerf_params(seed, density_boost=False, **NET_KW).items()})
    return net.cuda()


def pose(theta):
    return torch.from_numpy(synth.pose_spherical(theta, -30.0, 4.0)[:3, :4].astype(np.float32).copy()).cuda()


@pytest.mark.parametrize("ndc,use_viewdirs", [(False, True), (True, True), (False, False)])
def test_pack_pixel_rays_bit_identical_to_full_image(ndc, use_viewdirs):
    from plnerf_b200 import ops
    H, W, focal = 37, 53, 44.5
    K = np.array([[focal, 0, 0.5 * W], [0, focal, 0.5 * H], [0, 0, 1]])
    c2w = pose(25.0)
    near, far = (0.0, 1.0) if ndc else (2.0, 6.0)
    full, _ = ops.pack_rays(H, W, K, c2w=c2w, ndc=ndc, near=near, far=far, use_viewdirs=use_viewdirs)
    rs = np.random.RandomState(1)
    pix = np.concatenate([[0, H * W - 1, W - 1, W, 5, 5], rs.randint(0, H * W, 300)]).astype(np.int64)   # duplicates allowed
    pix_t = torch.from_numpy(pix).cuda()
    got = ops.pack_pixel_rays(H, W, K, c2w, pix_t, ndc=ndc, near=near, far=far, use_viewdirs=use_viewdirs)
    assert got.shape == (pix.size, 11 if use_viewdirs else 8)
    assert torch.equal(got, full[pix_t])
    empty = ops.pack_pixel_rays(H, W, K, c2w, pix_t[:0], ndc=ndc, near=near, far=far, use_viewdirs=use_viewdirs)
    assert empty.shape == (0, got.shape[1])
    with pytest.raises(RuntimeError):
        ops.pack_pixel_rays(H, W, K, c2w, pix_t.cpu(), ndc=ndc, near=near, far=far, use_viewdirs=use_viewdirs)
    with pytest.raises(RuntimeError):
        ops.pack_pixel_rays(H, W, K, c2w, pix_t.int(), ndc=ndc, near=near, far=far, use_viewdirs=use_viewdirs)


def _render_kwargs(net_c, net_f, **extra):
    kw = dict(network_query_fn=None, network_fn=net_c, network_fine=net_f, N_samples=32, N_importance=32, perturb=1.0,
              white_bkgd=True, raw_noise_std=0., mode="linear", color_mode="midpoint", use_viewdirs=True, ndc=False,
              near=2., far=6.)
    kw.update(extra)
    return kw


def test_train_step_matches_reference_loop_sequence():
    """Same pixels, same Philox seed: TrainStep's gradients, loss and first Adam update against the reference loop's
    sequence (full-image rays -> gather -> render -> img2mse x2 -> backward -> optimizer.step x2) run through
    this package's public API with stock (unfused, separate) Adam optimisers."""
    from plnerf_b200 import ops, run_plnerf as RP, train as T
    H, W, focal, B = 40, 48, 55.0, 256
    K = np.array([[focal, 0, 0.5 * W], [0, focal, 0.5 * H], [0, 0, 1]])
    c2w = pose(-60.0)
    gen = torch.Generator(device="cuda")
    gen.manual_seed(5)
    target = torch.rand(H, W, 3, device="cuda", generator=gen)
    pix = T.sample_pixels(H, W, B, "cuda", gen)

    # --- this package's step
    a_c, a_f = make_net(61), make_net(62)
    step = T.TrainStep(H, W, K, _render_kwargs(a_c, a_f, seed=99), N_rand=B, lrate=5e-4, coarse_lrate=5e-4, lrate_decay=500)
    before = torch.cat([p.detach().flatten().clone() for p in step.bucket.params])
    out = step(target, c2w, i=7, pix=pix)
    grads_a = step.bucket.flat.clone()
    after_a = torch.cat([p.detach().flatten() for p in step.bucket.params])
    assert torch.equal(out["pix"], pix)
    assert step.optimizer.param_groups[0]["lr"] == T.decayed_lrate(5e-4, 500, 6)      # global_step = i - 1

    # --- the reference loop's sequence
    b_c, b_f = make_net(61), make_net(62)
    opt = torch.optim.Adam(b_f.parameters(), lr=5e-4, betas=(0.9, 0.999))
    opt_c = torch.optim.Adam(b_c.parameters(), lr=5e-4, betas=(0.9, 0.999))
    # full-image rays, then the reference's gathers (run_plnerf.py:1259,1277-1280).  The rays come from the full-image
    # kernel (held to torch's get_rays in test_pack_rays_vs_torch) so that both paths see bit-identical rays and the
    # comparison below isolates the step logic.
    full, _ = ops.pack_rays(H, W, K, c2w=c2w, ndc=False, near=2., far=6., use_viewdirs=True)
    rays_o, rays_d = full[:, 0:3].reshape(H, W, 3), full[:, 3:6].reshape(H, W, 3)
    coords = torch.stack([pix // W, pix % W], -1)
    batch_rays = torch.stack([rays_o[coords[:, 0], coords[:, 1]], rays_d[coords[:, 0], coords[:, 1]]], 0)
    target_s = target[coords[:, 0], coords[:, 1]]
    kw = _render_kwargs(b_c, b_f, seed=99)
    rgb, disp, acc, extras = RP.render(H, W, K, chunk=1024 * 32, rays=batch_rays, retraw=True, **kw)
    opt.zero_grad(); opt_c.zero_grad()
    img_loss = torch.mean((rgb - target_s) ** 2)
    img_loss0 = torch.mean((extras["rgb0"] - target_s) ** 2)
    (img_loss + img_loss0).backward()
    grads_b = torch.cat([p.grad.flatten() for p in list(b_f.parameters()) + list(b_c.parameters())])
    opt.step(); opt_c.step()
    after_b = torch.cat([p.detach().flatten() for p in list(b_f.parameters()) + list(b_c.parameters())])

    assert abs(out["img_loss"].item() - img_loss.item()) <= 2e-5 * img_loss.item()
    assert abs(out["img_loss0"].item() - img_loss0.item()) <= 2e-5 * img_loss0.item()
    assert abs(out["loss"].item() - (img_loss + img_loss0).item()) <= 2e-5 * (img_loss + img_loss0).item()
    # same kernels on the same inputs: only the loss-gradient rounding (one fp32 ulp before the bf16 operand
    # rounding) and the order of the weight-gradient atomics differ
    rel = (grads_a - grads_b).norm().item() / grads_b.norm().item()
    assert rel < 5e-3, rel
    assert grads_b.norm().item() > 0
    da, db = (after_a - before).double(), (after_b - before).double()
    assert float(da.abs().max()) <= 5e-4 * 1.001 and float((da != 0).float().mean()) > 0.2     # first Adam step: |dp| <= lr
    cos = torch.dot(da, db).item() / (da.norm().item() * db.norm().item())
    assert cos > 0.98, cos        # entries whose gradient is ~0 can flip sign between the two roundings (Adam: +-lr)


def test_train_step_optimises_with_precrop_and_constant_init():
    from plnerf_b200 import train as T
    H, W, focal, B = 64, 64, 70.0, 512
    K = np.array([[focal, 0, 0.5 * W], [0, focal, 0.5 * H], [0, 0, 1]])
    net_c, net_f = make_net(71), make_net(72)
    step = T.TrainStep(H, W, K, _render_kwargs(net_c, net_f), N_rand=B, lrate=5e-4, coarse_lrate=5e-4, lrate_decay=250,
                       precrop_iters=3, precrop_frac=0.5, constant_init=2, seed=1)
    target = torch.empty(H, W, 3, device="cuda")
    target[..., 0], target[..., 1], target[..., 2] = 0.2, 0.5, 0.8
    poses = [pose(t) for t in (-120.0, -30.0, 45.0, 150.0)]
    losses = []
    r0, c0, rows, cols = T.crop_window(H, W, 0.5)
    for i in range(60):
        out = step(target, poses[i % len(poses)], i)
        losses.append(out["loss"])
        if i < 3:
            r, c = out["pix"] // W, out["pix"] % W
            assert int(r.min()) >= r0 and int(r.max()) < r0 + rows and int(c.min()) >= c0 and int(c.max()) < c0 + cols
        assert torch.unique(out["pix"]).numel() == B
    losses = torch.stack(losses).cpu().numpy()
    assert np.isfinite(losses).all()
    assert losses[-5:].mean() < 0.8 * losses[:5].mean(), losses


def test_step_rays_equals_pixel_step():
    """The use_batching entry (caller-supplied [2, B, 3] rays + targets, run_plnerf.py:1238-1250) against the pixel entry on
    the same rays: the same packed rays, hence the same forward; gradients agree to the weight-gradient atomics' noise."""
    from plnerf_b200 import ops, train as T
    H, W, focal, B = 40, 48, 55.0, 256
    K = np.array([[focal, 0, 0.5 * W], [0, focal, 0.5 * H], [0, 0, 1]])
    c2w = pose(100.0)
    gen = torch.Generator(device="cuda")
    gen.manual_seed(8)
    target = torch.rand(H, W, 3, device="cuda", generator=gen)
    pix = T.sample_pixels(H, W, B, "cuda", gen)
    full, _ = ops.pack_rays(H, W, K, c2w=c2w, ndc=False, near=2., far=6., use_viewdirs=True)
    batch_rays = torch.stack([full[pix, 0:3], full[pix, 3:6]], 0)
    target_s = target.reshape(-1, 3)[pix]
    res = []
    for entry in ("pixels", "rays"):
        net_c, net_f = make_net(91), make_net(92)
        step = T.TrainStep(H, W, K, _render_kwargs(net_c, net_f, seed=77), N_rand=B, lrate=5e-4, coarse_lrate=5e-4)
        out = step(target, c2w, 0, pix=pix) if entry == "pixels" else step.step_rays(batch_rays, target_s, 0)
        res.append((out["loss"].item(), step.bucket.flat.clone()))
    (loss_a, g_a), (loss_b, g_b) = res
    assert abs(loss_a - loss_b) <= 2e-5 * abs(loss_a)                      # the gates of the reference-sequence test above
    assert (g_a - g_b).norm().item() <= 5e-3 * g_a.norm().item()


def test_mse_loss_grad_vs_torch():
    """plnerf_mse_loss_grad against img2mse x2 + autograd (run_nerf_helpers.py:17, run_plnerf.py:1289-1297), with and
    without the pixel gather, accumulation over two calls, coarse map optional."""
    from plnerf_b200 import ops
    gen = torch.Generator(device="cuda").manual_seed(3)
    for n in (1, 37, 1024, 5000):
        B = 2 * n
        rgb = torch.rand(n, 3, device="cuda", generator=gen, requires_grad=True)
        rgb0 = torch.rand(n, 3, device="cuda", generator=gen, requires_grad=True)
        image = torch.rand(4096, 3, device="cuda", generator=gen)
        pix = torch.randint(0, 4096, (n,), device="cuda", generator=gen)
        tgt = image[pix]
        loss = ((rgb - tgt) ** 2).sum() / (3 * B) + ((rgb0 - tgt) ** 2).sum() / (3 * B)
        loss.backward()
        sq = torch.zeros(2, device="cuda")
        g, g0 = ops.mse_loss_grad(rgb.detach(), rgb0.detach(), image, 2.0 / (3.0 * B), sq, pix=pix)
        assert torch.allclose(g, rgb.grad, rtol=1e-6, atol=1e-9) and torch.allclose(g0, rgb0.grad, rtol=1e-6, atol=1e-9)
        want = torch.stack([((rgb - tgt) ** 2).sum(), ((rgb0 - tgt) ** 2).sum()]).detach()
        assert torch.allclose(sq, want, rtol=2e-6)
        g2, none = ops.mse_loss_grad(rgb.detach(), None, tgt, 2.0 / (3.0 * B), sq)           # explicit targets, no coarse map
        assert none is None and torch.equal(g2, g)
        assert torch.allclose(sq, want * torch.tensor([2.0, 1.0], device="cuda"), rtol=2e-6)
        sq_a, sq_b = torch.zeros(2, device="cuda"), torch.zeros(2, device="cuda")            # reproducible bit for bit
        ops.mse_loss_grad(rgb.detach(), rgb0.detach(), image, 1.0, sq_a, pix=pix)
        ops.mse_loss_grad(rgb.detach(), rgb0.detach(), image, 1.0, sq_b, pix=pix)
        assert torch.equal(sq_a, sq_b)
    with pytest.raises(RuntimeError):
        ops.mse_loss_grad(rgb.detach().cpu(), None, tgt, 1.0, sq)


@pytest.mark.parametrize("n", [1, 7, 4096, 595844 * 2 + 3])
def test_adam_step_vs_torch_adam(n):
    """plnerf_adam_step against torch.optim.Adam (the reference's optimiser, run_plnerf.py:431-447) over 5 updates with a
    changing learning rate; unaligned segments (a view starting at an odd element) take the scalar path."""
    from plnerf_b200 import ops
    gen = torch.Generator(device="cuda").manual_seed(n)
    for misalign in (0, 1):
        store = [torch.zeros(n + 1, device="cuda") for _ in range(4)]
        p, g, m, v = (t[misalign:misalign + n] for t in store)
        p.copy_(torch.randn(n, device="cuda", generator=gen))
        ref = torch.nn.Parameter(p.clone())
        opt = torch.optim.Adam([ref], lr=5e-4, betas=(0.9, 0.999))
        for step in range(1, 6):
            lr = 5e-4 * (0.1 ** (step / 7.0))
            grad = torch.randn(n, device="cuda", generator=gen) * (10.0 ** float(step - 3))
            g.copy_(grad)
            ref.grad = grad.clone()
            for group in opt.param_groups:
                group["lr"] = lr
            opt.step()
            ops.adam_step(p, g, m, v, lr, step)
            assert torch.equal(g, grad)                                    # gradients are kept unless asked otherwise
            st = opt.state[ref]
            assert torch.allclose(m, st["exp_avg"], rtol=1e-6, atol=1e-12)
            assert torch.allclose(v, st["exp_avg_sq"], rtol=1e-6, atol=1e-20)
            assert torch.allclose(p, ref.data, rtol=3e-7, atol=1e-9)        # <= 2 ulp of the parameter
        ops.adam_step(p, g, m, v, 0.0, 6, zero_grads=True)
        assert not g.any()
    with pytest.raises(RuntimeError):
        ops.adam_step(p.cpu(), g, m, v, 1e-3, 1)
    with pytest.raises(RuntimeError):
        ops.adam_step(p, g, m, v, 1e-3, 0)


def test_train_step_is_chunk_invariant():
    """N_rand rays rendered in several chunks (batchify_rays' loop, run_plnerf.py:95-107) give the step of one chunk: the
    device draws are keyed by the global ray id, the loss sums accumulate across chunks in order, the parameter gradients
    add up to the weight-gradient atomics' noise."""
    from plnerf_b200 import train as T
    H, W, focal, B = 40, 48, 55.0, 256
    K = np.array([[focal, 0, 0.5 * W], [0, focal, 0.5 * H], [0, 0, 1]])
    c2w = pose(40.0)
    gen = torch.Generator(device="cuda")
    gen.manual_seed(5)
    target = torch.rand(H, W, 3, device="cuda", generator=gen)
    pix = T.sample_pixels(H, W, B, "cuda", gen)
    res = []
    for chunk in (1024 * 32, 96):                     # 96: chunks of 96 + 96 + 64 rays
        net_c, net_f = make_net(91), make_net(92)
        step = T.TrainStep(H, W, K, _render_kwargs(net_c, net_f, seed=31), N_rand=B, chunk=chunk, lrate=5e-4, coarse_lrate=5e-4)
        out = step(target, c2w, 0, pix=pix)
        res.append((out["loss"].item(), out["img_loss"].item(), out["img_loss0"].item(), step.bucket.flat.clone(),
                    step.flat_params.clone()))
    a, b = res
    for i in range(3):
        assert abs(a[i] - b[i]) <= 2e-6 * abs(a[i]), (i, a[i], b[i])
    assert abs(a[0] - (a[1] + a[2])) <= 1e-6 * abs(a[0])
    assert (a[3] - b[3]).norm().item() <= 5e-3 * a[3].norm().item()
    assert (a[4] - b[4]).abs().max().item() <= 2 * 5e-4 + 1e-7    # one Adam step: at most +-lr where a ~0 gradient flips sign


@pytest.mark.parametrize("use_viewdirs", [True, False])
def test_train_step_coarse_only(use_viewdirs):
    """N_importance = 0 / network_fine = None (BASELINE config 1's shape, with and without view directions): one network, one
    optimiser segment, the loss is img2mse(rgb_map) alone; against render() + autograd + a stock Adam on a twin network."""
    from plnerf_b200 import ops, run_plnerf as RP, train as T
    from plnerf_b200.run_nerf_helpers import NeRF
    H, W, focal, B = 40, 48, 55.0, 192
    K = np.array([[focal, 0, 0.5 * W], [0, focal, 0.5 * H], [0, 0, 1]])
    c2w = pose(70.0)
    kwn = dict(D=8, W=256, input_ch=63, input_ch_views=27 if use_viewdirs else 0, output_ch=4, skips=(4,), use_viewdirs=use_viewdirs)

    def mk():
        net = NeRF(D=8, W=256, input_ch=63, input_ch_views=kwn["input_ch_views"], output_ch=4, skips=[4], use_viewdirs=use_viewdirs)
        net.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in synth.nerf_params(5, density_boost=False, **kwn).items()})
        return net.cuda()
    net, twin = mk(), mk()
    gen = torch.Generator(device="cuda")
    gen.manual_seed(2)
    target = torch.rand(H, W, 3, device="cuda", generator=gen)
    pix = T.sample_pixels(H, W, B, "cuda", gen)
    kw = dict(network_query_fn=None, network_fn=net, network_fine=None, N_samples=64, N_importance=0, perturb=1.0,
              white_bkgd=True, raw_noise_std=0., mode="linear", color_mode="midpoint", use_viewdirs=use_viewdirs, ndc=False,
              near=2., far=6., seed=11)
    step = T.TrainStep(H, W, K, kw, N_rand=B, lrate=5e-4, coarse_lrate=5e-4)
    assert len(step.optimizer.param_groups) == 1 and list(step.optimizer.segments) == ["coarse"]
    out = step(target, c2w, 0, pix=pix)
    assert out["img_loss0"] is None and out["loss"].item() == out["img_loss"].item()
    # the reference sequence on the twin
    full, _ = ops.pack_rays(H, W, K, c2w=c2w, ndc=False, near=2., far=6., use_viewdirs=use_viewdirs)
    rays = torch.stack([full[pix, 0:3], full[pix, 3:6]], 0)
    kw_t = dict(kw, network_fn=twin)
    opt = torch.optim.Adam(twin.parameters(), lr=5e-4, betas=(0.9, 0.999))
    rgb, disp, acc, extras = RP.render(H, W, K, chunk=1024 * 32, rays=rays, **kw_t)
    loss = torch.mean((rgb - target.reshape(-1, 3)[pix]) ** 2)
    opt.zero_grad()
    loss.backward()
    opt.step()
    assert abs(out["loss"].item() - loss.item()) <= 2e-5 * abs(loss.item())
    g_ref = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).flatten() for p in twin.parameters()])
    assert (step.bucket.flat - g_ref).norm().item() <= 5e-3 * g_ref.norm().item()
    p_a = torch.cat([p.detach().flatten() for p in net.parameters()])
    p_b = torch.cat([p.detach().flatten() for p in twin.parameters()])
    assert (p_a - p_b).abs().max().item() <= 2 * 5e-4 + 1e-7


@pytest.mark.parametrize("use_viewdirs", [True, False])
def test_repack_train_equals_separate_packs(use_viewdirs):
    """ops.repack_train (plnerf_pack_weights_train: forward stream + tail + transposed stream of both networks in one launch)
    writes the same bytes as plnerf_pack_weights / plnerf_pack_weights_bwd, and leaves both caches current."""
    from plnerf_b200 import ops
    from plnerf_b200.run_nerf_helpers import NeRF
    torch.manual_seed(3)
    nets = [NeRF(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=[4], use_viewdirs=use_viewdirs).cuda()
            for _ in range(2)]
    ref = [(ops.packed_of(n).get(n, "bf16").clone(), ops.packed_bwd_of(n).clone()) for n in nets]
    launches = ops.launch_count()
    with torch.no_grad():
        for n in nets:
            for p in n.parameters():
                p.data.mul_(1.25)                    # through .data: the version counters do not move (like the flat Adam)
    for n in nets:
        ops.invalidate_packed(n)
        ops.packed_of(n).buf[ops._prec("bf16")].zero_()
        n.__dict__["_plnerf_packed_bwd"]["buf"].zero_()
    launches = ops.launch_count()
    ops.repack_train(nets)
    assert ops.launch_count() - launches == 1
    launches = ops.launch_count()
    got = [(ops.packed_of(n).get(n, "bf16").clone(), ops.packed_bwd_of(n).clone()) for n in nets]
    assert ops.launch_count() == launches            # caches are current: no lazy repack
    for n in nets:
        ops.invalidate_packed(n)
    want = [(ops.packed_of(n).get(n, "bf16").clone(), ops.packed_bwd_of(n).clone()) for n in nets]
    for (g0, g1), (w0, w1), (r0, r1) in zip(got, want, ref):
        assert torch.equal(g0, w0) and torch.equal(g1, w1)
        assert not torch.equal(g0, r0) and not torch.equal(g1, r1)     # (the parameters did change)


@pytest.mark.parametrize("n_importance,shared,noisy", [(64, False, False), (64, True, False), (0, False, False), (64, False, True)])
def test_train_rays_mse_equals_the_three_calls(n_importance, shared, noisy):
    """plnerf_train_rays_mse (forward + both MSE terms + backward in one call, the coarse backward forked beside the fine
    pass) against plnerf_render_rays_fwd_train -> plnerf_mse_loss_grad -> plnerf_render_rays_bwd on the same rays and draws:
    maps and loss sums bit for bit, parameter gradients to the weight-gradient atomics' noise.  `shared`: network_fine=None
    (the coarse network serves both passes, both backward passes add into the same buffers concurrently); `noisy`: explicit
    density noise and explicit draws."""
    from plnerf_b200 import ops
    n, Ns = 200, 64
    gen = torch.Generator(device="cuda")
    gen.manual_seed(9)
    H, W, focal = 40, 48, 55.0
    K = np.array([[focal, 0, 0.5 * W], [0, focal, 0.5 * H], [0, 0, 1]])
    full, _ = ops.pack_rays(H, W, K, c2w=pose(25.0), ndc=False, near=2., far=6., use_viewdirs=True)
    pix = torch.randperm(H * W, device="cuda", generator=gen)[:n].contiguous()
    rays = full[pix].contiguous()
    image = torch.rand(H * W, 3, device="cuda", generator=gen)
    net_c, net_f = make_net(61), (None if (shared or n_importance == 0) else make_net(62))
    args = (rays, net_c, net_f, Ns, n_importance, "linear", "midpoint")
    kw = dict(perturb=True, white_bkgd=True, seed=77, ray_id_offset=1000)
    if noisy:       # explicit density noise (raw_noise_std > 0 in the training configs): the same arrays reach forward and backward
        kw.update(noise0=torch.randn(n, Ns, device="cuda", generator=gen), noise1=torch.randn(n, Ns + n_importance, device="cuda", generator=gen),
                  t_rand=torch.rand(n, Ns, device="cuda", generator=gen), u=torch.rand(n, n_importance, device="cuda", generator=gen))
    scale = 2.0 / (3.0 * n)

    def zero_grads(net):
        return None if net is None else {k: torch.zeros_like(p) for k, p in net.named_parameters()}
    # the three calls
    ga_c, ga_f, sq_a = zero_grads(net_c), zero_grads(net_f), torch.zeros(2, device="cuda")
    ret, ctx = ops.render_rays_fwd_train(*args, **kw)
    g, g0 = ops.mse_loss_grad(ret["rgb_map"], ret.get("rgb0"), image, scale, sq_a, pix=pix)
    if n_importance > 0:
        ops.render_rays_bwd(ctx, (g, None, None, None), (g0, None, None, None), ga_c, ga_f if net_f is not None else ga_c)
    else:
        ops.render_rays_bwd(ctx, (g, None, None, None), None, ga_c, None)
    # one call
    gb_c, gb_f, sq_b = zero_grads(net_c), zero_grads(net_f), torch.zeros(2, device="cuda")
    maps = ops.train_rays_mse(*args, image, scale, sq_b, gb_c, gb_f, pix=pix, want_maps=True, **kw)
    torch.cuda.synchronize()
    for k in maps:       # bit for bit (as int32: a ray with zero opacity has a NaN disparity in the reference too, run_plnerf.py:608)
        assert torch.equal(maps[k].view(torch.int32), ret[k].view(torch.int32)), k
    assert torch.equal(sq_a, sq_b) and sq_a[0].item() > 0 and (sq_a[1].item() > 0) == (n_importance > 0)
    for ga, gb in ((ga_c, gb_c), (ga_f, gb_f)):
        if ga is None:
            continue
        fa, fb = torch.cat([v.flatten() for v in ga.values()]), torch.cat([v.flatten() for v in gb.values()])
        assert fa.norm().item() > 0
        assert (fa - fb).norm().item() <= 1e-4 * fa.norm().item()
    # without map outputs: same gradients again (the maps then live in the workspace)
    gc_c, gc_f, sq_c = zero_grads(net_c), zero_grads(net_f), torch.zeros(2, device="cuda")
    assert ops.train_rays_mse(*args, image, scale, sq_c, gc_c, gc_f, pix=pix, **kw) is None
    assert torch.equal(sq_c, sq_a)
    fb, fc = torch.cat([v.flatten() for v in gb_c.values()]), torch.cat([v.flatten() for v in gc_c.values()])
    assert (fb - fc).norm().item() <= 1e-4 * fb.norm().item()
