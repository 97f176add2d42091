"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: ray sharding, global-ray-id offsets,
output gather and the single flat gradient all-reduce."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, ws, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    import plnerf_b200.dist as D
    try:
        n = 1001
        rays = torch.arange(n * 11, dtype=torch.float32).reshape(n, 11)
        local, lo = D.shard_rays(rays)
        lo2, hi2 = D.shard_bounds(n)
        assert lo == lo2 and local.shape[0] == hi2 - lo2
        assert torch.equal(local, rays[lo2:hi2])

        # a fake render_rays that returns the global id it was told about: offsets must be global
        seen = []

        def fake_render(chunk_rays, ray_id_offset=0, **kw):
            seen.append((ray_id_offset, chunk_rays.shape[0]))
            ids = torch.arange(chunk_rays.shape[0], dtype=torch.float32) + ray_id_offset
            return {"rgb_map": torch.stack([ids, ids, ids], -1), "acc_map": chunk_rays[:, 0]}
        out = D.render_sharded(fake_render, rays, chunk=128)
        assert seen[0][0] == lo and sum(c for _, c in seen) == hi2 - lo2
        full = D.gather_rays(out["rgb_map"], n)
        assert full.shape == (n, 3) and torch.equal(full[:, 0], torch.arange(n, dtype=torch.float32))
        acc = D.gather_rays(out["acc_map"], n)
        assert torch.equal(acc, rays[:, 0])

        # flat gradient bucket: one all-reduce, mean over ranks
        import plnerf_b200.run_nerf_helpers as H
        net = H.NeRF(D=2, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=[], use_viewdirs=True)
        bucket = D.FlatGradBucket([net, None])
        assert bucket.flat.numel() == sum(p.numel() for p in net.parameters())
        for p in net.parameters():
            p.grad.fill_(float(rank + 1))
        calls = {"n": 0}
        real = dist.all_reduce

        def counting(*a, **k):
            calls["n"] += 1
            return real(*a, **k)
        dist.all_reduce = counting
        D.allreduce_gradients(bucket)
        dist.all_reduce = real
        assert calls["n"] == 1
        want = sum(range(1, ws + 1)) / ws
        assert all(torch.all(p.grad == want) for p in net.parameters())
        # sum-only variant (train.TrainStep scales its local loss gradient by the GLOBAL batch size)
        for p in net.parameters():
            p.grad.fill_(float(rank + 1))
        bucket.allreduce_sum()
        assert all(torch.all(p.grad == float(sum(range(1, ws + 1)))) for p in net.parameters())

        # TrainStep's batch split: every rank draws the same global pixel batch, takes its contiguous shard
        import plnerf_b200.train as T
        gen = torch.Generator(device="cpu")
        gen.manual_seed(7)
        pix = T.sample_pixels(40, 50, 101, "cpu", gen)
        lo3, hi3 = D.shard_bounds(pix.shape[0])
        gathered = D.gather_rays(pix[lo3:hi3].float(), pix.shape[0])
        assert torch.equal(gathered, pix.float())          # same draw everywhere, shards tile it in order
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_sharding_and_allreduce_world2():
    ws = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, ws, port, q)) for r in range(ws)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(ws)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


def test_shard_bounds_cover_everything():
    import plnerf_b200.dist as D
    for n in (0, 1, 7, 640000, 1024):
        for ws in (1, 2, 3, 4, 8):
            spans = [D.shard_bounds(n, r, ws) for r in range(ws)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


# ------------------------------------------------------------------------------------------------
# TrainStep across two ranks: the W-rank job is the 1-rank job (host logic + the one exchange step, over gloo)
# ------------------------------------------------------------------------------------------------
def _fake_pack(H, W, K, pose, pix, ndc, near, far, use_viewdirs):
    r = torch.zeros(pix.shape[0], 11)
    r[:, 0], r[:, 1] = (pix % W).float() / W, (pix // W).float() / H
    return r


def _fake_forward(cfg, rays):
    """Stand-in for autograd.forward_stashed (no GPU here): a small differentiable function of the rays and of each
    network's first layer; the ray ids the step hands over are recorded through cfg['ray_id_offset']."""
    with torch.enable_grad():
        x = torch.cat([rays[:, :3]] * 21, -1)
        rgb = torch.sigmoid(cfg["net_f"].pts_linears[0](x)[:, :3])
        rgb0 = torch.sigmoid(cfg["net_c"].pts_linears[0](x)[:, :3])
    z = torch.zeros(rays.shape[0])
    _fake_forward.offsets.append((cfg["ray_id_offset"], rays.shape[0]))
    return (rgb.detach(), z, z, z, None, rgb0.detach(), z, z, z, z), (rgb, rgb0), (None, None)


_fake_forward.offsets = []


def _fake_backward(cfg, saved, stashes, g_fine, g_coarse, grads_c, grads_f):
    for out, g, net, grads in ((saved[0], g_fine[0], cfg["net_f"], grads_f), (saved[1], g_coarse[0], cfg["net_c"], grads_c)):
        got = torch.autograd.grad(out, list(net.parameters()), g, allow_unused=True)
        for (k, _), gk in zip(net.named_parameters(), got):
            if gk is not None:
                grads[k].add_(gk)


def _run_train_steps(n_steps):
    import plnerf_b200.train as T
    from plnerf_b200 import autograd as AG, ops
    from plnerf_b200.run_nerf_helpers import NeRF
    from util import fake_mse_loss_grad, fake_adam_step
    ops.pack_pixel_rays, ops.invalidate_packed = _fake_pack, (lambda net: None)
    ops.mse_loss_grad, ops.adam_step = fake_mse_loss_grad, fake_adam_step
    AG.forward_stashed, AG.backward_stashed = _fake_forward, _fake_backward

    def fake_train_rays_mse(cfg, rays, target, pix, scale, sqerr, grads_c, grads_f):   # the fused entry = the three in sequence
        outs, saved, stashes = _fake_forward(cfg, rays)
        g, g0 = fake_mse_loss_grad(outs[0], outs[5], target, scale, sqerr, pix=pix)
        _fake_backward(cfg, saved, stashes, (g, None, None, None), (g0, None, None, None), grads_c, grads_f)
    AG.train_rays_mse = fake_train_rays_mse

    class HostOnlyStep(T.TrainStep):
        def _check_device(self):
            pass
    torch.manual_seed(0)
    mk = lambda: NeRF(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=[4], use_viewdirs=True)
    net_c, net_f = mk(), mk()
    kw = dict(network_fn=net_c, network_fine=net_f, N_samples=8, N_importance=8, perturb=1., white_bkgd=True,
              raw_noise_std=0., mode="linear", color_mode="midpoint", use_viewdirs=True, ndc=False, near=2., far=6., seed=5)
    step = HostOnlyStep(16, 16, np.eye(3), kw, N_rand=65, lrate=5e-4, coarse_lrate=5e-4, lrate_decay=250, seed=3)
    target = torch.rand(16, 16, 3, generator=torch.Generator().manual_seed(1))
    losses = []
    for i in range(n_steps):
        losses.append(step(target, torch.eye(4)[:3], i)["loss"].clone())
    return step.flat_params.clone(), torch.stack(losses)


def _train_worker(rank, ws, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    import plnerf_b200.dist as D
    try:
        params, losses = _run_train_steps(3)
        # shards: rank r renders the contiguous range shard_bounds(65) of every global batch, ids are global
        lo, hi = D.shard_bounds(65)
        assert _fake_forward.offsets == [(lo, hi - lo)] * 3, _fake_forward.offsets
        dist.all_reduce(losses)                              # each rank holds its share of the global mean
        gathered = [torch.empty_like(params) for _ in range(ws)]
        dist.all_gather(gathered, params)
        assert all(torch.equal(g, gathered[0]) for g in gathered)        # replicas stay identical
        if rank == 0:
            real = D.world
            D.world = lambda: (0, 1)                         # the same job on one rank
            _fake_forward.offsets.clear()
            try:
                params1, losses1 = _run_train_steps(3)
            finally:
                D.world = real
            assert _fake_forward.offsets == [(0, 65)] * 3
            assert torch.allclose(losses, losses1, rtol=1e-5, atol=0), (losses, losses1)
            # Adam turns a ~0 gradient's rounding into +-lr; everything else agrees to fp32 sum-order noise
            diff = (params - params1).abs()
            assert float((diff > 1e-6).float().mean()) < 1e-3 and float(diff.max()) <= 3 * 2 * 5e-4, float(diff.max())
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_train_step_world2_equals_single_rank():
    ws = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_train_worker, args=(r, ws, port, q)) for r in range(ws)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(ws)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res
