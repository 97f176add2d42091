"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: ray sharding, global-ray-id offsets,
output gather and the single flat gradient all-reduce."""
import os
import sys

import numpy as np
import pytest
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
