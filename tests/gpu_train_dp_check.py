"""Data-parallel training step on real GPUs: the W-rank job is the 1-rank job.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/gpu_train_dp_check.py

Every rank builds the same two networks and a plnerf_b200.train.TrainStep over the same global pixel batches
(N_rand = 1024 split into contiguous shards, Philox draws keyed by global ray id, loss gradient scaled by the
global batch, ONE NCCL all-reduce (sum) of the flat gradient buffer per step).  Rank 0 then repeats the same
steps alone (world forced to 1) from the same initial weights and compares: loss per step, the flat gradient of
the first step and the parameters after the last step.  Prints one JSON line; exit code 1 on disagreement.
Not bit-exact by construction: the per-rank partial sums are added in a different order than the single-rank
atomics, and Adam turns a sign flip of a ~0 gradient into a +-lr difference."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import plnerf_b200  # noqa: E402
from plnerf_b200 import dist as PD, synth, train as T  # noqa: E402
from plnerf_b200.run_nerf_helpers import NeRF  # noqa: E402

KW = dict(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=(4,), use_viewdirs=True)
H = W = 200
STEPS, N_RAND = 6, 1024


def make_net(seed, dev):
    net = NeRF(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=[4], use_viewdirs=True)
    net.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in synth.nerf_params(seed, density_boost=False, **KW).items()})
    return net.to(dev)


def run(dev):
    net_c, net_f = make_net(81, dev), make_net(82, dev)
    K = np.array([[250.0, 0, 0.5 * W], [0, 250.0, 0.5 * H], [0, 0, 1]])
    kw = dict(network_query_fn=None, network_fn=net_c, network_fine=net_f, N_samples=64, N_importance=64, perturb=1.0,
              white_bkgd=True, raw_noise_std=0., mode="linear", color_mode="midpoint", use_viewdirs=True, ndc=False,
              near=2., far=6., seed=4242)
    step = T.TrainStep(H, W, K, kw, N_rand=N_RAND, lrate=5e-4, coarse_lrate=5e-4, lrate_decay=250, seed=7)
    gen = torch.Generator(device=dev)
    gen.manual_seed(3)
    target = torch.rand(H, W, 3, device=dev, generator=gen)
    pose = torch.from_numpy(synth.pose_spherical(30.0, -30.0, 4.0)[:3, :4].astype(np.float32).copy()).to(dev)
    losses, grad0 = [], None
    for i in range(STEPS):
        out = step(target, pose, i)
        loss = out["loss"].clone()
        if PD.world()[1] > 1:
            dist.all_reduce(loss)                      # each rank holds its share of the global mean
        losses.append(float(loss))
        if i == 0:
            grad0 = step.bucket.flat.clone()
    return losses, grad0, step.flat_params.clone()


def main():
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    losses_dp, grad_dp, params_dp = run(dev)
    ok = True
    if world > 1:
        # every rank must hold the same parameters after the all-reduced steps
        ref = params_dp.clone()
        dist.broadcast(ref, src=0)
        same = torch.tensor([float(torch.equal(ref, params_dp))], device=dev)
        dist.all_reduce(same, op=dist.ReduceOp.MIN)
        ranks_identical = bool(same.item())
        dist.barrier()
    else:
        ranks_identical = True
    if rank == 0:
        real_world = PD.world
        PD.world = lambda: (0, 1)                       # the same job on one rank
        try:
            losses_1, grad_1, params_1 = run(dev)
        finally:
            PD.world = real_world
        init = torch.cat([p.detach().flatten() for n in (make_net(82, dev), make_net(81, dev)) for p in n.parameters()])
        d_dp, d_1 = (params_dp - init).double(), (params_1 - init).double()
        rec = {"world": world, "steps": STEPS, "N_rand_global": N_RAND, "ranks_identical": ranks_identical,
               "loss_dp": losses_dp, "loss_single": losses_1,
               "loss_max_rel_diff": max(abs(a - b) / abs(b) for a, b in zip(losses_dp, losses_1)),
               "grad_step0_rel_diff": float((grad_dp - grad_1).norm() / grad_1.norm()),
               "update_cosine": float(torch.dot(d_dp, d_1) / (d_dp.norm() * d_1.norm())),
               "update_max_abs_diff": float((d_dp - d_1).abs().max())}
        ok = (ranks_identical and rec["loss_max_rel_diff"] < 2e-3 and rec["grad_step0_rel_diff"] < 5e-3
              and rec["update_cosine"] > 0.97)
        rec["ok"] = ok
        print(json.dumps(rec))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
