"""In-situ kernel timeline of TrainStep (torch.profiler / CUPTI activity records: start, duration and stream of every kernel
of a few iterations, without ncu's serialisation and cache flushes):
    python tests/gpu_train_timeline.py [n_rand]  > gpurun_out/train_timeline.txt"""
import os
import sys

import numpy as np
import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import gpu_train_step_target as G  # noqa: E402
from plnerf_b200 import synth, train as T  # noqa: E402


def main():
    n_rand = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    net_c, net_f = G.mk(11), G.mk(12)
    K = synth.intrinsics(G.H, G.W, 0.5 * G.W / np.tan(0.5 * 0.6911112070083618))
    kw = dict(network_query_fn=None, network_fn=net_c, network_fine=net_f, N_samples=128, N_importance=64, perturb=1.0,
              white_bkgd=True, raw_noise_std=0., mode="linear", color_mode="midpoint", use_viewdirs=True, ndc=False,
              near=2., far=6.)
    step = T.TrainStep(G.H, G.W, K, kw, N_rand=n_rand, lrate=5e-4, coarse_lrate=5e-4, lrate_decay=500, seed=1)
    target = torch.rand(G.H, G.W, 3, device="cuda")
    pose = torch.from_numpy(synth.pose_spherical(-180.0, -30.0, 4.0)[:3, :4].astype(np.float32).copy()).cuda()
    for i in range(20):
        step(target, pose, i)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for i in range(20, 26):
            step(target, pose, i)
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    evs.sort(key=lambda e: e.time_range.start)
    t0 = evs[0].time_range.start
    # iterations are delimited by k_pack_rays (the first plnerf kernel of an iteration)
    starts = [k for k, e in enumerate(evs) if "k_pack_rays" in e.name]
    lo, hi = starts[2], starts[4]
    print(f"# two iterations: {(evs[hi].time_range.start - evs[lo].time_range.start) / 2:.1f} us per iteration")
    prev_end = None
    for e in evs[lo:hi]:
        s, d = e.time_range.start - evs[lo].time_range.start, e.time_range.end - e.time_range.start
        gap = "" if prev_end is None else f"{e.time_range.start - prev_end:+7.1f}"
        prev_end = max(prev_end or 0, e.time_range.end)
        print(f"{s:9.1f} {d:8.1f} gap {gap:>8}  {e.name[:70]}")


if __name__ == "__main__":
    main()
