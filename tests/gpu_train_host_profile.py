"""Host-side profile of TrainStep (cProfile over 200 iterations): where the ~0.75 ms of issue time per iteration goes.
    python tests/gpu_train_host_profile.py"""
import cProfile
import os
import pstats
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import gpu_train_step_target as G  # noqa: E402
from plnerf_b200 import synth, train as T  # noqa: E402


def main():
    net_c, net_f = G.mk(11), G.mk(12)
    K = synth.intrinsics(G.H, G.W, 0.5 * G.W / np.tan(0.5 * 0.6911112070083618))
    kw = dict(network_query_fn=None, network_fn=net_c, network_fine=net_f, N_samples=128, N_importance=64, perturb=1.0,
              white_bkgd=True, raw_noise_std=0., mode="linear", color_mode="midpoint", use_viewdirs=True, ndc=False,
              near=2., far=6.)
    n_rand = int(sys.argv[1]) if len(sys.argv) > 1 else 128          # small batch: the device is never the bottleneck
    step = T.TrainStep(G.H, G.W, K, kw, N_rand=n_rand, lrate=5e-4, coarse_lrate=5e-4, lrate_decay=500, seed=1)
    target = torch.rand(G.H, G.W, 3, device="cuda")
    pose = torch.from_numpy(synth.pose_spherical(-180.0, -30.0, 4.0)[:3, :4].astype(np.float32).copy()).cuda()
    for i in range(10):
        step(target, pose, i)
    torch.cuda.synchronize()
    pr = cProfile.Profile()
    pr.enable()
    for i in range(10, 210):
        step(target, pose, i)
    pr.disable()
    torch.cuda.synchronize()
    st = pstats.Stats(pr)
    st.sort_stats("cumulative").print_stats(28)
    st.sort_stats("tottime").print_stats(18)


if __name__ == "__main__":
    main()
