"""Profile target for the device-side training step (plnerf_b200.train.TrainStep) at the bench shape
(N_rand=1024, 128+64 samples, 800x800 image):

    python tests/gpu_train_step_target.py                      # issue time vs device time per iteration
    ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
        --log-file gpurun_out/train_step_launches.csv python tests/gpu_train_step_target.py ncu

With `ncu` as argument, two iterations are bracketed by cudaProfilerStart/Stop (the launch list of the step)."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import plnerf_b200  # noqa: E402
from plnerf_b200 import ops, synth, train as T  # noqa: E402
from plnerf_b200.run_nerf_helpers import NeRF  # noqa: E402

KW = dict(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=(4,), use_viewdirs=True)
H = W = 800


def mk(seed):
    net = NeRF(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=[4], use_viewdirs=True)
    net.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in synth.nerf_params(seed, density_boost=False, **KW).items()})
    return net.cuda()


def main():
    under_ncu = len(sys.argv) > 1 and sys.argv[1] == "ncu"
    net_c, net_f = mk(11), mk(12)
    K = synth.intrinsics(H, W, 0.5 * W / np.tan(0.5 * 0.6911112070083618))
    kw = dict(network_query_fn=None, network_fn=net_c, network_fine=net_f, N_samples=128, N_importance=64, perturb=1.0,
              white_bkgd=True, raw_noise_std=0., mode="linear", color_mode="midpoint", use_viewdirs=True, ndc=False,
              near=2., far=6.)
    step = T.TrainStep(H, W, K, kw, N_rand=1024, lrate=5e-4, coarse_lrate=5e-4, lrate_decay=500, seed=1)
    target = torch.rand(H, W, 3, device="cuda")
    pose = torch.from_numpy(synth.pose_spherical(-180.0, -30.0, 4.0)[:3, :4].astype(np.float32).copy()).cuda()
    for i in range(4):
        step(target, pose, i)
    torch.cuda.synchronize()
    if under_ncu:
        torch.cuda.cudart().cudaProfilerStart()
        for i in range(4, 6):
            step(target, pose, i)
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStop()
        return
    iters = int(os.environ.get("PLNERF_ITERS", "40"))
    l0 = ops.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for i in range(4, 4 + iters):
        step(target, pose, i)
    e1.record()
    issue_ms = (time.perf_counter() - t0) * 1e3 / iters       # host time to ISSUE an iteration (no sync inside)
    torch.cuda.synchronize()
    print(json.dumps({"iters": iters, "device_ms_per_iter": e0.elapsed_time(e1) / iters, "host_issue_ms_per_iter": issue_ms,
                      "plnerf_launches_per_iter": (ops.launch_count() - l0) / iters}))


if __name__ == "__main__":
    main()
