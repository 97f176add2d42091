"""In-situ kernel timeline of render() at the bench shape (two 32 768-ray chunks, 64+128 samples; torch.profiler / CUPTI
activity records, no ncu serialisation):   python tests/gpu_render_timeline.py > gpurun_out/render_timeline.txt"""
import os
import sys

import numpy as np
import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import plnerf_b200  # noqa: E402,F401
from plnerf_b200 import run_plnerf as RP, synth  # noqa: E402
from plnerf_b200.run_nerf_helpers import NeRF  # noqa: E402

KW = dict(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=(4,), use_viewdirs=True)


def mk(seed):
    net = NeRF(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=[4], use_viewdirs=True)
    net.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in synth.nerf_params(seed, **KW).items()})
    return net.cuda()


def main():
    n = 65536
    ro, rd, K, (H, W, focal) = synth.lego_rays(n, seed=3)
    rays = torch.stack([torch.from_numpy(ro), torch.from_numpy(rd)]).cuda()
    kw = dict(ndc=False, near=2., far=6., use_viewdirs=True, network_query_fn=None, network_fn=mk(1), network_fine=mk(2),
              N_samples=64, N_importance=128, perturb=1.0, white_bkgd=True, mode="linear", color_mode="midpoint")
    with torch.no_grad():
        for _ in range(3):
            RP.render(H, W, K, chunk=32768, rays=rays, **kw)
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            RP.render(H, W, K, chunk=32768, rays=rays, **kw)
            torch.cuda.synchronize()
    evs = sorted((e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA), key=lambda e: e.time_range.start)
    t0 = evs[0].time_range.start
    tot = evs[-1].time_range.end - t0
    mine = sum(e.time_range.end - e.time_range.start for e in evs if "k_mlp3" in e.name)
    print(f"# render of {n} rays: {tot:.1f} us, k_mlp3 share {mine / tot:.4f}")
    for e in evs:
        print(f"{e.time_range.start - t0:9.1f} {e.time_range.end - e.time_range.start:8.1f}  {e.name[:90]}")


if __name__ == "__main__":
    main()
