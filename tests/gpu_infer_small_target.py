"""Launch list of ONE inference render at the training shape (1024 rays, 128 + 64 samples): how long the two k_mlp3 launches
take without the training stash (reference point for the stash forward).  Run under
ncu --metrics gpu__time_duration.sum --profile-from-start off."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import plnerf_b200  # noqa: E402
from plnerf_b200 import ops, synth, run_plnerf as RP  # noqa: E402
from plnerf_b200.run_nerf_helpers import NeRF  # noqa: E402

KW = dict(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=(4,), use_viewdirs=True)


def mk(seed):
    net = NeRF(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=[4], use_viewdirs=True)
    net.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in synth.nerf_params(seed, density_boost=False, **KW).items()})
    return net.cuda()


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    net_c, net_f = mk(11), mk(12)
    ro, rd, K, (H, W, focal) = synth.lego_rays(n, seed=3)
    rays = torch.stack([torch.from_numpy(ro), torch.from_numpy(rd)]).cuda()
    kw = dict(ndc=False, near=2., far=6., use_viewdirs=True, network_query_fn=None, network_fn=net_c, network_fine=net_f,
              N_samples=128, N_importance=64, perturb=1.0, white_bkgd=True, mode="linear", color_mode="midpoint")
    with torch.no_grad():
        for _ in range(3):
            RP.render(H, W, K, rays=rays, **kw)
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStart()
        RP.render(H, W, K, rays=rays, **kw)
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStop()


if __name__ == "__main__":
    main()
