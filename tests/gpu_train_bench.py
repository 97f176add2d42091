"""Training-step timing on one B200: fwd (stash) + loss + backward + two Adam steps, like the reference loop
(run_plnerf.py:1283-1303).  python tests/gpu_train_bench.py [N_rand] [Ns] [Ni] [iters]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import plnerf_b200  # noqa: E402
from plnerf_b200 import ops, run_plnerf as RP, synth  # noqa: E402
from plnerf_b200.run_nerf_helpers import NeRF  # noqa: E402

N_rand = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
Ns = int(sys.argv[2]) if len(sys.argv) > 2 else 128
Ni = int(sys.argv[3]) if len(sys.argv) > 3 else 64
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 30
kw = dict(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=(4,), use_viewdirs=True)


def mk(seed):
    net = NeRF(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=[4], use_viewdirs=True)
    net.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in synth.nerf_params(seed, density_boost=False, **kw).items()})
    return net.cuda()


net_c, net_f = mk(1), mk(2)
opt_f = torch.optim.Adam(net_f.parameters(), lr=5e-4, fused=True)
opt_c = torch.optim.Adam(net_c.parameters(), lr=5e-4, fused=True)
ro, rd, K, (H, W, focal) = synth.lego_rays(None)
ro_t, rd_t = torch.from_numpy(ro).cuda(), torch.from_numpy(rd).cuda()
target_img = torch.rand(H * W, 3, device="cuda")


def step():
    idx = torch.randint(0, H * W, (N_rand,), device="cuda")
    rays = torch.stack([ro_t[idx], rd_t[idx]])
    rgb, disp, acc, extras = RP.render(H, W, K, chunk=1024 * 32, rays=rays, ndc=False, near=2., far=6., use_viewdirs=True,
                                       network_query_fn=None, network_fn=net_c, network_fine=net_f, N_samples=Ns,
                                       N_importance=Ni, perturb=1.0, white_bkgd=True, mode="linear", color_mode="midpoint",
                                       retraw=True)
    tgt = target_img[idx]
    loss = torch.mean((rgb - tgt) ** 2) + torch.mean((extras["rgb0"] - tgt) ** 2)
    opt_f.zero_grad(set_to_none=True); opt_c.zero_grad(set_to_none=True)
    loss.backward()
    opt_f.step(); opt_c.step()
    return loss


for _ in range(5):
    l0 = step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
e0.record()
for _ in range(iters):
    l = step()
e1.record()
torch.cuda.synchronize()
wall = time.perf_counter() - t0
ms = e0.elapsed_time(e1) / iters
rows = N_rand * (2 * Ns + Ni)
flop = rows * 3489024      # fwd + dX + dW MACs*2 per evaluation (1 744 512 MAC)
print(f"train N_rand={N_rand} Ns={Ns} Ni={Ni}: {ms:.3f} ms/iter device ({1000/ms:.1f} it/s), wall {wall/iters*1e3:.3f} ms/iter, "
      f"{N_rand/ms*1e3:.0f} rays/s, {flop/ms/1e9:.1f} TFLOP/s algorithmic, loss {l0.item():.4f} -> {l.item():.4f}", flush=True)
