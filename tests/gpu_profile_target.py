"""Small driver for ncu: a few fused-MLP launches at the benchmark shape (32768 rays x 192)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import plnerf_b200  # noqa: E402
from plnerf_b200 import ops, synth  # noqa: E402
from plnerf_b200.run_nerf_helpers import NeRF  # noqa: E402

prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
n_launch = int(sys.argv[2]) if len(sys.argv) > 2 else 3
n, S = 32768, 192
kw = dict(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=(4,), use_viewdirs=True)
net = NeRF(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=[4], use_viewdirs=True)
net.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in synth.nerf_params(1, **kw).items()})
net = net.cuda()
ro, rd, K, _ = synth.lego_rays(n, seed=1)
vd = rd / np.linalg.norm(rd, axis=-1, keepdims=True)
rays = torch.from_numpy(np.concatenate([ro, rd, np.full((n, 1), 2, np.float32), np.full((n, 1), 6, np.float32), vd], -1)).cuda()
z = torch.sort(torch.rand(n, S, device="cuda") * 4 + 2, -1)[0]
with torch.no_grad():
    for i in range(n_launch):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        raw = ops.network_query(net, rays, z, precision=prec)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print(f"{prec} launch {i}: {ms:.3f} ms  {n*S*1186816/ms/1e9:.1f} TFLOP/s", flush=True)

