"""Timeline of one tile pair of k_mlp3 (developer library: run with PLNERF_DEBUG_LIB=1): clock64 stamps of
lane 0 of the first warp of each (tile, column half) epilogue group of block 0 on its third pair; argv[1] = region (tile*2 + column half)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import plnerf_b200  # noqa: E402
from plnerf_b200 import ops, synth, _lib as L  # noqa: E402
from plnerf_b200.run_nerf_helpers import NeRF  # noqa: E402

n, S = 32768, 192
kw = dict(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=(4,), use_viewdirs=True)
net = NeRF(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=[4], use_viewdirs=True)
net.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in synth.nerf_params(1, **kw).items()})
net = net.cuda()
ro, rd, K, _ = synth.lego_rays(n, seed=1)
vd = rd / np.linalg.norm(rd, axis=-1, keepdims=True)
rays = torch.from_numpy(np.concatenate([ro, rd, np.full((n, 1), 2, np.float32), np.full((n, 1), 6, np.float32), vd], -1)).cuda()
z = torch.sort(torch.rand(n, S, device="cuda") * 4 + 2, -1)[0]
trace = torch.zeros(4 * 256 * 2 + 256, dtype=torch.int64, device="cuda")
L.check(L.debug_lib().plnerf_debug_set_trace(trace.data_ptr()))
with torch.no_grad():
    for i in range(2):
        trace.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        raw = ops.network_query(net, rays, z, precision="bf16")
        e1.record()
        torch.cuda.synchronize()
        print(f"launch {i}: {e0.elapsed_time(e1):.3f} ms", flush=True)
raw_t = trace.cpu().numpy()
t = raw_t[:4 * 256 * 2].reshape(4, 256, 2)
issue = raw_t[4 * 256 * 2:].view(np.uint32)
ev = sorted((int(c), int(code), r) for r in range(4) for c, code in t[r] if code != 0)
t0 = ev[0][0]
what = {2: "top", 3: "d_full seen", 5: "loaded", 7: "converted", 8: "stored", 4: "arrived"}
only = int(sys.argv[1]) if len(sys.argv) > 1 else 0
for c, code, r in ev:
    if r != only:
        continue
    kind, sub = code // 1000, code % 1000
    l, h = sub // 10, sub % 10
    print(f"{c - t0:8d} tile={'XY'[r // 2]} ch={r % 2} l={l} h={h} {what.get(kind, kind)}")

# issue side: per program entry (start, waits done, stage issued), low 32 bits of the same SM clock
base32 = t0 & 0xFFFFFFFF
for tile in range(2):
    iss = issue[256 * tile: 256 * (tile + 1)]
    n = 0
    print(f"issuer {'XY'[tile]}: entry  start  waits_done  issued   (cycles relative to the first epilogue stamp)")
    while 3 * n + 2 < len(iss) and iss[3 * n] != 0:
        a, b, c = (int(iss[3 * n + k]) - base32 for k in range(3))
        print(f"  e{n:3d} {a:8d} {b:8d} {c:8d}   wait={b - a:5d} issue={c - b:5d}")
        n += 1
