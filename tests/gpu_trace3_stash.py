"""Timeline of one tile pair of k_mlp3 at the TRAINING shape (1 024 rays x 192 samples), inference vs stash forward
(developer library: PLNERF_DEBUG_LIB=1 python tests/gpu_trace3_stash.py [train|infer]).  Per layer of tile X, column half 0:
when the epilogue group reaches the layer, how long it waits for each accumulator half, how long its own work takes; and
the issuer's waits per program entry."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import plnerf_b200  # noqa: E402
from plnerf_b200 import ops, synth, _lib as L  # noqa: E402
from plnerf_b200.run_nerf_helpers import NeRF  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "train"
n, S = 1024, 192
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
    for i in range(3):
        trace.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if mode == "train":
            raw, stash = ops.network_query_train(net, rays, z)
        else:
            raw = ops.network_query(net, rays, z, precision="bf16")
        e1.record()
        torch.cuda.synchronize()
        print(f"{mode} launch {i}: {e0.elapsed_time(e1) * 1e3:.1f} us", flush=True)
raw_t = trace.cpu().numpy()
t = raw_t[:4 * 256 * 2].reshape(4, 256, 2)
issue = raw_t[4 * 256 * 2:].view(np.uint32)
for region in (0, 2):
    ev = [(int(c), int(code)) for c, code in t[region] if code != 0]
    if not ev:
        continue
    t0 = ev[0][0]
    what = {2: "top", 3: "d_full", 5: "loaded", 7: "conv", 8: "stored", 4: "arrived"}
    print(f"--- tile {'XY'[region // 2]} column half {region % 2}: cycles since the pair's first stamp")
    line = []
    for c, code in ev:
        kind, sub = code // 1000, code % 1000
        l, h = sub // 10, sub % 10
        line.append(f"l{l}h{h}:{what.get(kind, kind)}@{c - t0}")
    print("  " + "  ".join(line))
    print(f"  pair total: {ev[-1][0] - t0} cycles")
base32 = int(min(int(c) for r in range(4) for c, code in t[r] if code != 0)) & 0xFFFFFFFF
for tile in range(2):
    iss = issue[256 * tile: 256 * (tile + 1)]
    k = 0
    waits = []
    while 3 * k + 2 < len(iss) and iss[3 * k] != 0:
        a, b, c = (int(iss[3 * k + j]) - base32 for j in range(3))
        waits.append((a, b - a, c - b))
        k += 1
    print(f"issuer {'XY'[tile]}: {k} entries; sum of waits {sum(w[1] for w in waits)}, sum of issue {sum(w[2] for w in waits)}; "
          f"span {waits[-1][0] + waits[-1][1] + waits[-1][2] - waits[0][0] if waits else 0}")
    print("   (start, wait, issue) per entry: " + " ".join(f"({a},{w},{i})" for a, w, i in waits))
