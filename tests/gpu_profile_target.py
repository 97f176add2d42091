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
trace = None
if os.environ.get("PLNERF_TRACE"):
    from plnerf_b200 import _lib as L
    trace = torch.zeros(4 * 256 * 2, dtype=torch.int64, device="cuda")
    L.check(L.lib().plnerf_debug_set_trace(trace.data_ptr()))
with torch.no_grad():
    for i in range(n_launch):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        raw = ops.network_query(net, rays, z, precision=prec)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print(f"{prec} launch {i}: {ms:.3f} ms  {n*S*1186816/ms/1e9:.1f} TFLOP/s", flush=True)

if trace is not None:
    tt = trace.cpu().numpy()
    if tt[-1] > 0:
        print(f"kernel: {tt[-2]} SM cycles in {tt[-1]} ns -> {tt[-2] / tt[-1]:.3f} GHz; cycles per tile = {tt[-2] / (n * S / 128 / 148):.0f}")
    tt[-2:] = 0
    t = tt.reshape(4, 256, 2)
    ev = [(int(c), int(code), r) for r in range(4) for c, code in t[r] if code != 0]
    ev.sort()
    t0 = ev[0][0]
    names = {0: "MMA ", 1: "EPI0", 2: "EPI1", 3: "TMA "}
    for c, code, r in ev:
        kind = code // 1000
        if os.environ.get("PLNERF_MLP_KERNEL") != "v1":
            sub = code % 1000
            l, t, st = sub // 100, (sub % 100) // 50, sub % 50
            what = {1: "slot ready (a_ready seen)", 2: "stage full seen, issuing", 3: "d_full seen", 4: "a_ready signalled",
                    6: "stage empty seen, TMA issue", 7: "tmem loaded", 8: "activations stored", 9: "issue returned", 10: "d_full committed", 11: "item top"}[kind]
            print(f"{c - t0:8d}  {names[r]}  l={l} slot={t} st={st:2d}  {what}")
            continue
        desc = {1: "batch top   entry=%d" % (code % 1000),
                2: "batch issue entry=%d" % (code % 1000),
                3: "d_full seen l=%d h=%d" % ((code % 1000) // 10, code % 10),
                4: "arrived     l=%d h=%d" % ((code % 1000) // 10, code % 10),
                5: "issue ret   entry=%d" % (code % 1000),
                6: "tma issue   l=%d h=%d st=%d" % ((code % 1000) // 100, (code % 100) // 10, code % 10),
                7: "pe step %d" % (code % 1000)}[kind]
        print(f"{c - t0:8d}  {names[r]}  {desc}")
