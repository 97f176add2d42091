"""Diagnostic (not a test): per-ray error distribution of the smoke() configuration vs the oracle."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import plnerf_b200
from plnerf_b200 import run_plnerf as RP, synth
from plnerf_b200.run_nerf_helpers import NeRF
import plnerf_oracle as O

n, Ns, Ni = int(os.environ.get("N", 256)), 64, 128
kw = dict(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=(4,), use_viewdirs=True)
pc, pf = synth.nerf_params(1, **kw), synth.nerf_params(2, **kw)
def mk(p):
    net = NeRF(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=[4], use_viewdirs=True)
    net.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in p.items()})
    return net.cuda()
ro, rd, K, (H, W, focal) = synth.lego_rays(n, seed=3)
rs = np.random.RandomState(0)
t_rand = rs.rand(n, Ns).astype(np.float32)
u = rs.rand(n, Ni).astype(np.float32)
rays = torch.stack([torch.from_numpy(ro), torch.from_numpy(rd)]).cuda()
ref = O.render(H, W, K, ro, rd, ndc=False, near=2., far=6., use_viewdirs=True, t_rand=t_rand, u=u,
               params_coarse=pc, params_fine=pf, N_samples=Ns, mode="linear", color_mode="midpoint",
               N_importance=Ni, white_bkgd=True,
               net_kw=dict(D=8, skips=(4,), input_ch=63, input_ch_views=27, use_viewdirs=True))
for prec in ("bf16x3", "bf16"):
    with torch.no_grad():
        rgb, disp, acc, ex = RP.render(H, W, K, rays=rays, ndc=False, near=2., far=6., use_viewdirs=True,
                                       network_query_fn=None, network_fn=mk(pc), network_fine=mk(pf),
                                       N_samples=Ns, N_importance=Ni, perturb=1.0, white_bkgd=True,
                                       mode="linear", color_mode="midpoint", t_rand=torch.from_numpy(t_rand).cuda(),
                                       u=torch.from_numpy(u).cuda(), precision=prec, retraw=True)
    torch.cuda.synchronize()
    out = {"rgb_map": rgb, "disp_map": disp, "acc_map": acc}
    out.update(ex)
    print("==", prec)
    for k in ("rgb_map", "depth_map", "acc_map", "disp_map", "rgb0", "depth0", "acc0", "z_std", "raw"):
        if k not in ref or k not in out:
            continue
        a, b = out[k].cpu().numpy().astype(np.float64), ref[k].astype(np.float64)
        e = np.abs(a - b).reshape(a.shape[0], -1).max(1)
        srt = np.sort(e)[::-1]
        print(f"{k:10s} scale={np.abs(b).max():.3g} max={srt[0]:.3e} top5={srt[:5]} median={np.median(e):.2e} p99={np.percentile(e,99):.2e}")
