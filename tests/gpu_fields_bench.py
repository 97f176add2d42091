"""Throughput of the mesh extractor's density grid on a B200 (python tests/gpu_fields_bench.py [R ...]):
plnerf_b200.nerf_extract_mesh.extract_fields on an R^3 grid, device time of the grid query (CUDA events) and wall time
of the whole call including the final device->host copy of the grid.  Prints one JSON line per resolution."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import plnerf_b200  # noqa: E402
from plnerf_b200 import nerf_extract_mesh as NM, synth  # noqa: E402
from plnerf_b200.run_nerf_helpers import NeRF  # noqa: E402

KW = dict(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=(4,), use_viewdirs=True)
FLOP_PER_EVAL = 1186816


def main():
    res = [int(a) for a in sys.argv[1:]] or [256, 512]
    net = NeRF(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=[4], use_viewdirs=True)
    net.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in synth.nerf_params(3, **KW).items()})
    net = net.cuda()
    bmin, bmax = [-1.2, -1.2, -1.2], [1.2, 1.2, 1.2]
    NM.extract_fields(bmin, bmax, 64, None, net)          # warm-up (packs the weights, loads the kernels)
    for R in res:
        X, Y, Z = (NM._axis(bmin[k], bmax[k], R) for k in range(3))
        out = torch.empty((R, R, R), device="cuda")
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        with torch.no_grad():
            NM.query_density_grid(net, X, Y, Z, out=out)
        e1.record()
        torch.cuda.synchronize()
        dev_ms = e0.elapsed_time(e1)
        walls = []
        for _ in range(2):      # the first call at a new size also pays for the pinned host allocation (cached afterwards)
            u = None
            t0 = time.perf_counter()
            u = NM.extract_fields(bmin, bmax, R, None, net)
            walls.append((time.perf_counter() - t0) * 1e3)
        wall_ms = walls[1]
        n = R ** 3
        print(json.dumps({"resolution": R, "points": n, "query_device_ms": dev_ms,
                          "points_per_s_device": n / dev_ms * 1e3,
                          "algorithmic_tflops_device": n * FLOP_PER_EVAL / dev_ms / 1e9,
                          "extract_fields_wall_ms": wall_ms, "extract_fields_first_call_wall_ms": walls[0], "points_per_s_wall": n / wall_ms * 1e3,
                          "grid_bytes_d2h": int(u.nbytes), "nonzero_frac": float(np.mean(u > 0)), "precision": "bf16"}))


if __name__ == "__main__":
    main()
