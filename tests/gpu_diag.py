"""Bring-up diagnostics for the tcgen05 path (not collected by pytest).  Run on the GPU box:
    timeout 120 python tests/gpu_diag.py gemm      # descriptor encodings: which (lbo,sbo) is right
    timeout 120 python tests/gpu_diag.py mlp       # fused MLP vs oracle on a few hundred rows
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")]
import plnerf_b200  # noqa: E402
from plnerf_b200 import _lib as L, ops  # noqa: E402
import plnerf_oracle as O  # noqa: E402


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def gemm():
    rs = np.random.RandomState(0)
    for a_mode in (0, 1):
        for (lbo, sbo) in ((2048, 128), (128, 2048)):
            for N, K in ((128, 16), (128, 64), (256, 256)):
                A = rs.randn(128, K).astype(np.float32)
                B = rs.randn(N, K).astype(np.float32)
                ref = O.bf16_round(A).astype(np.float64) @ O.bf16_round(B).astype(np.float64).T
                D = torch.zeros((128, N), device="cuda")
                dA, dB = dev(A), dev(B)
                rc = L.debug_lib().plnerf_debug_umma_gemm_ex(dA.data_ptr(), dB.data_ptr(), N, K, a_mode, lbo, sbo,
                                                       D.data_ptr(), None)
                torch.cuda.synchronize()
                out = D.cpu().numpy()
                err = np.abs(out - ref).max() / np.abs(ref).max()
                print(f"a_mode={a_mode} lbo={lbo} sbo={sbo} N={N} K={K} rc={rc} relerr={err:.3e} "
                      f"out[0,:4]={out[0,:4]} ref[0,:4]={ref[0,:4]}", flush=True)


def mlp():
    from util import case_params, load_golden, oracle_net_kw
    from plnerf_b200.run_nerf_helpers import NeRF
    for name in ("lego_linear_mid", "lego_left_noise_lindisp"):
        g = load_golden(name)
        cfg, kw, pc, pf = case_params(name)
        net = NeRF(D=kw["D"], W=kw["W"], input_ch=kw["input_ch"], input_ch_views=kw["input_ch_views"],
                   output_ch=kw["output_ch"], skips=list(kw["skips"]), use_viewdirs=kw["use_viewdirs"])
        net.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in pc.items()})
        net = net.cuda()
        for prec in ("bf16", "bf16x3"):
            with torch.no_grad():
                raw = ops.network_query(net, dev(g["ray_batch"]), dev(g["z_vals0"]), precision=prec)
            torch.cuda.synchronize()
            raw = raw.cpu().numpy()
            ref = g["raw0"][..., :raw.shape[-1]]
            scale = np.abs(ref).reshape(-1, ref.shape[-1]).max(0)
            err = (np.abs(raw - ref) / scale).reshape(-1, ref.shape[-1]).max(0)
            print(f"{name} {prec}: per-channel max err / scale = {err}  raw[0,0]={raw[0,0]} ref[0,0]={ref[0,0]}", flush=True)


def gemm_mn():
    rs = np.random.RandomState(0)
    for N, K in ((128, 16), (256, 64), (64, 128), (32, 128)):
        for (lbo, sbo) in ((128, (K // 8) * 128),):
            X = rs.randn(K, 128).astype(np.float32)
            Y = rs.randn(K, N).astype(np.float32)
            ref = O.bf16_round(X).astype(np.float64).T @ O.bf16_round(Y).astype(np.float64)
            D = torch.zeros((128, N), device="cuda")
            dX, dY = dev(X), dev(Y)
            rc = L.debug_lib().plnerf_debug_umma_gemm_mn(dX.data_ptr(), dY.data_ptr(), N, K, lbo, sbo, D.data_ptr(), None)
            torch.cuda.synchronize()
            out = D.cpu().numpy()
            err = np.abs(out - ref).max() / np.abs(ref).max()
            print(f"MN-major N={N} K={K} lbo={lbo} sbo={sbo} rc={rc} relerr={err:.3e} out[0,:3]={out[0,:3]} ref[0,:3]={ref[0,:3]}",
                  flush=True)


def mmarate2():
    for mode, name in ((3, "SS N=256 none (old loop)"), (6, "SS N=256 none"), (7, "SS N=256 A+B sw128"), (8, "SS N=256 A sw128"),
                       (9, "SS N=256 B sw128"), (10, "SS N=128 A+B sw128"), (11, "SS N=256 none, A fixed"), (12, "SS N=256 none, B fixed"),
                       (13, "SS N=256 none + 16 warps polling an mbarrier"), (14, "SS N=256 none + 4 warps polling")):
        out = torch.zeros(148, dtype=torch.int64, device="cuda")
        iters = 200
        for rep in range(2):
            L.check(L.debug_lib().plnerf_debug_mma_rate(mode, iters, 148, out.data_ptr(), None))
            torch.cuda.synchronize()
        cyc = out.cpu().numpy().astype(np.float64) / (iters * 16)
        print(f"{name}: cycles/MMA mean={cyc.mean():.1f} min={cyc.min():.1f} max={cyc.max():.1f}", flush=True)


def alurate():
    for mode, name in ((20, "cvt.rn.relu.bf16x2.f32"), (23, "cvt.rn.bf16x2.f32"), (21, "fmax+iadd round-half-up + prmt"), (22, "add.f32x2")):
        out = torch.zeros(148, dtype=torch.int64, device="cuda")
        iters = 2000
        for rep in range(2):
            L.check(L.debug_lib().plnerf_debug_mma_rate(mode, iters, 148, out.data_ptr(), None))
            torch.cuda.synchronize()
        cyc = out.cpu().numpy().astype(np.float64) / (iters * 16)
        print(f"{name}: cycles per pair-op per warp (4 warps/SM, 1 per SMSP) = {cyc.mean():.2f}", flush=True)


def mmarate():
    for grid in (1, 148):
        for mode, name in ((0, "TS N=128"), (1, "TS N=256"), (2, "SS N=128"), (3, "SS N=256"),
                           (4, "TS N=128, 8/batch + commit + wait(prev)"), (5, "TS N=128 issue-return time of 16 MMAs")):
            out = torch.zeros(grid, dtype=torch.int64, device="cuda")
            iters = 200
            for rep in range(2):
                L.check(L.debug_lib().plnerf_debug_mma_rate(mode, iters, grid, out.data_ptr(), None))
                torch.cuda.synchronize()
            cyc = out.cpu().numpy().astype(np.float64) / (iters * 16)
            print(f"grid={grid} {name}: cycles/MMA mean={cyc.mean():.1f} min={cyc.min():.1f} max={cyc.max():.1f}", flush=True)


if __name__ == "__main__":
    {"gemm": gemm, "mlp": mlp, "mmarate": mmarate, "mmarate2": mmarate2, "alurate": alurate, "gemm_mn": gemm_mn, "train": lambda: None}[sys.argv[1]]()


def train_diag():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import test_gpu_train as T
    from util import case_params, load_golden
    n, S = 40, 96
    cfg, kw, pc, pf = case_params("lego_linear_mid")
    net = T.make_net(kw, pc)
    rs = np.random.RandomState(n)
    g = load_golden("lego_linear_mid")
    rb = np.tile(g["ray_batch"], (n // g["ray_batch"].shape[0] + 1, 1))[:n]
    rays = dev(rb)
    z = torch.sort(torch.rand(n, S, device="cuda") * 4 + 2, -1)[0]
    g_raw = dev(rs.randn(n, S, 4).astype(np.float32))
    with torch.no_grad():
        raw, stash = ops.network_query_train(net, rays, z)
        grads = ops.network_query_bwd(net, g_raw, stash, n, S)
    emu = T.emulated_backward(net, rays, z, g_raw)
    for k, p in net.named_parameters():
        a, b = grads[k].double().flatten(), emu[k].double().flatten()
        rel = (a - b).norm().item() / b.norm().item()
        print(f"EMU {k:28s} rel={rel:.5f}", flush=True)
    ref_raw = T.torch_ref_query(net, rays, z)
    (ref_raw * g_raw).sum().backward()
    for k, p in net.named_parameters():
        a, b = grads[k].double().flatten(), p.grad.double().flatten()
        rel = (a - b).norm().item() / b.norm().item()
        cos = torch.dot(a, b).item() / (a.norm().item() * b.norm().item() + 1e-300)
        print(f"{k:28s} rel={rel:.4f} cos={cos:.6f} |ref|={b.norm().item():.4e} |got|={a.norm().item():.4e}", flush=True)


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "train":
    train_diag()
