#!/usr/bin/env python
"""Benchmark of the PL-NeRF ray-rendering hot path on B200 (contract in the task brief / DESIGN.md).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--precision bf16|bf16x3]

Workload (BASELINE.json configs[1]): 800x800 Blender-lego-shaped synthetic rays (640 000 rays per
image), N_samples=64 + N_importance=128, PL ("linear") quadrature, midpoint colour, viewdirs, white
background, seeded random-init (density-boosted) coarse+fine 8x256 NeRF, reference chunking (32 768
rays per render_rays call).  One "step" = one full image through render_rays.  Weak scaling: every
rank renders its own image (different pose), no data-path collective.

Printed JSON (rank 0): `value` = rays/s with the packed rays already resident in HBM;
`e2e` = the same image through the public `render()` call from pinned HOST rays, with the H2D copy
of the rays and the D2H copy of rgb/disp/acc inside the timed region;
`roofline` = algorithmic MLP FLOPs / CUDA-event time of the k_mlp_fwd launches inside the timed
region, against the measured bf16 tensor peak; `cpu_baseline` = the numpy/torch-CPU oracle port of
the reference on a bounded ray sample on this box's host cores.
`--impl reference` times that same CPU port (the reference is pure PyTorch-CPU; its files do not
travel to the GPU box) on the same config with all host threads.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

H = W = 800
N_SAMPLES, N_IMPORTANCE = 64, 128
CHUNK = 1024 * 32
FLOP_PER_EVAL = 1186816            # 593 408 MAC, viewdirs network
NET_KW = dict(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=(4,), use_viewdirs=True)
METRIC = "rays/sec (64 coarse + 128 fine samples)"


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as f:
            d = json.load(f)
        return d.get("bf16_tflops_sustained", 1387.7), d.get("bf16_tflops", 1648.0), "measured (MEASURED_PEAKS.json)"
    return 1400.0, 1590.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.lines, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nme, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def build_rays(rank):
    from plnerf_b200 import synth
    theta = float(np.linspace(-180, 180, 41)[:-1][rank % 40])
    ro, rd, K, hwf = synth.lego_rays(None, H=H, W=W, theta=theta)
    return ro, rd, K


def oracle_render(ro, rd, K, pc, pf, seed=0):
    import plnerf_oracle as O
    n = ro.shape[0]
    rs = np.random.RandomState(seed)
    t_rand = rs.rand(n, N_SAMPLES).astype(np.float32)
    u = rs.rand(n, N_IMPORTANCE).astype(np.float32)
    return O.render(H, W, K, ro, rd, chunk=CHUNK, ndc=False, near=2., far=6., use_viewdirs=True, t_rand=t_rand, u=u,
                    params_coarse=pc, params_fine=pf, N_samples=N_SAMPLES, mode="linear", color_mode="midpoint",
                    N_importance=N_IMPORTANCE, white_bkgd=True,
                    net_kw=dict(D=8, skips=(4,), input_ch=63, input_ch_views=27, use_viewdirs=True))


def config_dict(args, extra=None):
    c = {"workload": "lego-shaped 800x800 synthetic rays (640000 rays/step), N_samples=64, N_importance=128, "
                     "mode=linear, color_mode=midpoint, use_viewdirs, white_bkgd, perturb=1, chunk=32768, "
                     "random-init density-boosted coarse+fine NeRF 8x256",
         "rays_per_step": H * W, "chunk": CHUNK, "parallelism": f"ray-sharded x{args.gpus} (one image per rank)",
         "l2_policy": "working set per step (raw [640000,192,4] fp32 = 1.97 GB + depths/weights) >> 126 MB L2; "
                      "no explicit flush"}
    if extra:
        c.update(extra)
    return c


def run_reference(args):
    """CPU arm: the oracle port of the reference on a bounded ray sample per step (rank 0 only)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from plnerf_b200 import synth
    cores = os.cpu_count()
    torch.set_num_threads(cores)
    pc, pf = synth.nerf_params(1, **NET_KW), synth.nerf_params(2, **NET_KW)
    ro, rd, K = build_rays(0)
    sample = args.cpu_rays
    idx = np.random.RandomState(0).choice(H * W, sample, replace=False)
    ro, rd = ro[idx], rd[idx]
    for _ in range(args.warmup):
        oracle_render(ro, rd, K, pc, pf)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle_render(ro, rd, K, pc, pf)
    dt = time.perf_counter() - t0
    v = sample * args.steps / dt
    desc = f"{sample} of the 640000 rays per step (chunks are independent; rays/s is chunk-size invariant)"
    OUT.emit(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "rays/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(args, {"rays_per_step": sample}),
        "cpu_baseline": {"value": v, "unit": "rays/s", "cores": cores, "kind": "port", "sample": desc},
        "e2e": {"value": v, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


def train_leg(args, dev, rank, world, ro, rd, K):
    """Secondary number: training iterations/s of the reference loop's step (run_plnerf.py:1283-1303) on the
    blender_linear.txt shape (N_rand=1024 rays per rank, N_samples=128, N_importance=64): stash-mode forward,
    loss, backward (compositing bwd + gradient chain + weight-gradient GEMMs), ONE flat NCCL all-reduce of both
    networks' gradients, two fused Adam steps.  Weak scaling: global batch = 1024 x ranks."""
    import torch
    import torch.distributed as dist
    from plnerf_b200 import run_plnerf as RP, synth
    from plnerf_b200 import dist as PD
    from plnerf_b200.run_nerf_helpers import NeRF
    N_rand, Ns, Ni, iters, warm = 1024, 128, 64, 30, 5

    def mk(seed):
        net = NeRF(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=[4], use_viewdirs=True)
        net.load_state_dict({k: torch.from_numpy(v.copy())
                             for k, v in synth.nerf_params(seed, density_boost=False, **NET_KW).items()})
        return net.to(dev)
    net_c, net_f = mk(11), mk(12)
    bucket = PD.FlatGradBucket([net_c, net_f])
    opt_f = torch.optim.Adam(net_f.parameters(), lr=5e-4, fused=True)
    opt_c = torch.optim.Adam(net_c.parameters(), lr=5e-4, fused=True)
    ro_t, rd_t = torch.from_numpy(ro).to(dev), torch.from_numpy(rd).to(dev)
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234 + rank)
    target_img = torch.rand(H * W, 3, device=dev, generator=gen)

    def step():
        idx = torch.randint(0, H * W, (N_rand,), device=dev, generator=gen)
        rays = torch.stack([ro_t[idx], rd_t[idx]])
        rgb, disp, acc, extras = RP.render(H, W, K, chunk=CHUNK, rays=rays, ndc=False, near=2., far=6.,
                                           use_viewdirs=True, network_query_fn=None, network_fn=net_c,
                                           network_fine=net_f, N_samples=Ns, N_importance=Ni, perturb=1.0,
                                           white_bkgd=True, mode="linear", color_mode="midpoint", retraw=True)
        tgt = target_img[idx]
        loss = torch.mean((rgb - tgt) ** 2) + torch.mean((extras["rgb0"] - tgt) ** 2)
        bucket.zero_()
        loss.backward()
        bucket.allreduce_mean()
        opt_f.step(); opt_c.step()
        return loss

    for _ in range(warm):
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t[0]) / iters
    rows = N_rand * (2 * Ns + Ni)
    return {"iters_per_s": 1e3 / ms, "ms_per_iter": ms, "rays_per_iter_global": N_rand * world,
            "algorithmic_tflops_per_gpu": rows * 3489024 / ms / 1e9, "final_loss": float(loss.item()),
            "config": f"N_rand={N_rand}/rank, N_samples={Ns}, N_importance={Ni}, linear/midpoint, viewdirs, white_bkgd, "
                      f"bf16 tensor-core operands, 2x fused Adam, 1 flat all-reduce ({bucket.flat.numel() * 4} B) per step"}


def train_step_leg(dev, K):
    """Extra (1 GPU only): the same training shape through plnerf_b200.train.TrainStep -- pixel draw, ray generation of
    the chosen pixels, loss gradient, one flat gradient buffer and ONE fused Adam all on the device,
    against train_leg's reference-loop-shaped step above."""
    import torch
    from plnerf_b200 import synth, train as T
    from plnerf_b200.run_nerf_helpers import NeRF
    N_rand, Ns, Ni, iters, warm = 1024, 128, 64, 30, 5

    def mk(seed):
        net = NeRF(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=[4], use_viewdirs=True)
        net.load_state_dict({k: torch.from_numpy(v.copy())
                             for k, v in synth.nerf_params(seed, density_boost=False, **NET_KW).items()})
        return net.to(dev)
    net_c, net_f = mk(11), mk(12)
    kw = dict(network_query_fn=None, network_fn=net_c, network_fine=net_f, N_samples=Ns, N_importance=Ni, perturb=1.0,
              white_bkgd=True, raw_noise_std=0., mode="linear", color_mode="midpoint", use_viewdirs=True, ndc=False,
              near=2., far=6.)
    step = T.TrainStep(H, W, K, kw, N_rand=N_rand, chunk=CHUNK, lrate=5e-4, coarse_lrate=5e-4, lrate_decay=500, seed=1234)
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234)
    target = torch.rand(H, W, 3, device=dev, generator=gen)
    pose = torch.from_numpy(synth.pose_spherical(-180.0, -30.0, 4.0)[:3, :4].astype(np.float32).copy()).to(dev)
    for i in range(warm):
        step(target, pose, i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(warm, warm + iters):
        out = step(target, pose, i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    return {"iters_per_s": 1e3 / ms, "ms_per_iter": ms, "final_loss": float(out["loss"].item()),
            "config": "plnerf_b200.train.TrainStep, same shape as `train`: device-side pixel draws (64 iterations per batched draw) + "
                      "pack_pixel_rays, direct loss gradient, forward/backward kernels called without an autograd graph into flat gradient + parameter buffers, 1 fused Adam launch"}


def run_ours(args):
    import torch
    import torch.distributed as dist
    import plnerf_b200
    from plnerf_b200 import ops, run_plnerf as RP, synth
    from plnerf_b200.run_nerf_helpers import NeRF

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (plnerf_b200 has no CPU path; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    ops.set_precision(args.precision)

    pc, pf = synth.nerf_params(1, **NET_KW), synth.nerf_params(2, **NET_KW)

    def mk(p):
        net = NeRF(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=[4], use_viewdirs=True)
        net.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in p.items()})
        return net.to(dev)
    net_c, net_f = mk(pc), mk(pf)
    ro, rd, K = build_rays(rank)
    n = ro.shape[0]
    host_rays = torch.stack([torch.from_numpy(ro), torch.from_numpy(rd)]).pin_memory()       # [2, n, 3]
    host_out = {k: torch.empty(s, dtype=torch.float32).pin_memory() for k, s in
                (("rgb", (n, 3)), ("disp", (n,)), ("acc", (n,)))}
    kwargs = dict(network_fn=net_c, network_query_fn=None, network_fine=net_f, N_samples=N_SAMPLES,
                  N_importance=N_IMPORTANCE, perturb=1.0, white_bkgd=True, raw_noise_std=0.0, mode="linear",
                  color_mode="midpoint", lindisp=False, seed=1234)
    # device-resident packed rays for the `value` leg: exactly what render() packs (run_plnerf.py:140-164)
    with torch.no_grad():
        d_o, d_d = host_rays[0].to(dev), host_rays[1].to(dev)
        vd = d_d / torch.norm(d_d, dim=-1, keepdim=True)
        near = 2.0 * torch.ones_like(d_d[..., :1]); far = 6.0 * torch.ones_like(d_d[..., :1])
        dev_rays = torch.cat([d_o, d_d, near, far, vd], -1).contiguous()

    def step_resident():
        with torch.no_grad():
            return RP.batchify_rays(dev_rays, CHUNK, **kwargs)

    def step_e2e():
        with torch.no_grad():
            r = host_rays.to(dev, non_blocking=True)
            rgb, disp, acc, _ = RP.render(H, W, K, chunk=CHUNK, rays=r, ndc=False, near=2., far=6., use_viewdirs=True,
                                          **kwargs)
            host_out["rgb"].copy_(rgb, non_blocking=True)
            host_out["disp"].copy_(disp, non_blocking=True)
            host_out["acc"].copy_(acc, non_blocking=True)
            torch.cuda.current_stream().synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, profile=False):
        barrier()
        if profile:
            ops.profile_enable(True)
        l0 = ops.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) * 1e3
        ms = max(e0.elapsed_time(e1), 0.0)
        launches = ops.launch_count() - l0
        prof = ops.profile_read() if profile else None
        if profile:
            ops.profile_enable(False)
        t = torch.tensor([ms, wall], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        barrier()
        return float(t[0]), float(t[1]), launches, prof

    for _ in range(args.warmup):
        step_resident()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms, wall, launches, prof = timed(step_resident, args.steps, profile=True)
    clocks = sampler.stop() if rank == 0 else None
    for _ in range(max(1, args.warmup // 2)):
        step_e2e()
    ms_e, wall_e, _, _ = timed(step_e2e, args.steps)
    # e2e timing: device events only see the stream; the step ends with a stream sync, so wall == device span
    e2e_ms = max(ms_e, wall_e)

    total_rays = world * n * args.steps
    value = total_rays / (ms / 1e3)
    e2e_value = total_rays / (e2e_ms / 1e3)
    sust, burst, peak_src = peaks()
    mlp_ms, mlp_n, mlp_rows = prof
    ach = mlp_rows * FLOP_PER_EVAL / (mlp_ms / 1e3) / 1e12 if mlp_ms > 0 else 0.0
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "mlp_fwd_traffic.json")
    if os.path.isfile(tpath):
        with open(tpath) as f:
            traffic = json.load(f).get(args.precision)
    out = {
        "metric": METRIC, "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16" if args.precision == "bf16" else "bf16x3 (hi/lo split, ~fp32 products)",
        "data": "synthetic", "config": config_dict(args, {"precision": args.precision}),
        "e2e": {"value": e2e_value, "unit": "rays/s", "h2d_bytes_per_step": int(host_rays.numel() * 4),
                "d2h_bytes_per_step": int(sum(v.numel() for v in host_out.values()) * 4),
                "ms_per_step": e2e_ms / args.steps},
        "gpu_launches": int(launches),
        "roofline": {"bound": "tensor", "achieved": ach, "peak": sust, "unit": "TFLOP/s",
                     "frac": ach / sust if sust else None, "traffic": traffic,
                     "kernel": "k_mlp_fwd", "launches": int(mlp_n), "avg_launch_ms": mlp_ms / max(1, mlp_n),
                     "share_of_step": mlp_ms / ms if ms > 0 else None, "peak_source": peak_src,
                     "frac_of_burst_peak": ach / burst if burst else None,
                     "algorithmic_flop_per_launch": mlp_rows * FLOP_PER_EVAL / max(1, mlp_n)},
        "clocks": clocks,
    }
    if not args.no_train:
        out["train"] = train_leg(args, dev, rank, world, ro, rd, K)
        if world == 1:
            try:
                out["train"]["device_side_step"] = train_step_leg(dev, K)
            except Exception as e:  # an extra: the lines above must survive its failure
                out["train"]["device_side_step"] = {"error": repr(e)}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count()
        torch.set_num_threads(cores)
        sample = args.cpu_rays
        idx = np.random.RandomState(0).choice(H * W, sample, replace=False)
        oracle_render(ro[idx[:256]], rd[idx[:256]], K, pc, pf)   # warm-up
        t0 = time.perf_counter()
        reps = 0
        while reps < 3 and (time.perf_counter() - t0) < 20.0:
            oracle_render(ro[idx], rd[idx], K, pc, pf)
            reps += 1
        dt = time.perf_counter() - t0
        # PSNR of the CUDA path against the CPU port on the SAME rays and the SAME random draws (the "PSNR vs ref" half of
        # BASELINE.json's metric): identical images would be +inf dB; bf16 operands land around 60-70 dB.
        try:
            rs = np.random.RandomState(0)
            t_rand = rs.rand(sample, N_SAMPLES).astype(np.float32)
            u = rs.rand(sample, N_IMPORTANCE).astype(np.float32)
            ref_img = oracle_render(ro[idx], rd[idx], K, pc, pf)          # same seed -> same t_rand / u as above
            with torch.no_grad():
                r = torch.stack([torch.from_numpy(ro[idx]), torch.from_numpy(rd[idx])]).to(dev)
                kw2 = dict(kwargs); kw2.pop("seed", None)
                rgb_g, _, _, ex_g = RP.render(H, W, K, chunk=CHUNK, rays=r, ndc=False, near=2., far=6., use_viewdirs=True,
                                              t_rand=torch.from_numpy(t_rand).to(dev), u=torch.from_numpy(u).to(dev), **kw2)
            mse = float(np.mean((rgb_g.cpu().numpy().astype(np.float64) - ref_img["rgb_map"]) ** 2))
            dmax = float(np.abs(ex_g["depth_map"].cpu().numpy() - ref_img["depth_map"]).max())
            out["psnr_vs_reference"] = {"psnr_db": (-10.0 * np.log10(mse)) if mse > 0 else float("inf"), "rgb_mse": mse,
                                        "max_abs_depth_err": dmax, "rays": int(sample), "precision": args.precision,
                                        "reference": "oracle/plnerf_oracle.py (fp32 CPU port pinned to the reference), same rays and draws"}
        except Exception as e:  # the timing lines above must survive a failure of this extra
            out["psnr_vs_reference"] = {"error": repr(e)}
        out["cpu_baseline"] = {"value": sample * reps / dt, "unit": "rays/s", "cores": cores, "kind": "port",
                               "sample": f"{reps} x {sample} rays of the same 640000-ray image, numpy/torch-CPU oracle "
                                         f"(oracle/plnerf_oracle.py), {cores} host threads"}
    if rank == 0:
        OUT.emit(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


class StdoutToStderr:
    """Everything libraries print on fd 1 (e.g. NCCL's version banner) goes to stderr; the ONE JSON line is
    written to the real stdout at the end."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def emit(self, line):
        sys.stdout.flush()
        os.write(self.saved, (line + "\n").encode())

    def __exit__(self, *a):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)


OUT = None


def main():
    global OUT
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=os.environ.get("PLNERF_PRECISION", "bf16"), choices=["bf16", "bf16x3"])
    ap.add_argument("--cpu-rays", type=int, default=4096, help="rays per CPU-baseline repetition")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-train", action="store_true", help="skip the secondary training-step timing")
    args = ap.parse_args()
    with StdoutToStderr() as OUT:
        if args.impl == "reference":
            run_reference(args)
        else:
            run_ours(args)


if __name__ == "__main__":
    main()
