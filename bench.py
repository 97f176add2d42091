#!/usr/bin/env python
"""Benchmark of the PL-NeRF ray-rendering hot path on B200 (contract in the task brief / DESIGN.md).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload lego|llff]

Workload `lego` (BASELINE.json configs[1], the default): 800x800 Blender-lego-shaped synthetic rays (640 000 rays per
image), N_samples=64 + N_importance=128, PL ("linear") quadrature, midpoint colour, viewdirs, white background, seeded
random-init (density-boosted) coarse+fine 8x256 NeRF, reference chunking (32 768 rays per render_rays call).  One "step"
= one full image through render_rays.  Weak scaling: every rank renders its own image (different pose), no data-path
collective.  Workload `llff` (configs[3]): 378x504 fern-shaped NDC rays, N_samples=64 + N_importance=64.

Printed JSON (rank 0), ONE line:
  value         rays/s with the packed rays already resident in HBM, bf16 operands / fp32 accumulate (north-star mode)
  e2e           the same image through the public `render()` call from pinned HOST rays, H2D copy of the rays and D2H
                copy of rgb/disp/acc inside the timed region
  roofline      algorithmic MLP FLOPs / CUDA-event time of the fused-MLP launches inside the timed region, against the
                measured sustained bf16 tensor peak (MEASURED_PEAKS.json)
  parity_mode   the SAME measurements (value, e2e, roofline, psnr_vs_reference) in the mode that meets the 1e-4 parity
                gate against the fp32 reference (bf16x3: hi/lo split operands, 3 MMAs per product)
  psnr_vs_reference   both modes' images against the UNMODIFIED reference's on the same rays and the same draws
  cpu_baseline  the unmodified reference's own render() (oracle/_ref, staged copies of the reference modules) on a
                bounded ray sample on this box's host cores (kind "reference"); the numpy/torch oracle port only if the
                staged reference is absent (kind "port")
`--impl reference` times that same unmodified reference on CPU with all host threads: one 32 768-ray render_rays chunk
of the workload per step (shrunk, and said so, if a first probe shows the run would not end within a few minutes).
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

CHUNK = 1024 * 32
FLOP_PER_EVAL = 1186816            # 593 408 MAC, viewdirs network
NET_KW = dict(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=(4,), use_viewdirs=True)
WORKLOADS = {
    "lego": dict(H=800, W=800, Ns=64, Ni=128, ndc=False, near=2., far=6., white_bkgd=True,
                 metric="rays/sec (64 coarse + 128 fine samples)",
                 desc="lego-shaped 800x800 synthetic rays (640000 rays/step), N_samples=64, N_importance=128, mode=linear, "
                      "color_mode=midpoint, use_viewdirs, white_bkgd, perturb=1, chunk=32768, random-init density-boosted "
                      "coarse+fine NeRF 8x256"),
    "llff": dict(H=378, W=504, Ns=64, Ni=64, ndc=True, near=0., far=1., white_bkgd=False,
                 metric="rays/sec (64 coarse + 64 fine samples, LLFF NDC)",
                 desc="fern-shaped 378x504 forward-facing synthetic rays through ndc_rays (190512 rays/step), N_samples=64, "
                      "N_importance=64, mode=linear, color_mode=midpoint, use_viewdirs, perturb=1, chunk=32768, random-init "
                      "density-boosted coarse+fine NeRF 8x256"),
}


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


def build_rays(wl, rank):
    """(rays_o, rays_d, K) of the workload's full image for this rank (numpy float32 [H*W,3])."""
    from plnerf_b200 import synth
    if wl is WORKLOADS["lego"]:
        theta = float(np.linspace(-180, 180, 41)[:-1][rank % 40])
        ro, rd, K, _ = synth.lego_rays(None, H=wl["H"], W=wl["W"], theta=theta)
    else:
        n = wl["H"] * wl["W"]
        ro, rd, K, _ = synth.llff_rays(n, H=wl["H"], W=wl["W"], seed=rank)
    return ro, rd, K


def config_dict(args, wl):
    n = wl["H"] * wl["W"]
    S_last = wl["Ns"] + wl["Ni"]
    return {"workload": wl["desc"], "rays_per_step": n, "chunk": CHUNK,
            "parallelism": f"ray-sharded x{args.gpus} (one image per rank)",
            "l2_policy": f"per-step working set (depths / weights / samples of {n} rays x {S_last} samples, "
                         f"{n * S_last * 4 * 6 / 1e9:.2f} GB) >> 126 MB L2; no explicit flush"}


# ---------------------------------------------------------------------------------------------------------------------
# the unmodified reference on CPU (oracle/_ref staged copies; /root/reference in the build container)
# ---------------------------------------------------------------------------------------------------------------------
class CpuReference:
    """render() of the UNMODIFIED reference (kind "reference"), or of the oracle port when it is not staged (kind "port")."""

    def __init__(self, wl, pc, pf):
        import torch
        import refimport
        self.wl, self.torch = wl, torch
        self.cores = os.cpu_count()
        torch.set_num_threads(self.cores)
        self.kind = "reference" if refimport.available() else "port"
        self.pc, self.pf = pc, pf
        if self.kind == "reference":
            H, R = refimport.load()
            R.device = torch.device("cpu")      # the module picks cuda when one is visible; this arm is the reference's CPU path
            self.R = R

            def mk(prm):
                net = H.NeRF(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=[4], use_viewdirs=True)
                net.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in prm.items()})
                return net
            self.net_c, self.net_f = mk(pc), mk(pf)
            embed_fn, _ = H.get_embedder(10, 0)
            embeddirs_fn, _ = H.get_embedder(4, 0)
            self.q = lambda p, v, fn: R.run_network(p, v, fn, embed_fn=embed_fn, embeddirs_fn=embeddirs_fn, netchunk=1024 * 64)

    def describe(self):
        return ("unmodified reference run_plnerf.render() on CPU (oracle/_ref), torch " + self.torch.__version__
                if self.kind == "reference" else "numpy/torch-CPU oracle port (oracle/plnerf_oracle.py)")

    def render(self, ro, rd, K):
        """One render() call with the reference's pytest draws (np.random.seed(0) heads): -> dict of numpy outputs."""
        wl, torch = self.wl, self.torch
        if self.kind == "reference":
            rays = torch.stack([torch.from_numpy(ro), torch.from_numpy(rd)])
            with torch.no_grad():
                rgb, disp, acc, ex = self.R.render(wl["H"], wl["W"], K, chunk=CHUNK, rays=rays, ndc=wl["ndc"], near=wl["near"],
                                                   far=wl["far"], use_viewdirs=True, network_query_fn=self.q,
                                                   network_fn=self.net_c, network_fine=self.net_f, N_samples=wl["Ns"],
                                                   N_importance=wl["Ni"], perturb=1.0, raw_noise_std=0.,
                                                   white_bkgd=wl["white_bkgd"], mode="linear", color_mode="midpoint",
                                                   lindisp=False, pytest=True)
            return {"rgb_map": rgb.numpy(), "depth_map": ex["depth_map"].numpy()}
        import plnerf_oracle as O
        n = ro.shape[0]

        def head(shape):
            np.random.seed(0)
            return np.random.rand(*shape).astype(np.float32)
        return O.render(wl["H"], wl["W"], K, ro, rd, chunk=CHUNK, ndc=wl["ndc"], near=wl["near"], far=wl["far"],
                        use_viewdirs=True, t_rand=head((n, wl["Ns"])), u=head((n, wl["Ni"])), params_coarse=self.pc,
                        params_fine=self.pf, N_samples=wl["Ns"], mode="linear", color_mode="midpoint",
                        N_importance=wl["Ni"], white_bkgd=wl["white_bkgd"],
                        net_kw=dict(D=8, skips=(4,), input_ch=63, input_ch_views=27, use_viewdirs=True))


def run_reference(args, wl):
    """CPU arm: the unmodified reference, one render_rays chunk of the workload per step (rank 0 only)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from plnerf_b200 import synth
    pc, pf = synth.nerf_params(1, **NET_KW), synth.nerf_params(2, **NET_KW)
    ref = CpuReference(wl, pc, pf)
    ro, rd, K = build_rays(wl, 0)
    n_img = ro.shape[0]
    sample = min(args.cpu_rays if args.cpu_rays > 0 else CHUNK, n_img)
    order = np.random.RandomState(0).permutation(n_img)
    # probe: would (warmup + steps) chunks of `sample` rays end within a few minutes?  shrink the chunk if not
    t0 = time.perf_counter()
    ref.render(ro[order[:1024]], rd[order[:1024]], K)
    probe_rate = 1024 / (time.perf_counter() - t0)
    budget_s = 420.0
    fit = int(budget_s * probe_rate / max(1, args.steps + args.warmup))
    shrunk = fit < sample
    if shrunk:
        sample = max(1024, fit // 1024 * 1024)
    idx = order[:sample]
    ro, rd = ro[idx], rd[idx]
    for _ in range(args.warmup):
        ref.render(ro, rd, K)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ref.render(ro, rd, K)
    dt = time.perf_counter() - t0
    v = sample * args.steps / dt
    desc = (f"{sample} rays per step = one render() call of the {n_img}-ray image's rays"
            + (" (a full 32768-ray render_rays chunk)" if sample == CHUNK else "")
            + (f"; shrunk from {CHUNK} so that {args.steps}+{args.warmup} steps fit {budget_s:.0f} s" if shrunk else "")
            + f"; chunks are independent, rays/s is chunk-size invariant; {ref.describe()}, {ref.cores} host threads")
    OUT.emit(json.dumps({
        "impl": "reference", "metric": wl["metric"], "value": v, "unit": "rays/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(args, wl), "rays_timed_per_step": int(sample),
        "cpu_baseline": {"value": v, "unit": "rays/s", "cores": ref.cores, "kind": ref.kind, "sample": desc},
        "e2e": {"value": v, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


# ---------------------------------------------------------------------------------------------------------------------
# training legs (secondary numbers)
# ---------------------------------------------------------------------------------------------------------------------
def train_leg(args, dev, rank, world, ro, rd, K, H, W):
    """Training iterations/s of the reference loop's step (run_plnerf.py:1283-1303) on the blender_linear.txt shape
    (N_rand=1024 rays per rank, N_samples=128, N_importance=64): stash-mode forward, loss, backward (compositing bwd +
    gradient chain + weight-gradient GEMMs), ONE flat NCCL all-reduce of both networks' gradients, two fused Adam steps.
    Weak scaling: global batch = 1024 x ranks."""
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


def train_step_leg(dev, K, H, W, rank, world, weak=False):
    """The reference's JOB (global N_rand = 1024 rays per iteration, blender_linear.txt) through plnerf_b200.train.TrainStep:
    device-side pixel draw, ray generation of the chosen pixels, loss gradient, one flat gradient buffer, ONE fused Adam.
    With W ranks every rank draws the same global batch and renders its contiguous 1024/W-ray shard (strong scaling);
    the flat gradient is summed by one all-reduce."""
    import torch
    import torch.distributed as dist
    from plnerf_b200 import synth, train as T
    from plnerf_b200.run_nerf_helpers import NeRF
    # weak: 1024 rays per rank.  Two timed windows back to back: the first 30 iterations after the warm-up (what rounds 1-2
    # reported; the SM clock is still near its maximum) and the following 500 (~0.9 s: the step under the power cap, like the
    # rendering legs and like a real training run) -- `iters_per_s` is the SUSTAINED one.
    N_rand, Ns, Ni, burst, iters, warm = 1024 * (world if weak else 1), 128, 64, 30, 500, 10

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
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    for i in range(warm, warm + burst):
        out = step(target, pose, i)
    e1.record()
    for i in range(warm + burst, warm + burst + iters):
        out = step(target, pose, i)
    e2.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1), e1.elapsed_time(e2)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_burst, ms = float(t[0]) / burst, float(t[1]) / iters
    # the step's one collective in isolation: the flat 4.77 MB gradient all-reduce (device time, max over ranks)
    ar_ms = None
    if world > 1:
        for _ in range(5):
            step.bucket.allreduce_sum()
        dist.barrier()
        torch.cuda.synchronize()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(20):
            step.bucket.allreduce_sum()
        a1.record()
        torch.cuda.synchronize()
        ta = torch.tensor([a0.elapsed_time(a1) / 20], device=dev, dtype=torch.float64)
        dist.all_reduce(ta, op=dist.ReduceOp.MAX)
        ar_ms = float(ta[0])
    # fraction of the tensor roofline of SURVEY.md 8d: 3 489 024 FLOP per network evaluation of a training step
    # (forward + input-gradient chain + weight gradients), (2 Ns + Ni) evaluations per ray, against the sustained bf16 peak
    peak_sus, _, _ = peaks()
    tflops = (N_rand / world) * (2 * Ns + Ni) * 3489024 / ms / 1e9
    return {"iters_per_s": 1e3 / ms, "ms_per_iter": ms, "iters": iters,
            "first_30_iters": {"iters_per_s": 1e3 / ms_burst, "ms_per_iter": ms_burst},
            "final_loss": float(out["loss"].item()),
            "scaling": "weak" if weak else "strong", "algorithmic_tflops_per_gpu": tflops,
            "frac_of_tensor_roofline": tflops / peak_sus,
            "rays_per_iter_global": N_rand, "allreduce_ms": ar_ms, "allreduce_bytes": int(step.bucket.flat.numel() * 4),
            "config": f"plnerf_b200.train.TrainStep, global N_rand={N_rand} sharded over {world} rank(s), N_samples={Ns}, "
                      f"N_importance={Ni}: device-side pixel draws (64 iterations per batched draw) + pack_pixel_rays, direct "
                      "loss gradient, forward/backward kernels called without an autograd graph into flat gradient + "
                      "parameter buffers, 1 flat all-reduce (sum), 1 flat Adam launch (plnerf_adam_step)"}


# ---------------------------------------------------------------------------------------------------------------------
def run_ours(args, wl):
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
    H, W, Ns, Ni = wl["H"], wl["W"], wl["Ns"], wl["Ni"]

    pc, pf = synth.nerf_params(1, **NET_KW), synth.nerf_params(2, **NET_KW)

    def mk(p):
        net = NeRF(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=[4], use_viewdirs=True)
        net.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in p.items()})
        return net.to(dev)
    net_c, net_f = mk(pc), mk(pf)
    ro, rd, K = build_rays(wl, rank)
    n = ro.shape[0]
    host_rays = torch.stack([torch.from_numpy(ro), torch.from_numpy(rd)]).pin_memory()       # [2, n, 3]
    host_out = {k: torch.empty(s, dtype=torch.float32).pin_memory() for k, s in
                (("rgb", (n, 3)), ("disp", (n,)), ("acc", (n,)))}
    render_kw = dict(ndc=wl["ndc"], near=wl["near"], far=wl["far"], use_viewdirs=True)
    kwargs = dict(network_fn=net_c, network_query_fn=None, network_fine=net_f, N_samples=Ns, N_importance=Ni, perturb=1.0,
                  white_bkgd=wl["white_bkgd"], raw_noise_std=0.0, mode="linear", color_mode="midpoint", lindisp=False,
                  seed=1234)
    # device-resident packed rays for the `value` leg: exactly what render() packs (run_plnerf.py:140-164)
    with torch.no_grad():
        dev_rays, _ = ops.pack_rays(H, W, K, rays=host_rays.to(dev), **render_kw)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def measure(precision, steps, warmup, sample_clocks):
        """value / e2e / roofline of one precision mode."""
        kw = dict(kwargs, precision=precision)

        def step_resident():
            with torch.no_grad():
                return RP.batchify_rays(dev_rays, CHUNK, **kw)

        def step_e2e():
            with torch.no_grad():
                r = host_rays.to(dev, non_blocking=True)
                rgb, disp, acc, _ = RP.render(H, W, K, chunk=CHUNK, rays=r, **render_kw, **kw)
                host_out["rgb"].copy_(rgb, non_blocking=True)
                host_out["disp"].copy_(disp, non_blocking=True)
                host_out["acc"].copy_(acc, non_blocking=True)
                torch.cuda.current_stream().synchronize()

        def timed(fn, profile=False):
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

        for _ in range(warmup):
            step_resident()
        sampler = ClockSampler(local) if (sample_clocks and rank == 0) else None
        if sampler:
            sampler.start()
        ms, wall, launches, prof = timed(step_resident, profile=True)
        clocks = sampler.stop() if sampler else None
        for _ in range(max(1, warmup // 2)):
            step_e2e()
        ms_e, wall_e, _, _ = timed(step_e2e)
        e2e_ms = max(ms_e, wall_e)    # the step ends with a stream sync, so wall == device span; take the larger
        total_rays = world * n * steps
        sust, burst, peak_src = peaks()
        mlp_ms, mlp_n, mlp_rows = prof
        ach = mlp_rows * FLOP_PER_EVAL / (mlp_ms / 1e3) / 1e12 if mlp_ms > 0 else 0.0
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "mlp_fwd_traffic.json")
        if os.path.isfile(tpath):
            with open(tpath) as f:
                traffic = json.load(f).get(precision)
        mmas = 3 if precision == "bf16x3" else 1
        return {
            "value": total_rays / (ms / 1e3), "ms_per_step": ms / steps, "steps": steps, "gpu_launches": int(launches),
            "e2e": {"value": total_rays / (e2e_ms / 1e3), "unit": "rays/s", "h2d_bytes_per_step": int(host_rays.numel() * 4),
                    "d2h_bytes_per_step": int(sum(v.numel() for v in host_out.values()) * 4), "ms_per_step": e2e_ms / steps},
            "roofline": {"bound": "tensor", "achieved": ach, "peak": sust, "unit": "TFLOP/s",
                         "frac": ach / sust if sust else None, "traffic": traffic,
                         "kernel": "k_mlp3" if precision == "bf16" else "k_mlp_fwd<bf16x3>", "launches": int(mlp_n),
                         "avg_launch_ms": mlp_ms / max(1, mlp_n), "share_of_step": mlp_ms / ms if ms > 0 else None,
                         "peak_source": peak_src, "frac_of_burst_peak": ach / burst if burst else None,
                         "algorithmic_flop_per_launch": mlp_rows * FLOP_PER_EVAL / max(1, mlp_n),
                         "issued_mma_tflops": ach * mmas},
            "clocks": clocks}

    main = measure("bf16", args.steps, args.warmup, True)
    out = {
        "metric": wl["metric"], "value": main["value"], "unit": "rays/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": main["ms_per_step"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": config_dict(args, wl),
        "e2e": main["e2e"], "gpu_launches": main["gpu_launches"], "roofline": main["roofline"], "clocks": main["clocks"],
    }
    if not args.no_parity:
        psteps = max(1, min(args.steps, 5))
        pm = measure("bf16x3", psteps, 3, False)
        out["parity_mode"] = {"precision": "bf16x3", "dtype": "bf16x3 (hi/lo split operands, 3 tcgen05 MMAs per product, ~2^-16 products)",
                              "why": "the mode that meets the 1e-4 parity gate against the fp32 reference (tests/test_sized_golden.py); "
                                     "the headline bf16 mode is gated at its measured bounds",
                              "value": pm["value"], "unit": "rays/s", "steps": psteps, "warmup": 3, "ms_per_step": pm["ms_per_step"],
                              "e2e": pm["e2e"], "roofline": pm["roofline"], "gpu_launches": pm["gpu_launches"]}
    if not args.no_train and wl is WORKLOADS["lego"]:
        out["train"] = train_leg(args, dev, rank, world, ro, rd, K, H, W)
        try:
            out["train"]["device_side_step"] = train_step_leg(dev, K, H, W, rank, world)
            if world > 1:      # the same step with 1024 rays per rank (weak scaling; equals the line above at one rank)
                out["train"]["device_side_step_weak"] = train_step_leg(dev, K, H, W, rank, world, weak=True)
        except Exception as e:  # an extra: the lines above must survive its failure
            out["train"]["device_side_step"] = {"error": repr(e)}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        ref = CpuReference(wl, pc, pf)
        sample = min(args.cpu_rays if args.cpu_rays > 0 else 4096, n)
        idx = np.random.RandomState(0).choice(n, sample, replace=False)
        ref.render(ro[idx[:256]], rd[idx[:256]], K)   # warm-up
        t0 = time.perf_counter()
        reps = 0
        while reps < 3 and (time.perf_counter() - t0) < 20.0:
            ref_img = ref.render(ro[idx], rd[idx], K)
            reps += 1
        dt = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": sample * reps / dt, "unit": "rays/s", "cores": ref.cores, "kind": ref.kind,
                               "sample": f"{reps} x {sample} rays of the same {n}-ray image, {ref.describe()}, {ref.cores} host threads"}
        # PSNR of the CUDA path against the CPU reference on the SAME rays and the SAME (pytest) draws: the "PSNR vs ref"
        # half of BASELINE.json's metric, for both modes
        psnr = {}
        for prec in ("bf16",) + (() if args.no_parity else ("bf16x3",)):
            try:
                with torch.no_grad():
                    r = torch.stack([torch.from_numpy(ro[idx]), torch.from_numpy(rd[idx])]).to(dev)
                    kw2 = dict(kwargs, precision=prec, pytest=True)
                    kw2.pop("seed", None)
                    rgb_g, _, _, ex_g = RP.render(H, W, K, chunk=CHUNK, rays=r, **render_kw, **kw2)
                mse = float(np.mean((rgb_g.cpu().numpy().astype(np.float64) - ref_img["rgb_map"]) ** 2))
                psnr[prec] = {"psnr_db": (-10.0 * np.log10(mse)) if mse > 0 else float("inf"), "rgb_mse": mse,
                              "max_abs_rgb_err": float(np.abs(rgb_g.cpu().numpy() - ref_img["rgb_map"]).max()),
                              "max_abs_depth_err_over_far": float(np.abs(ex_g["depth_map"].cpu().numpy() - ref_img["depth_map"]).max() / wl["far"]),
                              "rays": int(sample)}
            except Exception as e:  # the timing lines above must survive a failure of this extra
                psnr[prec] = {"error": repr(e)}
        out["psnr_vs_reference"] = dict(psnr.get("bf16", {}), precision="bf16",
                                        reference=f"{ref.describe()}, same rays and draws (pytest=True)")
        if "parity_mode" in out and "bf16x3" in psnr:
            out["parity_mode"]["psnr_vs_reference"] = psnr["bf16x3"]
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
    ap.add_argument("--workload", default="lego", choices=list(WORKLOADS))
    ap.add_argument("--cpu-rays", type=int, default=0, help="rays per CPU repetition (0 = 32768 for --impl reference, 4096 for cpu_baseline)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the bf16x3 parity-mode block")
    ap.add_argument("--no-train", action="store_true", help="skip the secondary training-step timing")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    with StdoutToStderr() as OUT:
        if args.impl == "reference":
            run_reference(args, wl)
        else:
            run_ours(args, wl)


if __name__ == "__main__":
    main()
