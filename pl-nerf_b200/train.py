"""One optimisation step of the reference's training loop with the per-iteration host work moved to the
device (reference: run_plnerf.py:1253-1315, the ``no_batching`` branch).

What the reference does every iteration around ``render()`` and what replaces it here:

* ``get_rays`` for the whole image + a ``[H*W, 2]`` coordinate grid + ``np.random.choice(H*W, N_rand,
  replace=False)`` on the host + three fancy-index gathers (:1259-1280)  ->  ``PixelSampler`` (device-side draws with
  the same "N_rand distinct pixels, uniform, random order" law, 64 iterations per batched draw) and
  ``ops.pack_pixel_rays`` (ONE kernel that generates and packs the rays of the chosen pixels only);
* ``img2mse`` twice + ``loss.backward()`` (:1289-1299)  ->  the loss gradient ``2 (rgb - target) / (3 B)`` is formed
  directly and the forward (stash mode) and backward kernels are called back to back, the weight-gradient kernels
  accumulating straight into the flat gradient buffer -- no autograd graph, no per-parameter AccumulateGrad add
  (48 small launches per step), no zero placeholders for the unused map gradients; the loss value stays on the
  device (no ``.item()`` in the step);
* ``optimizer.zero_grad()`` x2, ``optimizer.step()`` x2 (:1286-1303)  ->  both networks' gradients alias ONE flat
  buffer (``dist.FlatGradBucket``: one fill, one NCCL all-reduce when several ranks train), the parameters alias a
  second flat buffer, and ONE fused Adam launch updates the two flat segments (fine, coarse) instead of 48 tensors
  (measured 33 us instead of 111 us);
* the learning-rate decay loop (:1307-1315) is kept as is, including the reference's quirk that the coarse group
  is assigned the *fine* rate;
* ``retraw=True`` (:1284) is not requested: the only reader (``trans``, :1291) is unused.

Data parallel: every rank draws the SAME global pixel batch (same generator seed), renders the
contiguous shard ``dist.shard_bounds(N_rand)`` of it with ``ray_id_offset`` = the shard's first global ray, scales
its loss gradient by the GLOBAL batch size and sums gradients over ranks -- the W-rank job is the 1-rank job.

CUDA only; gradients need ``precision='bf16'`` (see autograd.py).
"""
import torch

from . import autograd as AG
from . import dist as pdist
from . import ops
from . import run_plnerf as RP


def crop_window(H, W, precrop_frac):
    """(row0, col0, rows, cols) of the centre crop used for the first ``precrop_iters`` iterations
    (run_plnerf.py:1262-1270): rows H//2-dH .. H//2+dH-1, cols W//2-dW .. W//2+dW-1."""
    dH = int(H // 2 * precrop_frac)
    dW = int(W // 2 * precrop_frac)
    return H // 2 - dH, W // 2 - dW, 2 * dH, 2 * dW


def _window(H, W, precrop_frac):
    return (0, 0, H, W) if precrop_frac is None else crop_window(H, W, precrop_frac)


def _to_flat_ids(k, H, W, window):
    r0, c0, rows, cols = window
    if window == (0, 0, H, W):
        return k
    return (r0 + torch.div(k, cols, rounding_mode="floor")) * W + (c0 + k % cols)


def distinct_draws(n_items, n_draws, n_rows, device, generator=None):
    """[n_rows, n_draws] int64: every row is n_draws DISTINCT items of range(n_items), uniformly, in random order.

    Large windows (n_items >= 64 n_draws): each row is the first n_draws distinct values of 2 n_draws i.i.d. uniform
    draws -- the first-occurrence subsequence of an i.i.d. stream is exactly a uniform sample without replacement in
    random order; more than n_draws repeats among 2 n_draws draws from >= 64 n_draws items has probability < 8^-n_draws
    (the missing slots would then repeat item 0).  All rows are produced by one batched stable sort + scan with fixed
    shapes (no host round trip).  Small windows: a batched permutation (argsort of uniform keys)."""
    if n_draws > n_items:
        raise ValueError(f"cannot take {n_draws} distinct items from {n_items}")   # np.random.choice raises too
    if n_items < 64 * n_draws:
        keys = torch.rand((n_rows, n_items), device=device, generator=generator)
        return torch.argsort(keys, dim=1)[:, :n_draws].contiguous()
    m = 2 * n_draws
    c = torch.randint(0, n_items, (n_rows, m), device=device, generator=generator)
    vals, order = torch.sort(c, dim=1, stable=True)
    dup_sorted = torch.zeros((n_rows, m), dtype=torch.bool, device=device)
    dup_sorted[:, 1:] = vals[:, 1:] == vals[:, :-1]               # stable sort: the later draw of an equal pair is the repeat
    keep = torch.ones((n_rows, m), dtype=torch.bool, device=device)
    keep.scatter_(1, order, ~dup_sorted)                           # back to draw order
    pos = torch.cumsum(keep, 1) - 1
    slot = torch.where(keep & (pos < n_draws), pos, torch.full_like(pos, n_draws))   # column n_draws = discard
    out = torch.zeros((n_rows, n_draws + 1), dtype=torch.int64, device=device)
    out.scatter_(1, slot, c)
    return out[:, :n_draws].contiguous()


def sample_pixels(H, W, N_rand, device, generator=None, precrop_frac=None):
    """N_rand distinct pixels of an H x W image (or of its centre crop), uniformly, in random order -- the law of
    ``np.random.choice(coords.shape[0], size=[N_rand], replace=False)`` (run_plnerf.py:1276) -- as int64 flat ids
    row*W + col on ``device``.  No host round trip."""
    window = _window(H, W, precrop_frac)
    k = distinct_draws(window[2] * window[3], N_rand, 1, device, generator)[0]
    return _to_flat_ids(k, H, W, window)


class PixelSampler:
    """The pixel batches of consecutive iterations, drawn ``block`` iterations at a time: one batched draw costs about
    as much as a single one (a dozen small launches, ~130 us on a B200), so the per-iteration cost drops to a slice.
    A block never mixes windows: it is redrawn when the crop window changes (end of the precrop phase)."""

    def __init__(self, H, W, N_rand, device, generator=None, block=64):
        self.H, self.W, self.N_rand, self.device, self.generator, self.block = H, W, N_rand, device, generator, block
        self._window, self._buf, self._next = None, None, 0

    def next(self, precrop_frac=None):
        window = _window(self.H, self.W, precrop_frac)
        if self._buf is None or window != self._window or self._next >= self._buf.shape[0]:
            k = distinct_draws(window[2] * window[3], self.N_rand, self.block, self.device, self.generator)
            self._buf, self._window, self._next = _to_flat_ids(k, self.H, self.W, window), window, 0
        row = self._buf[self._next]
        self._next += 1
        return row


def decayed_lrate(lrate, lrate_decay, global_step, decay_rate=0.1):
    """run_plnerf.py:1307-1309: lrate * 0.1 ** (global_step / (lrate_decay * 1000))."""
    return lrate * (decay_rate ** (global_step / (lrate_decay * 1000)))


def alias_parameters_flat(params):
    """Move the storage of ``params`` into ONE flat fp32 buffer: every ``p.data`` becomes a view of it (values kept,
    ``state_dict`` / ``load_state_dict`` keep working through the views).  An in-place update of the flat buffer
    does not move the views' version counters -- callers must ``ops.invalidate_packed(net)`` afterwards."""
    dev = params[0].device
    flat = torch.empty(sum(p.numel() for p in params), dtype=torch.float32, device=dev)
    off = 0
    for p in params:
        n = p.numel()
        view = flat[off:off + n].view_as(p)
        view.copy_(p.data)
        p.data = view
        off += n
    return flat


class FlatAdam:
    """The reference's two ``torch.optim.Adam(params, lr, betas=(0.9, 0.999))`` (run_plnerf.py:431-447) as ONE optimiser over
    contiguous segments of a flat parameter / gradient buffer: one ``plnerf_adam_step`` launch per parameter group (one
    group when all segments share the learning rate) instead of a multi-tensor launch over 48 tensors.

    ``segments``: [(name, n_elements, lr)] in buffer order.  ``param_groups`` is a list of {"lr", "segments"} dicts that
    the learning-rate loop writes like the reference does (:1310-1315).  State (``exp_avg``, ``exp_avg_sq`` flat like the
    parameters, one ``step`` count) converts to and from the per-tensor ``optimizer.state_dict()`` of the reference's
    checkpoints with ``export_reference_state`` / ``import_reference_state``."""

    def __init__(self, flat_params, flat_grads, segments, betas=(0.9, 0.999), eps=1e-8):
        self.flat_params, self.flat_grads = flat_params, flat_grads
        self.betas, self.eps = betas, eps
        self.exp_avg = torch.zeros_like(flat_params)
        self.exp_avg_sq = torch.zeros_like(flat_params)
        self.step_count = 0
        self.segments, off = {}, 0
        for name, n, _ in segments:
            self.segments[name] = (off, off + n)
            off += n
        if off != flat_params.numel():
            raise ValueError("segments do not cover the flat buffer")
        lrs = [lr for _, _, lr in segments]
        if all(lr == lrs[0] for lr in lrs):
            self.param_groups = [{"lr": lrs[0], "segments": [name for name, _, _ in segments]}]
        else:
            self.param_groups = [{"lr": lr, "segments": [name]} for name, _, lr in segments]

    def _range(self, group):
        lo = min(self.segments[k][0] for k in group["segments"])
        hi = max(self.segments[k][1] for k in group["segments"])
        return lo, hi

    def step(self):
        self.step_count += 1
        for g in self.param_groups:
            lo, hi = self._range(g)
            ops.adam_step(self.flat_params[lo:hi], self.flat_grads[lo:hi], self.exp_avg[lo:hi], self.exp_avg_sq[lo:hi],
                          g["lr"], self.step_count, self.betas, self.eps)

    def zero_grad(self):
        self.flat_grads.zero_()

    def state_dict(self):
        return {"step": self.step_count, "exp_avg": self.exp_avg, "exp_avg_sq": self.exp_avg_sq,
                "param_groups": [dict(g) for g in self.param_groups]}

    def load_state_dict(self, sd):
        self.step_count = int(sd["step"])
        self.exp_avg.copy_(sd["exp_avg"])
        self.exp_avg_sq.copy_(sd["exp_avg_sq"])
        for g, src in zip(self.param_groups, sd["param_groups"]):
            g["lr"] = src["lr"]

    def export_reference_state(self, name, params):
        """``torch.optim.Adam(params).state_dict()`` of segment ``name`` (e.g. "fine": the ``optimizer_state_dict`` entry of
        the reference's checkpoints, run_plnerf.py:1326-1331), ``params`` = that network's parameters in order."""
        lo, hi = self.segments[name]
        lr = next(g["lr"] for g in self.param_groups if name in g["segments"])
        state, off = {}, lo
        for i, p in enumerate(params):
            n = p.numel()
            if self.step_count > 0:
                state[i] = {"step": torch.tensor(float(self.step_count)), "exp_avg": self.exp_avg[off:off + n].view_as(p).clone(),
                            "exp_avg_sq": self.exp_avg_sq[off:off + n].view_as(p).clone()}
            off += n
        if off != hi:
            raise ValueError(f"parameters do not match segment {name!r}")
        group = {"lr": lr, "betas": self.betas, "eps": self.eps, "weight_decay": 0, "amsgrad": False, "maximize": False,
                 "foreach": None, "capturable": False, "differentiable": False, "fused": None,
                 "params": list(range(len(params)))}
        return {"state": state, "param_groups": [group]}

    def import_reference_state(self, name, params, sd):
        """Load a reference ``optimizer.state_dict()`` (per-tensor moments) into segment ``name``.  The flat optimiser keeps
        ONE step count: it is taken from the loaded state."""
        lo, hi = self.segments[name]
        off = lo
        for i, p in enumerate(params):
            n = p.numel()
            st = sd["state"].get(i)
            if st is not None:
                self.exp_avg[off:off + n].copy_(st["exp_avg"].reshape(-1))
                self.exp_avg_sq[off:off + n].copy_(st["exp_avg_sq"].reshape(-1))
                self.step_count = int(st["step"])
            off += n
        if off != hi:
            raise ValueError(f"parameters do not match segment {name!r}")
        for g in self.param_groups:
            if name in g["segments"]:
                g["lr"] = sd["param_groups"][0]["lr"]


class TrainStep:
    """Callable optimisation step.  ``render_kwargs`` is the reference's ``render_kwargs_train`` dict
    (run_plnerf.py:475-487: network_fn, network_fine, N_samples, N_importance, perturb, white_bkgd, raw_noise_std,
    mode, color_mode, [lindisp], plus use_viewdirs / ndc / near / far which ``render`` consumes itself).

    Construct it after the networks are on their device: the constructor re-homes their parameters into one flat
    buffer (``alias_parameters_flat``); a later ``net.to(...)`` would break that aliasing."""

    def __init__(self, H, W, K, render_kwargs, N_rand=1024, chunk=1024 * 32, lrate=5e-4, coarse_lrate=5e-4,
                 lrate_decay=250, precrop_iters=0, precrop_frac=.5, constant_init=0, seed=0):
        kw = dict(render_kwargs)
        self.H, self.W, self.K = int(H), int(W), K
        self.N_rand, self.chunk = int(N_rand), int(chunk)
        self.use_viewdirs = bool(kw.pop("use_viewdirs", False))
        self.ndc = bool(kw.pop("ndc", True))
        self.near, self.far = float(kw.pop("near", 0.)), float(kw.pop("far", 1.))
        for k in ("retraw", "constant_init", "verbose"):   # set per step below / unused
            kw.pop(k, None)
        kw.setdefault("network_query_fn", None)            # positional in render_rays; the query is fused
        self.net_c, self.net_f = kw["network_fn"], kw.get("network_fine")
        self.render_kwargs = kw
        self.lrate, self.coarse_lrate, self.lrate_decay = lrate, coarse_lrate, lrate_decay
        self.precrop_iters, self.precrop_frac, self.constant_init = precrop_iters, precrop_frac, constant_init
        self.device = next(self.net_c.parameters()).device
        self._check_device()
        self.nets = [n for n in (self.net_f, self.net_c) if n is not None]
        self.bucket = pdist.FlatGradBucket(self.nets, extra=2)     # + the two squared-error accumulators of the loss
        self.flat_params = alias_parameters_flat(self.bucket.params)
        # one optimiser segment per network (fine first, like the bucket); equal rates -> one launch over both
        segs = [(name, sum(p.numel() for p in net.parameters() if p.requires_grad), lr)
                for name, net, lr in zip(["fine", "coarse"] if self.net_f is not None else ["coarse"], self.nets,
                                         [lrate, coarse_lrate] if self.net_f is not None else [coarse_lrate])]
        self.optimizer = FlatAdam(self.flat_params, self.bucket.flat, segs, betas=(0.9, 0.999))
        # {state_dict name: view of the flat gradient buffer} per network: what the weight-gradient kernels add into
        self._grads = {net: {k: p.grad for k, p in net.named_parameters()} for net in self.nets}
        self._prepared = {}            # the same, as the ABI's gradient struct (built on first use: needs the CUDA library)
        # explicit depth / sampler draws (t_rand [B, N_samples], u [B, N_importance] of the GLOBAL batch) are sliced per
        # shard and chunk; explicit noise or the pytest hook go through render_rays' autograd.Function instead
        self._direct = not any(kw.get(k) is not None for k in ("noise0", "noise1")) and not kw.get("pytest")
        if self._direct:
            if kw.get("precision") not in (None, "bf16") or (kw.get("precision") is None and ops.get_precision() != "bf16"):
                raise NotImplementedError("plnerf_b200: gradients are implemented for precision='bf16' only")
            if not all(p.requires_grad for n in self.nets for p in n.parameters()):
                raise NotImplementedError("plnerf_b200: TrainStep trains every parameter of both networks")
        # the same stream of pixel draws on every rank (the shard is taken after the draw)
        self.generator = torch.Generator(device=self.device)
        self.generator.manual_seed(int(seed))
        self.sampler = PixelSampler(self.H, self.W, self.N_rand, self.device, self.generator)

    def _check_device(self):
        if self.device.type != "cuda":
            raise RuntimeError("plnerf_b200: the NeRF modules must live on a CUDA device (no CPU fallback)")

    def pixels(self, i):
        """The global pixel batch of iteration i (precrop window for the first precrop_iters iterations)."""
        return self.sampler.next(self.precrop_frac if i < self.precrop_iters else None)

    def _grad_arg(self, net):
        """The network's gradient views as the backward entry wants them (None for a missing fine network)."""
        if net is None:
            return None
        if not self.flat_params.is_cuda:       # (host-logic tests drive the step with stand-ins that take the dict)
            return self._grads[net]
        if net not in self._prepared:
            self._prepared[net] = ops.PreparedGrads(net, self._grads[net])
        return self._prepared[net]

    def _forward_backward_direct(self, rays, target_s, scale, ray0, constant_init, pix=None):
        """Forward (stash mode), the two MSE terms and the backward of every chunk as ONE C call
        (``autograd.train_rays_mse``), the weight-gradient kernels accumulating straight into the flat gradient buffer: no
        autograd graph, no per-parameter AccumulateGrad add (48 launches per step through ``loss.backward()``), no
        zero-filled placeholders for the unused map gradients; the coarse pass's backward runs beside the fine pass.
        ``target_s`` [n, 3] holds the rays' targets, or -- with ``pix`` (their pixel ids) -- the whole image [H*W, 3]
        (the gather :1280 then happens inside the loss kernel).  The sums of squared errors of rgb_map / rgb0 are added
        to ``self.bucket.extra``; returns whether a coarse term exists."""
        kw = self.render_kwargs
        Ns, Ni = int(kw["N_samples"]), int(kw.get("N_importance", 0))
        std = float(kw.get("raw_noise_std", 0.))
        sqerr = self.bucket.extra
        with torch.no_grad():
            for c0 in range(0, rays.shape[0], self.chunk):
                r = rays[c0:c0 + self.chunk]
                n = r.shape[0]
                cfg = dict(net_c=self.net_c, net_f=self.net_f if Ni > 0 else None, N_samples=Ns, N_importance=Ni,
                           mode="constant" if constant_init else kw["mode"], color_mode=kw["color_mode"],
                           perturb=kw.get("perturb", 0.) > 0., white_bkgd=bool(kw.get("white_bkgd", False)),
                           lindisp=bool(kw.get("lindisp", False)), zero_tol=kw.get("zero_tol", 1e-4),
                           epsilon=kw.get("epsilon", 1e-3), farcolorfix=bool(kw.get("farcolorfix", False)),
                           t_rand=None if kw.get("t_rand") is None else kw["t_rand"][ray0 + c0:ray0 + c0 + n],
                           u=None if kw.get("u") is None else kw["u"][ray0 + c0:ray0 + c0 + n], noise0=torch.randn((n, Ns), device=r.device) * std if std > 0. else None,
                           noise1=torch.randn((n, Ns + Ni), device=r.device) * std if (std > 0. and Ni > 0) else None,
                           seed=kw["seed"] if kw.get("seed") is not None else RP._next_seed(), ray_id_offset=ray0 + c0)
                if Ni > 0 and not cfg["perturb"] and cfg["u"] is None:      # det=True: the reference's linspace u (see render_rays)
                    cfg["u"] = torch.linspace(0., 1., steps=Ni, device=r.device).expand(n, Ni).contiguous()
                t, px = (target_s, pix[c0:c0 + n]) if pix is not None else (target_s[c0:c0 + n], None)
                AG.train_rays_mse(cfg, r, t, px, scale, sqerr, self._grad_arg(self.net_c),
                                  self._grad_arg(self.net_f) if Ni > 0 else None)
        return Ni > 0

    def _forward_backward_autograd(self, rays, target_s, scale, ray0, constant_init):
        """The same step through render_rays' autograd.Function (used when explicit draws / the pytest hook are asked
        for in the render kwargs)."""
        with torch.enable_grad():
            ret = RP.batchify_rays(rays, self.chunk, ray_id_offset=ray0, retraw=False, constant_init=constant_init,
                                   **self.render_kwargs)
        rgb, rgb0 = ret["rgb_map"], ret.get("rgb0")
        g, g0 = ops.mse_loss_grad(rgb.detach(), None if rgb0 is None else rgb0.detach(), target_s, scale, self.bucket.extra)
        torch.autograd.backward([rgb] if rgb0 is None else [rgb, rgb0], [g] if rgb0 is None else [g, g0])
        return rgb0 is not None

    def __call__(self, target, pose, i, global_step=None, pix=None):
        """target [H, W, 3] (or [H*W, 3]) device image, pose = c2w [3|4, 4], i = iteration number (drives precrop /
        constant_init exactly like the reference's loop variable).  ``global_step`` defaults to i - 1: the reference's
        loop variable starts at start + 1 while its global_step starts at start and is incremented at the END of the
        iteration (run_plnerf.py:1153, 1235, 1400), so the decayed rate of iteration i uses i - 1.  Returns {"loss", "img_loss", "img_loss0", "pix"}
        as device tensors (the loss values are detached scalars; nothing is synchronised).  With several ranks each
        loss value is this rank's share of the global mean (their sum over ranks is the reference's loss)."""
        global_step = max(i - 1, 0) if global_step is None else global_step
        if pix is None:
            pix = self.pixels(i)
        B = pix.shape[0]
        lo, hi = pdist.shard_bounds(B)
        local = pix[lo:hi]
        rays = ops.pack_pixel_rays(self.H, self.W, self.K, pose, local, ndc=self.ndc, near=self.near, far=self.far,
                                   use_viewdirs=self.use_viewdirs)
        out = self._optimise(rays, target.reshape(-1, 3), B, lo, i, global_step, pix=local)
        out["pix"] = pix
        return out

    def step_rays(self, batch_rays, target_s, i, global_step=None):
        """The ``use_batching`` branch of the reference loop (run_plnerf.py:1238-1250, the LLFF configs): the caller
        slices ``batch_rays`` [2, B, 3] (origins, directions) and ``target_s`` [B, 3] out of its pre-shuffled ray bank;
        everything after that is the same step.  Several ranks: every rank passes the same GLOBAL batch and renders its
        shard of it."""
        global_step = max(i - 1, 0) if global_step is None else global_step
        B = batch_rays.shape[1]
        lo, hi = pdist.shard_bounds(B)
        rays, _ = ops.pack_rays(self.H, self.W, self.K, rays=(batch_rays[0][lo:hi], batch_rays[1][lo:hi]), ndc=self.ndc,
                                near=self.near, far=self.far, use_viewdirs=self.use_viewdirs)
        return self._optimise(rays, target_s[lo:hi], B, lo, i, global_step)

    def _optimise(self, rays, target_s, B, lo, i, global_step, pix=None):
        """rays: this rank's packed shard [n, 8|11] of a global batch of B rays starting at global ray ``lo``; target_s:
        their targets [n, 3], or the whole image [H*W, 3] with ``pix`` = their pixel ids."""
        self.bucket.zero_()                                      # gradients and the two squared-error sums: one fill
        scale = 2.0 / (3.0 * B)                                  # d mean((x - t)^2) / dx over the GLOBAL batch
        if self._direct:
            coarse = self._forward_backward_direct(rays, target_s, scale, lo, i < self.constant_init, pix=pix)
        else:
            if pix is not None:
                target_s = target_s[pix]
            coarse = self._forward_backward_autograd(rays, target_s, scale, lo, i < self.constant_init)
        losses = self.bucket.extra * (1.0 / (3.0 * B))           # img2mse of the fine / coarse map (this rank's share)
        img_loss, img_loss0 = losses[0], (losses[1] if coarse else None)
        self.bucket.allreduce_sum()
        self.optimizer.step()
        for net in self.nets:          # the update went through the flat alias: the views' version counters did not move
            ops.invalidate_packed(net)
        if self._direct and self.flat_params.is_cuda:      # ... and both packed copies of both networks are rebuilt in one launch
            ops.repack_train(self.nets)
        new_lrate = decayed_lrate(self.lrate, self.lrate_decay, global_step)
        for g in self.optimizer.param_groups:                    # both groups get the fine rate (Appendix B.1)
            g["lr"] = new_lrate
        loss = img_loss if img_loss0 is None else img_loss + img_loss0
        return {"loss": loss, "img_loss": img_loss, "img_loss0": img_loss0}
