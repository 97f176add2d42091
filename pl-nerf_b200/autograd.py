"""Autograd bridge for training: loss.backward() through render_rays / NeRF.forward.

Forward = the same sm_100a kernels as inference with the fused MLP in "stash" mode (bf16 activations +
ReLU masks of every layer written as weight-gradient operand tiles); backward = compositing backward
(k_composite_bwd), the input-gradient chain (the fused MLP kernel run on W^T), the weight-gradient
GEMMs (k_wgrad, MN-major tcgen05 operands, TMEM-resident accumulators) and the small heads.
What is differentiated follows the reference: gradients flow from rgb_map / depth_map /
acc_map / disp_map (fine and coarse) to both networks' parameters; the importance samples are
detached (run_plnerf.py:728) and ray inputs carry no gradient.  bf16 tensor-core operands with fp32
accumulation; networks with and without view directions (output_linear head, run_nerf_helpers.py:100-103).

``forward_stashed`` / ``backward_stashed`` are the two plain functions behind the autograd.Function; the training
step (train.TrainStep) issues forward, loss and backward of a ray batch as one C call (``train_rays_mse``) without
building an autograd graph.
"""
import torch

from . import ops


def _names(net):
    return [k for k, _ in net.named_parameters()]


def forward_stashed(cfg, rays):
    """The training forward of render_rays: ONE C call (plnerf_render_rays_fwd_train: the inference kernels with the
    fused MLP in stash mode).  Returns (outs, saved, stashes): ``outs`` = (rgb, disp, acc, depth, raw) for
    N_importance == 0, else (rgb, disp, acc, depth, raw1, rgb0, disp0, acc0, depth0, z_std); ``saved`` feeds
    ``backward_stashed`` (it owns the workspace holding depths, raw and the two stashes); ``stashes`` is unused."""
    Ni = cfg["N_importance"]
    ret, ctx = ops.render_rays_fwd_train(rays, cfg["net_c"], cfg["net_f"], cfg["N_samples"], Ni, cfg["mode"], cfg["color_mode"],
                                         perturb=cfg["perturb"], white_bkgd=cfg["white_bkgd"], lindisp=cfg["lindisp"],
                                         raw_noise_std=0.0, zero_tol=cfg["zero_tol"], epsilon=cfg["epsilon"],
                                         farcolorfix=cfg["farcolorfix"], t_rand=cfg["t_rand"], u=cfg["u"], noise0=cfg["noise0"],
                                         noise1=cfg["noise1"], seed=cfg["seed"], ray_id_offset=cfg["ray_id_offset"],
                                         retraw=cfg.get("retraw", False))
    raw = ret.get("raw")
    if raw is None:
        raw = torch.empty(0, device=rays.device)
    outs = (ret["rgb_map"], ret["disp_map"], ret["acc_map"], ret["depth_map"], raw)
    if Ni > 0:
        outs += (ret["rgb0"], ret["disp0"], ret["acc0"], ret["depth0"], ret["z_std"])
    return outs, ctx, None


def backward_stashed(cfg, saved, stashes, g_fine, g_coarse, grads_c, grads_f):
    """Backward of ``forward_stashed``: ONE C call (plnerf_render_rays_bwd).  ``g_fine`` / ``g_coarse`` = (g_rgb, g_disp,
    g_acc, g_depth) of the fine / coarse maps (entries may be None; for N_importance == 0 only ``g_coarse`` is used).
    Parameter gradients are ACCUMULATED into ``grads_c`` / ``grads_f`` ({state_dict name: fp32 tensor}; ``grads_f`` is
    ignored when there is no separate fine network)."""
    if cfg["N_importance"] == 0:
        ops.render_rays_bwd(saved, g_coarse, None, grads_c, None)
    else:
        ops.render_rays_bwd(saved, g_fine, g_coarse, grads_c, grads_f if cfg["net_f"] is not None else grads_c)


def train_rays_mse(cfg, rays, target, pix, scale, sqerr, grads_c, grads_f):
    """``forward_stashed`` -> ``ops.mse_loss_grad`` -> ``backward_stashed`` of one ray batch as ONE C call
    (plnerf_train_rays_mse; the coarse pass's loss + backward run beside the fine pass): what train.TrainStep issues per
    chunk.  ``target`` [n, 3] (or the whole image [H*W, 3] with ``pix`` = the rays' pixel ids); the two sums of squared
    errors are added to ``sqerr`` [2]; parameter gradients are ACCUMULATED into grads_c / grads_f."""
    ops.train_rays_mse(rays, cfg["net_c"], cfg["net_f"], cfg["N_samples"], cfg["N_importance"], cfg["mode"], cfg["color_mode"],
                       target, scale, sqerr, grads_c, grads_f if cfg["net_f"] is not None else None, pix=pix,
                       perturb=cfg["perturb"], white_bkgd=cfg["white_bkgd"], lindisp=cfg["lindisp"], raw_noise_std=0.0,
                       zero_tol=cfg["zero_tol"], epsilon=cfg["epsilon"], farcolorfix=cfg["farcolorfix"], t_rand=cfg["t_rand"],
                       u=cfg["u"], noise0=cfg["noise0"], noise1=cfg["noise1"], seed=cfg["seed"],
                       ray_id_offset=cfg["ray_id_offset"])


class _RenderRaysFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, cfg, rays, *params):
        outs, saved, stashes = forward_stashed(cfg, rays)
        ctx.cfg = cfg
        ctx.saved = saved
        if cfg["N_importance"] == 0:
            ctx.mark_non_differentiable(outs[4])
        else:
            ctx.mark_non_differentiable(outs[4], outs[9])
        return outs

    @staticmethod
    def backward(ctx, *g):
        cfg = ctx.cfg
        net_c, net_f = cfg["net_c"], cfg["net_f"]
        grads_c = ops.zero_grads_like(net_c)      # one flat zeroed buffer per network (one fill kernel, not 24)
        grads_f = ops.zero_grads_like(net_f) if (net_f is not None and cfg["N_importance"] > 0) else None
        if cfg["N_importance"] == 0:
            backward_stashed(cfg, ctx.saved, None, None, (g[0], g[1], g[2], g[3]), grads_c, None)
        else:
            backward_stashed(cfg, ctx.saved, None, (g[0], g[1], g[2], g[3]), (g[5], g[6], g[7], g[8]),
                             grads_c, grads_f)
        ctx.saved = None
        out = [None, None] + [grads_c[k] for k in _names(net_c)]
        if net_f is not None:
            out += [grads_f[k] for k in _names(net_f)] if grads_f is not None else [None] * len(_names(net_f))
        return tuple(out)


def render_rays_autograd(ray_batch, network_fn, network_fine, N_samples, N_importance, mode, color_mode, perturb_on,
                         white_bkgd, lindisp, raw_noise_std, zero_tol, epsilon, farcolorfix, t_rand, u, noise0, noise1,
                         seed, ray_id_offset, retraw, precision):
    if precision not in (None, "bf16") or (precision is None and ops.get_precision() != "bf16"):
        raise NotImplementedError("plnerf_b200: gradients are implemented for precision='bf16' only")
    rays = ray_batch.detach().float().contiguous()
    n = rays.shape[0]
    dev = rays.device
    if raw_noise_std > 0.:   # the backward must see the same draws: make them explicit
        if noise0 is None:
            noise0 = torch.randn((n, N_samples), device=dev) * raw_noise_std
        if noise1 is None and N_importance > 0:
            noise1 = torch.randn((n, N_samples + N_importance), device=dev) * raw_noise_std
    cfg = dict(net_c=network_fn, net_f=network_fine if N_importance > 0 else None, N_samples=N_samples,
               N_importance=N_importance, mode=mode, color_mode=color_mode, perturb=perturb_on, white_bkgd=white_bkgd,
               lindisp=lindisp, zero_tol=zero_tol, epsilon=epsilon, farcolorfix=farcolorfix, t_rand=t_rand, u=u,
               noise0=noise0, noise1=noise1, seed=seed, ray_id_offset=ray_id_offset, retraw=retraw)
    params = list(network_fn.parameters())
    if cfg["net_f"] is not None:
        params += list(cfg["net_f"].parameters())
    outs = _RenderRaysFn.apply(cfg, rays, *params)
    if N_importance == 0:
        rgb, disp, acc, depth, raw = outs
        ret = {"rgb_map": rgb, "disp_map": disp, "acc_map": acc, "depth_map": depth}
        if retraw:
            ret["raw"] = raw
        return ret
    rgb, disp, acc, depth, raw, rgb0, disp0, acc0, depth0, z_std = outs
    ret = {"rgb_map": rgb, "disp_map": disp, "acc_map": acc, "depth_map": depth}
    if retraw:
        ret["raw"] = raw
    ret.update(rgb0=rgb0, disp0=disp0, depth0=depth0, acc0=acc0, z_std=z_std)
    return ret


def mlp_forward_autograd(net, x):
    raise NotImplementedError(
        "plnerf_b200: NeRF.forward on pre-embedded rows has no backward kernel (gradients flow through render_rays / "
        "render, whose fused query recomputes the encoding); call it under torch.no_grad().")
