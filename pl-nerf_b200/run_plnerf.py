"""Drop-in for the hot-path functions of the reference's ``run_plnerf`` module.

``render`` / ``batchify_rays`` / ``render_rays`` / ``raw2outputs`` / ``run_network`` /
``compute_weights`` / ``compute_weights_piecewise_linear`` keep the reference's signatures, argument
meaning and return conventions (run_plnerf.py:68-175, 504-758); the work is done by the sm_100a
kernels behind the C ABI.  ``install(module)`` rebinds these names inside an imported reference
``run_plnerf`` module so that the unmodified training / evaluation script drives this path
(``batchify_rays`` looks ``render_rays`` up as a module global, run_plnerf.py:100).
"""
import numpy as np
import torch

from . import nerf_extract_mesh
from . import ops
from . import run_nerf_helpers as helpers
from .run_nerf_helpers import get_rays, ndc_rays

DEBUG = False
_call_counter = 0


def _next_seed():
    """Philox key for one render_rays call: torch's global seed mixed with a call counter (no
    device sync; reproducible after torch.manual_seed)."""
    global _call_counter
    _call_counter += 1
    return (torch.initial_seed() * 0x9E3779B97F4A7C15 + _call_counter * 0xD1B54A32D192ED03) & ((1 << 63) - 1)


def batchify(fn, chunk):
    """run_plnerf.py:68-75."""
    if chunk is None:
        return fn

    def ret(inputs):
        return torch.cat([fn(inputs[i:i + chunk]) for i in range(0, inputs.shape[0], chunk)], 0)
    return ret


def run_network(inputs, viewdirs, fn, embed_fn, embeddirs_fn, netchunk=1024 * 64):
    """run_plnerf.py:78-92: PE + chunked MLP on explicit points [N,S,3] (op-level entry; the fused
    path inside render_rays never materialises the points)."""
    inputs_flat = torch.reshape(inputs, [-1, inputs.shape[-1]])
    embedded = embed_fn(inputs_flat)
    if viewdirs is not None:
        input_dirs = viewdirs[:, None].expand(inputs.shape)
        input_dirs_flat = torch.reshape(input_dirs, [-1, input_dirs.shape[-1]])
        embedded = torch.cat([embedded, embeddirs_fn(input_dirs_flat)], -1)
    outputs_flat = batchify(fn, netchunk)(embedded)
    return torch.reshape(outputs_flat, list(inputs.shape[:-1]) + [outputs_flat.shape[-1]])


def _pack_bounds(z_vals, rays_d, near=None, far=None):
    n = z_vals.shape[0]
    rays = torch.zeros((n, 8), device=z_vals.device, dtype=torch.float32)
    rays[:, 3:6] = rays_d
    if near is not None:
        rays[:, 6:7] = near.reshape(n, 1)
        rays[:, 7:8] = far.reshape(n, 1)
    return rays


def compute_weights(raw, z_vals, rays_d, noise=0.):
    """run_plnerf.py:504-513."""
    nz = None if (isinstance(noise, float) and noise == 0.) else noise
    return ops.raw2outputs(raw, z_vals, _pack_bounds(z_vals, rays_d), "constant", "midpoint", noise=nz)[3]


def compute_weights_piecewise_linear(raw, z_vals, near, far, rays_d, noise=0., return_tau=False):
    """run_plnerf.py:516-550."""
    nz = None if (isinstance(noise, float) and noise == 0.) else noise
    r = ops.raw2outputs(raw, z_vals, _pack_bounds(z_vals, rays_d, near, far), "linear", "midpoint", noise=nz)
    return (r[3], r[5], r[6]) if return_tau else r[3]


def _noise_like(shape, raw_noise_std, pytest, device):
    """The additive density noise of raw2outputs (run_plnerf.py:567-576)."""
    if raw_noise_std <= 0.:
        return None
    if pytest:
        np.random.seed(0)
        return torch.Tensor(np.random.rand(*shape) * raw_noise_std).to(device)
    return torch.randn(shape, device=device) * raw_noise_std


def raw2outputs(raw, z_vals, near, far, rays_d, mode, color_mode, raw_noise_std=0, pytest=False, white_bkgd=False,
                farcolorfix=False):
    """run_plnerf.py:553-624.  Returns (rgb_map, disp_map, acc_map, weights, depth_map, tau, T)."""
    noise = _noise_like(list(raw[..., 3].shape), raw_noise_std, pytest, raw.device)
    rays = _pack_bounds(z_vals, rays_d, near, far)
    return ops.raw2outputs(raw, z_vals, rays, mode, color_mode, noise=noise, white_bkgd=white_bkgd,
                           farcolorfix=farcolorfix)


def render_rays(ray_batch, network_fn, network_query_fn, N_samples, mode, color_mode, retraw=False, lindisp=False,
                perturb=0., N_importance=0, network_fine=None, white_bkgd=False, raw_noise_std=0., verbose=False,
                pytest=False, quad_solution_v2=False, zero_tol=1e-4, epsilon=1e-3, farcolorfix=False,
                constant_init=False, t_rand=None, u=None, noise0=None, noise1=None, seed=None, ray_id_offset=0,
                precision=None):
    """Volumetric rendering of a ray batch: run_plnerf.py:627-758, same arguments and returned dict.

    ``network_query_fn`` is accepted for signature compatibility; the positional-encoding widths are
    read from ``network_fn.input_ch / input_ch_views`` and the query is fused into the MLP kernel.
    Extra keyword-only knobs (not in the reference): explicit random draws ``t_rand/u/noise0/noise1``,
    Philox ``seed`` / ``ray_id_offset`` for draws made on the device, and ``precision``.
    """
    N_rays = ray_batch.shape[0]
    dev = ray_batch.device
    if constant_init:
        mode = "constant"
    perturb_on = perturb > 0.
    if pytest:  # the reference's determinism hook: every draw is the head of np.random.seed(0)
        def head(shape):
            np.random.seed(0)
            return np.random.rand(*shape)
        if perturb_on and t_rand is None:
            t_rand = torch.Tensor(head((N_rays, N_samples))).to(dev)
        if N_importance > 0 and perturb_on and u is None:
            u = torch.Tensor(head((N_rays, N_importance))).to(dev)
        if raw_noise_std > 0.:
            if noise0 is None:
                noise0 = torch.Tensor(head((N_rays, N_samples)) * raw_noise_std).to(dev)
            if noise1 is None and N_importance > 0:
                noise1 = torch.Tensor(head((N_rays, N_samples + N_importance)) * raw_noise_std).to(dev)
    if N_importance > 0 and not perturb_on and u is None:
        # det=True (run_nerf_helpers.py:377-379).  NOTE: in linear mode the reference raises an
        # IndexError here (u == 1.0 indexes past tau_diff); the kernel clamps that index instead.
        u = torch.linspace(0., 1., steps=N_importance, device=dev).expand(N_rays, N_importance).contiguous()
    if seed is None:
        seed = _next_seed()
    if torch.is_grad_enabled() and any(p.requires_grad for net in (network_fn, network_fine) if net is not None
                                       for p in net.parameters()):
        from .autograd import render_rays_autograd
        return render_rays_autograd(ray_batch, network_fn, network_fine, N_samples, N_importance, mode, color_mode,
                                    perturb_on, white_bkgd, lindisp, raw_noise_std, zero_tol, epsilon, farcolorfix,
                                    t_rand, u, noise0, noise1, seed, ray_id_offset, retraw, precision)
    ret = ops.render_rays_fwd(ray_batch, network_fn, network_fine, N_samples, N_importance, mode, color_mode,
                              perturb=perturb_on, white_bkgd=white_bkgd, lindisp=lindisp, raw_noise_std=raw_noise_std,
                              zero_tol=zero_tol, epsilon=epsilon, farcolorfix=farcolorfix, t_rand=t_rand, u=u,
                              noise0=noise0, noise1=noise1, seed=seed, ray_id_offset=ray_id_offset, retraw=retraw,
                              precision=precision)
    if DEBUG:
        for k in ret:
            if torch.isnan(ret[k]).any() or torch.isinf(ret[k]).any():
                print(f"! [Numerical Error] {k} contains nan or inf.")
    return ret


def batchify_rays(rays_flat, chunk=1024 * 32, **kwargs):
    """run_plnerf.py:95-107."""
    all_ret = {}
    for i in range(0, rays_flat.shape[0], chunk):
        kw = dict(kwargs)
        kw["ray_id_offset"] = kwargs.get("ray_id_offset", 0) + i
        for k in ("t_rand", "u", "noise0", "noise1"):
            if kw.get(k) is not None:
                kw[k] = kw[k][i:i + chunk]
        ret = render_rays(rays_flat[i:i + chunk], **kw)
        for k in ret:
            all_ret.setdefault(k, []).append(ret[k])
    return {k: (v[0] if len(v) == 1 else torch.cat(v, 0)) for k, v in all_ret.items()}


def render(H, W, K, chunk=1024 * 32, rays=None, c2w=None, ndc=True, near=0., far=1., use_viewdirs=False,
           c2w_staticcam=None, **kwargs):
    """run_plnerf.py:110-175: same arguments, returns [rgb_map, disp_map, acc_map, extras]."""
    # ray generation, viewdirs (before NDC), ndc_rays and the [o, d, near, far, viewdir] packing: one kernel
    # (plnerf_pack_rays) instead of ~15 full-image torch ops (run_plnerf.py:138-164)
    rays, sh = ops.pack_rays(H, W, K, c2w=c2w, rays=rays, ndc=ndc, near=near, far=far, use_viewdirs=use_viewdirs,
                             c2w_staticcam=c2w_staticcam)
    all_ret = batchify_rays(rays, chunk, **kwargs)
    for k in all_ret:
        k_sh = list(sh[:-1]) + list(all_ret[k].shape[1:])
        all_ret[k] = torch.reshape(all_ret[k], k_sh)
    k_extract = ['rgb_map', 'disp_map', 'acc_map']
    ret_list = [all_ret[k] for k in k_extract]
    ret_dict = {k: all_ret[k] for k in all_ret if k not in k_extract}
    return ret_list + [ret_dict]


_PATCHED = ("render_rays", "raw2outputs", "run_network", "batchify_rays", "render", "compute_weights",
            "compute_weights_piecewise_linear")
_PATCHED_HELPERS = ("NeRF", "get_embedder", "Embedder", "sample_pdf", "sample_pdf_reformulation", "sample_pdf_return_u",
                    "sample_pdf_reformulation_return_u")


def install(ref_run_plnerf, include_helpers=True):
    """Rebind the hot-path names inside an imported reference ``run_plnerf`` module.
    Returns the dict of replaced originals so that ``uninstall`` can restore them."""
    saved = {}
    g = globals()
    for name in _PATCHED:
        saved[name] = getattr(ref_run_plnerf, name, None)
        setattr(ref_run_plnerf, name, g[name])
    if include_helpers:
        for name in _PATCHED_HELPERS:
            saved[name] = getattr(ref_run_plnerf, name, None)
            setattr(ref_run_plnerf, name, getattr(helpers, name))
    if hasattr(ref_run_plnerf, "extract_fields"):   # nerf_extract_mesh.py carries its own copy of the path + the grid query
        saved["extract_fields"] = ref_run_plnerf.extract_fields
        ref_run_plnerf.extract_fields = nerf_extract_mesh.extract_fields
    return saved


def uninstall(ref_run_plnerf, saved):
    for name, fn in saved.items():
        if fn is not None:
            setattr(ref_run_plnerf, name, fn)
