"""Drop-in for the reference's ``run_nerf_helpers`` module on the hot path.

Same exported names, constructor signatures, parameter names/shapes and return conventions as the
reference (run_nerf_helpers.py), with the arithmetic done by the sm_100a kernels behind the C ABI:

* ``Embedder`` / ``get_embedder``               run_nerf_helpers.py:24-72
* ``NeRF`` (nn.Module, same state_dict)         run_nerf_helpers.py:76-157
* ``sample_pdf``                                run_nerf_helpers.py:241-284
* ``pw_linear_sample_increasing/decreasing``    run_nerf_helpers.py:340-361 (host formulas for API parity)
* ``sample_pdf_reformulation``                  run_nerf_helpers.py:364-445

plus the small utilities run_plnerf.py star-imports (img2mse, mse2psnr, to8b, to16b, get_rays,
get_rays_np, ndc_rays) which are plain torch/numpy one-liners outside the accelerated path.
CUDA tensors only: there is no CPU implementation in this package.
"""
import numpy as np
import torch
import torch.nn as nn

from . import ops

# Misc (run_nerf_helpers.py:17-20) -- not on the accelerated path, kept for star-import parity
img2mse = lambda x, y: torch.mean((x - y) ** 2)
mse2psnr = lambda x: -10. * torch.log(x) / torch.log(torch.tensor([10.], device=x.device))
to8b = lambda x: (255 * np.clip(x, 0, 1)).astype(np.uint8)
to16b = lambda x: ((2 ** 16 - 1) * np.clip(x, 0, 1)).astype(np.uint16)


class Embedder:
    """run_nerf_helpers.py:24-54.  Only the configuration get_embedder() builds is implemented
    (include_input, log-sampled power-of-two bands, [sin, cos])."""

    def __init__(self, **kwargs):
        self.kwargs = kwargs
        if not kwargs.get("include_input", True) or not kwargs.get("log_sampling", True) \
                or kwargs.get("input_dims", 3) != 3 \
                or kwargs.get("max_freq_log2") != kwargs.get("num_freqs") - 1:
            raise NotImplementedError("plnerf_b200 Embedder implements get_embedder()'s configuration only")
        self.multires = int(kwargs["num_freqs"])
        self.out_dim = 3 + 6 * self.multires

    def embed(self, inputs):
        return ops.encode(inputs, self.multires)


def get_embedder(multires, i=0):
    """run_nerf_helpers.py:57-72."""
    if i == -1:
        return nn.Identity(), 3
    embed_kwargs = {'include_input': True, 'input_dims': 3, 'max_freq_log2': multires - 1,
                    'num_freqs': multires, 'log_sampling': True, 'periodic_fns': [torch.sin, torch.cos]}
    embedder_obj = Embedder(**embed_kwargs)
    embed = lambda x, eo=embedder_obj: eo.embed(x)
    embed.multires = multires
    return embed, embedder_obj.out_dim


class NeRF(nn.Module):
    """Same module tree / state_dict as the reference NeRF (run_nerf_helpers.py:76-103); forward
    (:105-128) runs the fused tcgen05 kernel on the packed copy of the parameters."""

    def __init__(self, D=8, W=256, input_ch=3, input_ch_views=3, output_ch=4, skips=[4], use_viewdirs=False):
        super(NeRF, self).__init__()
        self.D = D
        self.W = W
        self.input_ch = input_ch
        self.input_ch_views = input_ch_views
        self.skips = skips
        self.use_viewdirs = use_viewdirs
        self.pts_linears = nn.ModuleList(
            [nn.Linear(input_ch, W)] + [nn.Linear(W, W) if i not in self.skips else nn.Linear(W + input_ch, W)
                                        for i in range(D - 1)])
        self.views_linears = nn.ModuleList([nn.Linear(input_ch_views + W, W // 2)])
        if use_viewdirs:
            self.feature_linear = nn.Linear(W, W)
            self.alpha_linear = nn.Linear(W, 1)
            self.rgb_linear = nn.Linear(W // 2, 3)
        else:
            self.output_linear = nn.Linear(W, output_ch)

    def forward(self, x):
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            from .autograd import mlp_forward_autograd
            return mlp_forward_autograd(self, x)
        return ops.mlp_forward(self, x)

    def load_weights_from_keras(self, weights):
        """run_nerf_helpers.py:130-157."""
        assert self.use_viewdirs, "Not implemented if use_viewdirs=False"
        for i in range(self.D):
            self.pts_linears[i].weight.data = torch.from_numpy(np.transpose(weights[2 * i]))
            self.pts_linears[i].bias.data = torch.from_numpy(np.transpose(weights[2 * i + 1]))
        k = 2 * self.D
        self.feature_linear.weight.data = torch.from_numpy(np.transpose(weights[k]))
        self.feature_linear.bias.data = torch.from_numpy(np.transpose(weights[k + 1]))
        self.views_linears[0].weight.data = torch.from_numpy(np.transpose(weights[k + 2]))
        self.views_linears[0].bias.data = torch.from_numpy(np.transpose(weights[k + 3]))
        self.rgb_linear.weight.data = torch.from_numpy(np.transpose(weights[k + 4]))
        self.rgb_linear.bias.data = torch.from_numpy(np.transpose(weights[k + 5]))
        self.alpha_linear.weight.data = torch.from_numpy(np.transpose(weights[k + 6]))
        self.alpha_linear.bias.data = torch.from_numpy(np.transpose(weights[k + 7]))


# Ray helpers (run_nerf_helpers.py:162-201) -- adjacent to the path ("next" row f-1), plain torch
def get_rays(H, W, K, c2w):
    i, j = torch.meshgrid(torch.linspace(0, W - 1, W, device=c2w.device),
                          torch.linspace(0, H - 1, H, device=c2w.device), indexing="ij")
    i, j = i.t(), j.t()
    dirs = torch.stack([(i - K[0][2]) / K[0][0], -(j - K[1][2]) / K[1][1], -torch.ones_like(i)], -1)
    rays_d = torch.sum(dirs[..., np.newaxis, :] * c2w[:3, :3], -1)
    rays_o = c2w[:3, -1].expand(rays_d.shape)
    return rays_o, rays_d


def get_rays_np(H, W, K, c2w):
    i, j = np.meshgrid(np.arange(W, dtype=np.float32), np.arange(H, dtype=np.float32), indexing='xy')
    dirs = np.stack([(i - K[0][2]) / K[0][0], -(j - K[1][2]) / K[1][1], -np.ones_like(i)], -1)
    rays_d = np.sum(dirs[..., np.newaxis, :] * c2w[:3, :3], -1)
    rays_o = np.broadcast_to(c2w[:3, -1], np.shape(rays_d))
    return rays_o, rays_d


def ndc_rays(H, W, focal, near, rays_o, rays_d):
    t = -(near + rays_o[..., 2]) / rays_d[..., 2]
    rays_o = rays_o + t[..., None] * rays_d
    o0 = -1. / (W / (2. * focal)) * rays_o[..., 0] / rays_o[..., 2]
    o1 = -1. / (H / (2. * focal)) * rays_o[..., 1] / rays_o[..., 2]
    o2 = 1. + 2. * near / rays_o[..., 2]
    d0 = -1. / (W / (2. * focal)) * (rays_d[..., 0] / rays_d[..., 2] - rays_o[..., 0] / rays_o[..., 2])
    d1 = -1. / (H / (2. * focal)) * (rays_d[..., 1] / rays_d[..., 2] - rays_o[..., 1] / rays_o[..., 2])
    d2 = -2. * near / rays_o[..., 2]
    return torch.stack([o0, o1, o2], -1), torch.stack([d0, d1, d2], -1)


def _draw_u(shape, N_samples, det, pytest, device):
    """The u the reference would use (run_nerf_helpers.py:248-264 / 376-392)."""
    if pytest:
        np.random.seed(0)
        if det:
            u = np.broadcast_to(np.linspace(0., 1., N_samples), shape)
        else:
            u = np.random.rand(*shape)
        return torch.Tensor(np.ascontiguousarray(u)).to(device)
    if det:
        return torch.linspace(0., 1., steps=N_samples, device=device).expand(shape).contiguous()
    return None  # drawn on device (Philox)


def _rays_with_bounds(near, far):
    rays = torch.zeros((near.shape[0], 8), device=near.device, dtype=torch.float32)
    rays[:, 6:7] = near
    rays[:, 7:8] = far
    return rays


def sample_pdf(bins, weights, N_samples, det=False, pytest=False, u=None, seed=0):
    """run_nerf_helpers.py:241-284 (piecewise-constant inverse CDF)."""
    n = bins.shape[0]
    if u is None:
        u = _draw_u([n, N_samples], N_samples, det, pytest, bins.device)
    return ops.sample_pdf(bins, weights, N_samples, u=u, seed=seed)


def pw_linear_sample_increasing(s_left, s_right, T_left, tau_left, tau_right, u, epsilon=1e-3):
    """run_nerf_helpers.py:340-349 (elementwise closed form; the fused kernel has its own copy)."""
    e = torch.ones_like(T_left) * epsilon
    ln_term = -torch.log(torch.max(e, torch.div(1 - u, torch.max(e, T_left))))
    disc = tau_left ** 2 + torch.div(2 * (tau_right - tau_left) * ln_term, torch.max(e, s_right - s_left))
    t = torch.div((s_right - s_left) * (-tau_left + torch.sqrt(torch.max(e, disc))), torch.max(e, tau_right - tau_left))
    t = torch.clamp(t, e, s_right - s_left)
    return s_left + t


def pw_linear_sample_decreasing(s_left, s_right, T_left, tau_left, tau_right, u, epsilon=1e-3):
    """run_nerf_helpers.py:352-361."""
    e = torch.ones_like(T_left) * epsilon
    ln_term = -torch.log(torch.max(e, torch.div(1 - u, torch.max(e, T_left))))
    disc = tau_left ** 2 - torch.div(2 * (tau_left - tau_right) * ln_term, torch.max(e, s_right - s_left))
    t = torch.div((s_right - s_left) * (tau_left - torch.sqrt(torch.max(e, disc))), torch.max(e, tau_left - tau_right))
    t = torch.clamp(t, e, s_right - s_left)
    return s_left + t


def sample_pdf_reformulation(bins, weights, tau, T, near, far, N_samples, det=False, pytest=False,
                             quad_solution_v2=False, zero_threshold=1e-4, epsilon_=1e-3, u=None, seed=0):
    """run_nerf_helpers.py:364-445.  Returns (samples, T_below, tau_below, bin_below) like the
    reference; the last three are gathered on the host side from the returned indices."""
    n = bins.shape[0]
    if u is None:
        u = _draw_u([n, N_samples], N_samples, det, pytest, bins.device)
    rays = _rays_with_bounds(near.reshape(n, 1), far.reshape(n, 1))
    samples, inds = ops.sample_pdf_pl(bins, weights, tau, T, rays, N_samples, u=u, seed=seed,
                                      zero_tol=zero_threshold, epsilon=epsilon_, return_inds=True)
    below = torch.clamp(inds - 1, min=0)
    full_bins = torch.cat([near.reshape(n, 1), bins, far.reshape(n, 1)], -1)
    return samples, torch.gather(T, 1, below), torch.gather(tau, 1, below), torch.gather(full_bins, 1, below)


# ---- f-4: the depth-experiment sampler variants (run_nerf_helpers.py:286-337, 448-533), forward --------------------------
_return_u_calls = 0


def _load_or_draw(load_u, shape, N_samples, det, pytest, device):
    if load_u is not None:
        return load_u.to(device).contiguous(), 0
    global _return_u_calls
    _return_u_calls += 1
    return _draw_u(shape, N_samples, det, pytest, device), (torch.initial_seed() * 0x9E3779B97F4A7C15 + _return_u_calls) & ((1 << 63) - 1)


class _SamplePdfReturnU(torch.autograd.Function):
    """sample_pdf_return_u with the gradient autograd gives the reference (bins directly, weights through the cdf)."""

    @staticmethod
    def forward(ctx, bins, weights, N_samples, u, seed):
        samples, u_used, _ = ops.sample_pdf_return_u(bins, weights, N_samples, load_u=u, seed=seed)
        ctx.save_for_backward(bins, weights, u_used)
        ctx.mark_non_differentiable(u_used)
        return samples, u_used

    @staticmethod
    def backward(ctx, g_samples, _g_u):
        bins, weights, u = ctx.saved_tensors
        g_bins, g_w = ops.sample_pdf_return_u_bwd(bins, weights, u, g_samples)
        return g_bins, g_w, None, None, None


def sample_pdf_return_u(bins, weights, N_samples, det=False, pytest=False, load_u=None):
    """run_nerf_helpers.py:286-337: sample_pdf that takes a saved u (``load_u``) and returns (samples, u).  Differentiable
    like the reference's (the depth experiments back-propagate through the samples): gradients reach ``bins`` and
    ``weights`` (plnerf_sample_pdf_return_u_bwd)."""
    u, seed = _load_or_draw(load_u, [bins.shape[0], N_samples], N_samples, det, pytest, bins.device)
    if torch.is_grad_enabled() and (bins.requires_grad or weights.requires_grad):
        return _SamplePdfReturnU.apply(bins, weights, N_samples, u, seed)
    samples, u_used, _ = ops.sample_pdf_return_u(bins, weights, N_samples, load_u=u, seed=seed)
    return samples, u_used


class _SamplePdfPLReturnU(torch.autograd.Function):
    """sample_pdf_reformulation_return_u with the gradient autograd gives the reference: every sample sends gradient to the
    two knots of its bracket (bins / near / far, tau, T); the weights only choose the bracket (searchsorted) and get none."""

    @staticmethod
    def forward(ctx, bins, weights, tau, T, near, far, N_samples, u, seed, zero_tol, eps):
        n = bins.shape[0]
        rays = _rays_with_bounds(near.reshape(n, 1), far.reshape(n, 1))
        samples, T_b, tau_b, bin_b, u_used, _ = ops.sample_pdf_pl_return_u(bins, weights, tau, T, rays, N_samples, load_u=u,
                                                                            seed=seed, zero_tol=zero_tol, epsilon=eps)
        ctx.save_for_backward(bins, weights, tau, T, rays, u_used)
        ctx.cfg = (zero_tol, eps, near.shape, far.shape)
        ctx.mark_non_differentiable(u_used)
        return samples, T_b, tau_b, bin_b, u_used

    @staticmethod
    def backward(ctx, g_samples, g_T_b, g_tau_b, g_bin_b, _g_u):
        bins, weights, tau, T, rays, u = ctx.saved_tensors
        zero_tol, eps, near_shape, far_shape = ctx.cfg
        g_z, g_near, g_far, g_tau, g_T = ops.sample_pdf_pl_return_u_bwd(bins, weights, tau, T, rays, u, g_samples, g_T_b, g_tau_b,
                                                                       g_bin_b, zero_tol=zero_tol, epsilon=eps)
        return g_z, None, g_tau, g_T, g_near.reshape(near_shape), g_far.reshape(far_shape), None, None, None, None, None


def sample_pdf_reformulation_return_u(bins, weights, tau, T, near, far, N_samples, det=False, pytest=False, load_u=None,
                                      quad_solution_v2=True, zero_threshold=1e-4, epsilon_=1e-3):
    """run_nerf_helpers.py:448-533: returns (samples, T_below, tau_below, bin_below, u).  Differentiable like the
    reference's: gradients of all four outputs reach ``bins``, ``near``, ``far``, ``tau`` and ``T``
    (plnerf_sample_pdf_pl_return_u_bwd)."""
    n = bins.shape[0]
    u, seed = _load_or_draw(load_u, [n, N_samples], N_samples, det, pytest, bins.device)
    if torch.is_grad_enabled() and any(t.requires_grad for t in (bins, tau, T, near, far)):
        return _SamplePdfPLReturnU.apply(bins, weights, tau, T, near, far, N_samples, u, seed, zero_threshold, epsilon_)
    rays = _rays_with_bounds(near.reshape(n, 1), far.reshape(n, 1))
    samples, T_b, tau_b, bin_b, u_used, _ = ops.sample_pdf_pl_return_u(bins, weights, tau, T, rays, N_samples, load_u=u, seed=seed,
                                                                        zero_tol=zero_threshold, epsilon=epsilon_)
    return samples, T_b, tau_b, bin_b, u_used


def compute_space_carving_loss_corrected(pred_depth, target_hypothesis, is_joint=False, mask=None, norm_p=2, threshold=0.0):
    """Space-carving depth loss of the depth-supervised experiments (run_nerf_helpers.py:203-238), plain torch (not on
    the accelerated path; present because run_plnerf.py star-imports the module).  pred_depth [n_rays, n_points];
    target_hypothesis [n_hyp, n_rays, 1 | n_points].  Per sample, the distance to the closest hypothesis (``is_joint``:
    the hypothesis is chosen per image from the ray-averaged distances), averaged."""
    n_points = pred_depth.shape[1]
    hyp = target_hypothesis.repeat(1, 1, n_points) if target_hypothesis.shape[-1] == 1 else target_hypothesis
    dist = torch.norm((pred_depth.unsqueeze(-1) - hyp.unsqueeze(-1)), p=norm_p, dim=-1)        # [n_hyp, n_rays, n_points]
    if mask is not None:
        dist = dist * mask.unsqueeze(0).repeat(dist.shape[0], 1).unsqueeze(-1)
    if threshold > 0:
        dist = torch.where(dist < threshold, torch.zeros((), device=dist.device, dtype=dist.dtype), dist)
    if is_joint:
        return torch.mean(torch.min(torch.mean(dist, dim=1), dim=0)[0], dim=-1)
    return torch.mean(torch.mean(torch.min(dist, dim=0)[0], dim=-1))


# ---- utilities for evaluation (run_nerf_helpers.py:537-570), star-imported by run_plnerf.py ------------------------------
def compute_rmse(prediction, target):
    """Root-mean-square error of two tensors (run_nerf_helpers.py:537-538)."""
    return ((prediction - target) ** 2).mean().sqrt()


class MeanTracker(object):
    """Running weighted means of the entries of metric dicts (run_nerf_helpers.py:541-570): same methods
    (add / has / get / as_dict / reset / print) and the same update rule."""

    def __init__(self):
        self.reset()

    def reset(self):
        self.mean_dict = {}
        self.total_weight = 0

    def add(self, input, weight=1.):
        new_total = self.total_weight + weight
        for key, value in input.items():
            old = self.mean_dict.get(key, 0)
            self.mean_dict[key] = (old * self.total_weight + value) / new_total
        self.total_weight = new_total

    def has(self, key):
        return key in self.mean_dict

    def get(self, key):
        return self.mean_dict[key]

    def as_dict(self):
        return self.mean_dict

    def print(self, f=None):
        for key, value in self.mean_dict.items():
            print("{}: {}".format(key, value), **({} if f is None else {"file": f}))
