"""torch-tensor front end of the C ABI: every function takes/returns CUDA float32 tensors and
enqueues the sm_100a kernels on ``torch.cuda.current_stream()``.  torch is only the allocator and
stream provider here.  CPU tensors are rejected: there is no CPU path.
"""
import ctypes as C
import os

import torch

from . import _lib as L

_PREC = {"bf16": L.PREC_BF16, "bf16x3": L.PREC_BF16X3}
_default_precision = os.environ.get("PLNERF_PRECISION", "bf16")


def set_precision(name):
    """'bf16' (one tcgen05 MMA per product, default) or 'bf16x3' (hi/lo split, ~fp32 products)."""
    global _default_precision
    if name not in _PREC:
        raise ValueError(f"precision must be one of {list(_PREC)}")
    _default_precision = name


def get_precision():
    return _default_precision


def _prec(p):
    return _PREC[p if p is not None else _default_precision]


def _stream():
    # (the raw handle of torch's current stream on the current device; torch.cuda.current_stream() builds a Stream object
    # per call, which shows in the issue time of a training step's ~10 calls)
    return C.c_void_p(torch._C._cuda_getCurrentRawStream(torch.cuda.current_device()))


def _f32(t, name):
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name}: expected a torch.Tensor")
    if not t.is_cuda:
        raise RuntimeError(f"{name}: plnerf_b200 only runs on CUDA tensors (no CPU fallback); got device {t.device}")
    if t.device.index != torch.cuda.current_device():
        # the kernels are enqueued on the CURRENT device's current stream (one process per GPU is the intended use)
        raise RuntimeError(f"{name}: tensor lives on {t.device} but the current CUDA device is cuda:{torch.cuda.current_device()}; "
                           "wrap the call in torch.cuda.device(...)")
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _p(t):
    return C.c_void_p(0 if t is None else t.data_ptr())


def launch_count():
    return int(L.lib().plnerf_launch_count())


def profile_enable(on=True):
    """Start/stop CUDA-event timing of every fused-MLP launch (bench.py roofline leg)."""
    L.check(L.lib().plnerf_profile_enable(int(bool(on))))


def profile_read():
    """-> (summed k_mlp_fwd device ms, launches, rows) since profile_enable(True)."""
    ms, n, rows = C.c_double(0), C.c_int64(0), C.c_int64(0)
    L.check(L.lib().plnerf_profile_read(C.byref(ms), C.byref(n), C.byref(rows)))
    return ms.value, n.value, rows.value


# ------------------------------------------------------------------------------------------------
def encode(x, multires):
    """Embedder.embed (run_nerf_helpers.py:53-54): [..., 3] -> [..., 3+6*multires]."""
    x = _f32(x, "x")
    shp = x.shape
    x2 = x.reshape(-1, 3)
    od = 3 if multires < 0 else 3 + 6 * multires
    out = torch.empty((x2.shape[0], od), device=x.device, dtype=torch.float32)
    L.check(L.lib().plnerf_encode(_p(x2), x2.shape[0], multires, _p(out), _stream()))
    return out.reshape(*shp[:-1], od)


def pack_rays(H, W, K, c2w=None, rays=None, ndc=True, near=0., far=1., use_viewdirs=False, c2w_staticcam=None):
    """The ray generation + packing of render() (run_plnerf.py:138-164) as one kernel: get_rays from a pose (or the given
    (rays_o, rays_d)), viewdirs before NDC, ndc_rays(H, W, K[0][0], 1., ...), near/far columns.
    Returns (packed [n, 8|11] fp32, shape of rays_d) -- `shape` is what render() reshapes its outputs to."""
    fx, fy, cx, cy = float(K[0][0]), float(K[1][1]), float(K[0][2]), float(K[1][2])
    ndc_cx, ndc_cy = -1. / (W / (2. * fx)), -1. / (H / (2. * fx))          # python float64 arithmetic like the reference
    if c2w is not None:
        c2w = _f32(c2w, "c2w")
        dev, n, sh = c2w.device, H * W, (H, W, 3)
        ro = rd = None
        if c2w_staticcam is not None:
            c2w_staticcam = _f32(c2w_staticcam, "c2w_staticcam")
    else:
        ro, rd = _f32(rays[0], "rays_o"), _f32(rays[1], "rays_d")
        sh = tuple(rd.shape)
        ro, rd = ro.reshape(-1, 3), rd.reshape(-1, 3)
        dev, n = rd.device, rd.shape[0]
    out = torch.empty((n, 11 if use_viewdirs else 8), device=dev, dtype=torch.float32)
    L.check(L.lib().plnerf_pack_rays(int(H), int(W), fx, fy, cx, cy, _p(c2w), 0 if c2w is None else c2w.stride(0),
                                      _p(c2w_staticcam), 0 if c2w_staticcam is None else c2w_staticcam.stride(0),
                                      _p(ro), _p(rd), n, int(bool(ndc)), ndc_cx, ndc_cy, 1.0, float(near), float(far),
                                      int(bool(use_viewdirs)), _p(out), out.shape[1], _stream()))
    return out, sh


def pack_pixel_rays(H, W, K, c2w, pix, ndc=True, near=0., far=1., use_viewdirs=False):
    """The training loop's per-iteration ray selection (run_plnerf.py:1259-1280) followed by render()'s packing
    (:145-164) as one kernel: packed rays [n, 8|11] of the pixels ``pix`` (int64 flat ids row*W + col) of the
    camera ``c2w`` -- row i is bit-identical to row pix[i] of ``pack_rays(H, W, K, c2w=c2w, ...)``; the full
    image's rays and the [H*W, 2] coordinate grid are never built."""
    fx, fy, cx, cy = float(K[0][0]), float(K[1][1]), float(K[0][2]), float(K[1][2])
    ndc_cx, ndc_cy = -1. / (W / (2. * fx)), -1. / (H / (2. * fx))
    c2w = _f32(c2w, "c2w")
    if not pix.is_cuda or pix.dtype != torch.int64 or pix.dim() != 1:
        raise RuntimeError("pack_pixel_rays: pix must be a 1-D int64 CUDA tensor")
    pix = pix.contiguous()
    n = pix.shape[0]
    out = torch.empty((n, 11 if use_viewdirs else 8), device=c2w.device, dtype=torch.float32)
    L.check(L.lib().plnerf_pack_pixel_rays(int(H), int(W), fx, fy, cx, cy, _p(c2w), c2w.stride(0), _p(pix), n,
                                            int(bool(ndc)), ndc_cx, ndc_cy, 1.0, float(near), float(far),
                                            int(bool(use_viewdirs)), _p(out), out.shape[1], _stream()))
    return out


def stratified_z(rays, N_samples, lindisp=False, perturb=True, t_rand=None, seed=0, ray_id_offset=0):
    """render_rays' depth sampling (run_plnerf.py:683-705).  rays [n, >=8]."""
    rays = _f32(rays, "rays")
    n = rays.shape[0]
    z = torch.empty((n, N_samples), device=rays.device, dtype=torch.float32)
    if t_rand is not None:
        t_rand = _f32(t_rand, "t_rand")
        assert t_rand.shape == (n, N_samples)
    L.check(L.lib().plnerf_stratified_z(_p(rays), n, rays.shape[1], N_samples, int(bool(lindisp)),
                                         int(bool(perturb)), _p(t_rand), seed, ray_id_offset, _p(z), _stream()))
    return z


def raw2outputs(raw, z_vals, rays, mode, color_mode, noise=None, white_bkgd=False, farcolorfix=False,
                want_weights=True):
    """raw2outputs (run_plnerf.py:553-624).  rays [n, >=8] carries rays_d (cols 3-5), near, far (6,7).
    Returns (rgb_map, disp_map, acc_map, weights, depth_map, tau, T)."""
    raw = _f32(raw, "raw")
    z_vals = _f32(z_vals, "z_vals")
    rays = _f32(rays, "rays")
    n, S = z_vals.shape
    dev = raw.device
    lin = mode == "linear"
    if mode not in ("linear", "constant"):
        raise ValueError(f"mode must be 'linear' or 'constant', got {mode!r}")
    if color_mode not in ("midpoint", "left"):
        raise ValueError(f"color_mode must be 'midpoint' or 'left', got {color_mode!r}")
    rgb = torch.empty((n, 3), device=dev)
    disp = torch.empty((n,), device=dev)
    acc = torch.empty((n,), device=dev)
    depth = torch.empty((n,), device=dev)
    w = torch.empty((n, S + 1 if lin else S), device=dev) if want_weights else None
    tau = torch.empty((n, S + 2), device=dev) if (lin and want_weights) else None
    T = torch.empty((n, S + 2), device=dev) if (lin and want_weights) else None
    if noise is not None:
        noise = _f32(noise, "noise")
    L.check(L.lib().plnerf_raw2outputs(_p(raw), raw.shape[-1], _p(z_vals), _p(rays), n, rays.shape[1], S,
                                        L.MODE_LINEAR if lin else L.MODE_CONSTANT,
                                        L.COLOR_MIDPOINT if color_mode == "midpoint" else L.COLOR_LEFT,
                                        int(bool(white_bkgd)), int(bool(farcolorfix)), _p(noise), _p(rgb), _p(disp),
                                        _p(acc), _p(depth), _p(w), _p(tau), _p(T), _stream()))
    return rgb, disp, acc, w, depth, tau, T


def raw2outputs_bwd(raw, z_vals, rays, mode, color_mode, g_rgb=None, g_depth=None, g_acc=None, g_disp=None,
                    noise=None, white_bkgd=False, farcolorfix=False):
    """Gradient of raw2outputs w.r.t. raw given upstream grads of (rgb_map, depth_map, acc_map, disp_map)."""
    raw, z_vals, rays = _f32(raw, "raw"), _f32(z_vals, "z_vals"), _f32(rays, "rays")
    n, S = z_vals.shape
    if mode not in ("linear", "constant"):
        raise ValueError(f"mode must be 'linear' or 'constant', got {mode!r}")
    opt = lambda t, nm: None if t is None else _f32(t, nm)
    g_rgb, g_depth, g_acc, g_disp, noise = (opt(g_rgb, "g_rgb"), opt(g_depth, "g_depth"), opt(g_acc, "g_acc"),
                                            opt(g_disp, "g_disp"), opt(noise, "noise"))
    g_raw = torch.empty_like(raw)
    L.check(L.lib().plnerf_raw2outputs_bwd(_p(raw), raw.shape[-1], _p(z_vals), _p(rays), n, rays.shape[1], S,
                                            L.MODE_LINEAR if mode == "linear" else L.MODE_CONSTANT,
                                            L.COLOR_MIDPOINT if color_mode == "midpoint" else L.COLOR_LEFT,
                                            int(bool(white_bkgd)), int(bool(farcolorfix)), _p(noise), _p(g_rgb),
                                            _p(g_depth), _p(g_acc), _p(g_disp), _p(g_raw), _stream()))
    return g_raw


def sample_pdf_pl(z_vals, weights, tau, T, rays, N_importance, u=None, seed=0, ray_id_offset=0, zero_tol=1e-4,
                  epsilon=1e-3, return_inds=False):
    """sample_pdf_reformulation (run_nerf_helpers.py:364-445) with near/far taken from rays cols 6,7."""
    z_vals, weights, tau, T, rays = (_f32(t, k) for t, k in ((z_vals, "z_vals"), (weights, "weights"),
                                                              (tau, "tau"), (T, "T"), (rays, "rays")))
    n, S = z_vals.shape
    out = torch.empty((n, N_importance), device=z_vals.device)
    inds = torch.empty((n, N_importance), device=z_vals.device, dtype=torch.int64) if return_inds else None
    if u is not None:
        u = _f32(u, "u")
    L.check(L.lib().plnerf_sample_pdf_pl(_p(z_vals), _p(weights), _p(tau), _p(T), _p(rays), n, rays.shape[1], S,
                                          N_importance, _p(u), seed, ray_id_offset, zero_tol, epsilon, _p(out),
                                          _p(inds), _stream()))
    return (out, inds) if return_inds else out


def sample_pdf(bins, weights, N_importance, u=None, seed=0, ray_id_offset=0, return_inds=False):
    """sample_pdf (run_nerf_helpers.py:241-284): bins [n,nb], weights [n,nb-1]."""
    bins, weights = _f32(bins, "bins"), _f32(weights, "weights")
    n, nb = bins.shape
    assert weights.shape == (n, nb - 1)
    out = torch.empty((n, N_importance), device=bins.device)
    inds = torch.empty((n, N_importance), device=bins.device, dtype=torch.int64) if return_inds else None
    if u is not None:
        u = _f32(u, "u")
    L.check(L.lib().plnerf_sample_pdf(_p(bins), _p(weights), n, nb, N_importance, _p(u), seed, ray_id_offset,
                                       _p(out), _p(inds), _stream()))
    return (out, inds) if return_inds else out


def sample_pdf_pl_return_u(z_vals, weights, tau, T, rays, N_importance, load_u=None, seed=0, ray_id_offset=0,
                           zero_tol=1e-4, epsilon=1e-3):
    """sample_pdf_reformulation_return_u (run_nerf_helpers.py:448-533), forward: -> (samples, T_below, tau_below,
    bin_below, u, inds), each [n, N_importance]."""
    z_vals, weights, tau, T, rays = (_f32(t, k) for t, k in ((z_vals, "z_vals"), (weights, "weights"),
                                                              (tau, "tau"), (T, "T"), (rays, "rays")))
    n, S = z_vals.shape
    dev = z_vals.device
    outs = [torch.empty((n, N_importance), device=dev) for _ in range(5)]
    inds = torch.empty((n, N_importance), device=dev, dtype=torch.int64)
    if load_u is not None:
        load_u = _f32(load_u, "load_u")
    L.check(L.lib().plnerf_sample_pdf_pl_return_u(_p(z_vals), _p(weights), _p(tau), _p(T), _p(rays), n, rays.shape[1], S,
                                                   N_importance, _p(load_u), seed, ray_id_offset, zero_tol, epsilon,
                                                   _p(outs[0]), _p(outs[1]), _p(outs[2]), _p(outs[3]), _p(outs[4]),
                                                   _p(inds), _stream()))
    return (*outs, inds)


def sample_pdf_return_u(bins, weights, N_importance, load_u=None, seed=0, ray_id_offset=0):
    """sample_pdf_return_u (run_nerf_helpers.py:286-337), forward: -> (samples, u, inds)."""
    bins, weights = _f32(bins, "bins"), _f32(weights, "weights")
    n, nb = bins.shape
    assert weights.shape == (n, nb - 1)
    out = torch.empty((n, N_importance), device=bins.device)
    u_out = torch.empty((n, N_importance), device=bins.device)
    inds = torch.empty((n, N_importance), device=bins.device, dtype=torch.int64)
    if load_u is not None:
        load_u = _f32(load_u, "load_u")
    L.check(L.lib().plnerf_sample_pdf_return_u(_p(bins), _p(weights), n, nb, N_importance, _p(load_u), seed, ray_id_offset,
                                                _p(out), _p(u_out), _p(inds), _stream()))
    return out, u_out, inds


def sample_pdf_pl_return_u_bwd(z_vals, weights, tau, T, rays, u, g_samples=None, g_T_below=None, g_tau_below=None,
                               g_bin_below=None, zero_tol=1e-4, epsilon=1e-3):
    """Backward of sample_pdf_reformulation_return_u (run_nerf_helpers.py:448-533) as autograd computes it: cotangents
    [n, Ni] (None = zero) -> (g_z [n,S], g_near [n], g_far [n], g_tau [n,S+2], g_T [n,S+2]).  ``u`` = the forward's draws."""
    z_vals, weights, tau, T, rays, u = (_f32(t, k) for t, k in ((z_vals, "z_vals"), (weights, "weights"), (tau, "tau"),
                                                                 (T, "T"), (rays, "rays"), (u, "u")))
    n, S = z_vals.shape
    Ni = u.shape[1]
    dev = z_vals.device
    gs = [None if g is None else _f32(g, "cotangent") for g in (g_samples, g_T_below, g_tau_below, g_bin_below)]
    g_z, g_near, g_far = torch.empty((n, S), device=dev), torch.empty((n,), device=dev), torch.empty((n,), device=dev)
    g_tau, g_T = torch.empty((n, S + 2), device=dev), torch.empty((n, S + 2), device=dev)
    L.check(L.lib().plnerf_sample_pdf_pl_return_u_bwd(_p(z_vals), _p(weights), _p(tau), _p(T), _p(rays), n, rays.shape[1], S, Ni,
                                                       _p(u), zero_tol, epsilon, _p(gs[0]), _p(gs[1]), _p(gs[2]), _p(gs[3]),
                                                       _p(g_z), _p(g_near), _p(g_far), _p(g_tau), _p(g_T), _stream()))
    return g_z, g_near, g_far, g_tau, g_T


def sample_pdf_return_u_bwd(bins, weights, u, g_samples):
    """Backward of sample_pdf_return_u (run_nerf_helpers.py:286-337): -> (g_bins [n,nb], g_weights [n,nb-1])."""
    bins, weights, u, g_samples = (_f32(t, k) for t, k in ((bins, "bins"), (weights, "weights"), (u, "u"),
                                                           (g_samples, "g_samples")))
    n, nb = bins.shape
    assert weights.shape == (n, nb - 1) and g_samples.shape == u.shape
    g_bins, g_w = torch.empty((n, nb), device=bins.device), torch.empty((n, nb - 1), device=bins.device)
    L.check(L.lib().plnerf_sample_pdf_return_u_bwd(_p(bins), _p(weights), n, nb, u.shape[1], _p(u), _p(g_samples),
                                                    _p(g_bins), _p(g_w), _stream()))
    return g_bins, g_w


def merge_samples(z_vals, z_samples, rays):
    """clamp + sort(cat) + std (run_plnerf.py:728-734, :752) -> (z_merged, z_std)."""
    z_vals, z_samples, rays = _f32(z_vals, "z_vals"), _f32(z_samples, "z_samples"), _f32(rays, "rays")
    n, S = z_vals.shape
    Ni = z_samples.shape[1]
    out = torch.empty((n, S + Ni), device=z_vals.device)
    std = torch.empty((n,), device=z_vals.device)
    L.check(L.lib().plnerf_merge_samples(_p(z_vals), _p(z_samples), _p(rays), n, rays.shape[1], S, Ni, _p(out),
                                          _p(std), _stream()))
    return out, std


# ------------------------------------------------------------------------------------------------
# packed networks
# ------------------------------------------------------------------------------------------------
def net_desc_of(net):
    """plnerf_net_desc from a NeRF-like module (attrs of run_nerf_helpers.py:81-86)."""
    d = L.NetDesc()
    d.D, d.W = int(net.D), int(net.W)
    d.input_ch, d.input_ch_views = int(net.input_ch), int(net.input_ch_views)
    d.use_viewdirs = int(bool(net.use_viewdirs))
    d.output_ch = int(net.output_linear.out_features) if not net.use_viewdirs else 4
    skips = list(net.skips)
    d.n_skips = len(skips)
    for i, s in enumerate(skips):
        d.skips[i] = int(s)
    return d


class PackedNet:
    """Device-side packed copy of one NeRF's parameters (bf16 K-major panels + fp32 tail),
    refreshed lazily whenever a parameter's version counter changes (e.g. after optimizer.step())."""

    def __init__(self, net):
        self.desc = net_desc_of(net)
        self.buf = {}
        self.version = {}
        self._params = None
        self._prm = None

    def _versions(self, net):
        # the module's Parameter objects, listed once (walking net.parameters() costs ~25 us a time; a NeRF's parameter set
        # is fixed, `.to()` / `load_state_dict` keep the objects): (storage address, version counter) of each
        if self._params is None:
            self._params = list(net.parameters())
        return tuple((p.data_ptr(), p._version) for p in self._params)

    def get(self, net, precision):
        prec = _prec(precision)
        ver = self._versions(net)
        if self.version.get(prec) == ver and prec in self.buf:
            return self.buf[prec]
        first = next(net.parameters())
        if not first.is_cuda:
            raise RuntimeError("plnerf_b200: the NeRF module must live on a CUDA device (no CPU fallback)")
        nbytes = L.lib().plnerf_packed_bytes(C.byref(self.desc), prec)
        if nbytes == 0:
            L.check(-2)
        if prec not in self.buf or self.buf[prec].numel() != nbytes or self.buf[prec].device != first.device:
            self.buf[prec] = torch.empty(nbytes, dtype=torch.uint8, device=first.device)
        prm, keep = self.param_struct(net, ver)
        L.check(L.lib().plnerf_pack_weights(C.byref(self.desc), C.byref(prm), prec, _p(self.buf[prec]), _stream()))
        del keep
        self.version[prec] = ver
        return self.buf[prec]

    def param_struct(self, net, ver=None):
        """plnerf_net_params of the module's parameters (+ the tensors it points into).  Rebuilt only when a parameter's
        storage moved: a training loop repacks after every optimiser step, with the same addresses each time."""
        ver = ver if ver is not None else self._versions(net)
        addrs = tuple(v[0] for v in ver)
        cached = self._prm
        if cached is not None and cached[0] == addrs:
            return cached[1], cached[2]
        prm = L.NetParams()
        keep = []
        converted = _fill_params(prm, self.desc, {k: v.detach() for k, v in net.named_parameters()}, keep, convert=True)
        if not converted:                      # (converted copies are snapshots: never cached)
            self._prm = (addrs, prm, keep)
        return prm, keep


def _fill_params(struct, desc, tensors, keep, convert=False):
    """Fill a NetParams / NetGrads ctypes struct from a {state_dict name: tensor} mapping.  ``convert``: tensors that are
    not contiguous float32 are copied (parameters of a half / strided module) instead of refused; returns whether any was."""
    converted = [False]

    def ptr(name):
        t = tensors[name]
        if t.dtype != torch.float32 or not t.is_contiguous():
            if not convert:
                raise RuntimeError(f"{name}: expected a contiguous float32 tensor")
            t = t.float().contiguous()
            converted[0] = True
        keep.append(t)
        return t.data_ptr()
    for i in range(desc.D):
        struct.pts_w[i] = ptr(f"pts_linears.{i}.weight")
        struct.pts_b[i] = ptr(f"pts_linears.{i}.bias")
    if desc.use_viewdirs:
        for f, nme in (("views_w", "views_linears.0.weight"), ("views_b", "views_linears.0.bias"),
                       ("feature_w", "feature_linear.weight"), ("feature_b", "feature_linear.bias"),
                       ("alpha_w", "alpha_linear.weight"), ("alpha_b", "alpha_linear.bias"),
                       ("rgb_w", "rgb_linear.weight"), ("rgb_b", "rgb_linear.bias")):
            setattr(struct, f, ptr(nme))
    else:
        struct.output_w = ptr("output_linear.weight")
        struct.output_b = ptr("output_linear.bias")
    return converted[0]


def packed_bwd_of(net):
    """Transposed (input-gradient) weight stream of a network, cached like PackedNet."""
    pk = packed_of(net)
    ver = pk._versions(net)
    cache = net.__dict__.setdefault("_plnerf_packed_bwd", {})
    if cache.get("ver") == ver:
        return cache["buf"]
    nbytes = L.lib().plnerf_packed_bwd_bytes(C.byref(pk.desc))
    if nbytes == 0:
        L.check(-2)
    first = next(net.parameters())
    buf = cache.get("buf")
    if buf is None or buf.numel() != nbytes or buf.device != first.device:
        buf = torch.empty(nbytes, dtype=torch.uint8, device=first.device)
    prm, keep = pk.param_struct(net, ver)
    L.check(L.lib().plnerf_pack_weights_bwd(C.byref(pk.desc), C.byref(prm), _p(buf), _stream()))
    del keep
    cache["ver"], cache["buf"] = ver, buf
    return buf


def repack_train(nets):
    """Refresh the packed bf16 copy AND the transposed (input-gradient) copy of 1 or 2 networks in ONE launch
    (plnerf_pack_weights_train) and mark both caches current: what a training loop does after ``optimizer.step()``
    instead of letting the next forward / backward repack lazily (three launches per network)."""
    if not 1 <= len(nets) <= 2:
        raise ValueError("repack_train: 1 or 2 networks")
    n = len(nets)
    descs, prms = (C.POINTER(L.NetDesc) * n)(), (C.POINTER(L.NetParams) * n)()
    bufs, bwds = (C.c_void_p * n)(), (C.c_void_p * n)()
    keep_all, done = [], []
    prec = _prec("bf16")
    for i, net in enumerate(nets):
        pk = packed_of(net)
        ver = pk._versions(net)
        first = pk._params[0]
        if not first.is_cuda:
            raise RuntimeError("plnerf_b200: the NeRF module must live on a CUDA device (no CPU fallback)")
        nb, nbb = L.lib().plnerf_packed_bytes(C.byref(pk.desc), prec), L.lib().plnerf_packed_bwd_bytes(C.byref(pk.desc))
        if nb == 0 or nbb == 0:
            L.check(-2)
        buf = pk.buf.get(prec)
        if buf is None or buf.numel() != nb or buf.device != first.device:
            buf = pk.buf[prec] = torch.empty(nb, dtype=torch.uint8, device=first.device)
        cache = net.__dict__.setdefault("_plnerf_packed_bwd", {})
        bwd = cache.get("buf")
        if bwd is None or bwd.numel() != nbb or bwd.device != first.device:
            bwd = cache["buf"] = torch.empty(nbb, dtype=torch.uint8, device=first.device)
        prm, keep = pk.param_struct(net, ver)
        keep_all.append((prm, keep))
        descs[i], prms[i] = C.pointer(pk.desc), C.pointer(prm)
        bufs[i], bwds[i] = _p(buf), _p(bwd)
        done.append((pk, cache, ver))
    with torch.cuda.device(done[0][0]._params[0].device):
        L.check(L.lib().plnerf_pack_weights_train(n, descs, prms, bufs, bwds, _stream()))
    for pk, cache, ver in done:
        pk.version[prec] = ver
        cache["ver"] = ver
    del keep_all


def network_query_train(net, rays, z_vals):
    """run_network forward that also stashes what the backward needs.  -> (raw [n,S,4], stash)."""
    rays, z_vals = _f32(rays, "rays"), _f32(z_vals, "z_vals")
    pk = packed_of(net)
    buf = pk.get(net, "bf16")
    n, S = z_vals.shape
    raw = torch.empty((n, S, 4 if pk.desc.use_viewdirs else pk.desc.output_ch), device=rays.device)
    sb = L.lib().plnerf_train_stash_bytes(C.byref(pk.desc), n, S)
    if sb == 0:
        L.check(-2)
    if sb > 64 * 2 ** 30:
        raise RuntimeError(f"plnerf_b200: training stash for {n} rays x {S} samples would need {sb / 2**30:.1f} GiB; "
                           "use a smaller ray batch / chunk for calls that require gradients")
    stash = torch.empty(sb + 1024, dtype=torch.uint8, device=rays.device)
    off = (-stash.data_ptr()) % 1024
    wsb = L.lib().plnerf_query_workspace_bytes(C.byref(pk.desc), n)
    ws = torch.empty(wsb, dtype=torch.uint8, device=rays.device)
    mr = _multires_of(pk.desc.input_ch)
    mrv = _multires_of(pk.desc.input_ch_views) if pk.desc.use_viewdirs else -1
    L.check(L.lib().plnerf_network_query_train(C.byref(pk.desc), _p(buf), mr, mrv, _p(rays), n, rays.shape[1],
                                                _p(z_vals), S, _p(raw), C.c_void_p(stash.data_ptr() + off), sb,
                                                _p(ws), wsb, _stream()))
    return raw, (stash, off, sb)


def zero_grads_like(net):
    """{state_dict name: zeroed fp32 gradient} as views of ONE flat buffer (the weight-gradient kernels accumulate with
    atomics, so the buffers must start at zero; a fill kernel per parameter is 24 launches per network and step)."""
    params = list(net.named_parameters())
    flat = torch.zeros(sum(p.numel() for _, p in params), device=params[0][1].device, dtype=torch.float32)
    out, off = {}, 0
    for k, p in params:
        out[k] = flat[off:off + p.numel()].view(p.shape)
        off += p.numel()
    return out


def network_query_bwd(net, g_raw, stash, n, S, grads=None):
    """Parameter gradients of one network query.  g_raw [n,S,>=4].  Returns {state_dict name: grad} (fp32);
    `grads` may supply zero-initialised / accumulating buffers."""
    stash_t, off, sb = stash
    g_raw = _f32(g_raw, "g_raw")
    pk = packed_of(net)
    buf = pk.get(net, "bf16")
    bwd = packed_bwd_of(net)
    if grads is None:
        grads = zero_grads_like(net)
    gs = L.NetGrads()
    keep = []
    _fill_params(gs, pk.desc, grads, keep)
    L.check(L.lib().plnerf_network_query_bwd(C.byref(pk.desc), _p(buf), _p(bwd), n, S, _p(g_raw), g_raw.shape[-1],
                                              C.c_void_p(stash_t.data_ptr() + off), sb, C.byref(gs), _stream()))
    return grads


def mse_loss_grad(rgb, rgb0, target, scale, sqerr, pix=None):
    """img2mse of the fine and the coarse map and the gradient loss.backward() starts from (run_plnerf.py:1289-1297):
    returns (g_rgb, g_rgb0) = scale * (rgb - t), scale * (rgb0 - t) and ADDS the two sums of squared errors to
    ``sqerr`` [2] (device tensor).  ``target`` [*, 3]: row i is the target of ray i, or row pix[i] when ``pix`` (int64 ids)
    is given.  rgb0 may be None."""
    rgb = _f32(rgb, "rgb")
    target = _f32(target, "target")
    n = rgb.shape[0]
    g = torch.empty_like(rgb)
    g0 = None
    if rgb0 is not None:
        rgb0 = _f32(rgb0, "rgb0")
        g0 = torch.empty_like(rgb0)
    if pix is not None and (pix.dtype != torch.int64 or not pix.is_contiguous() or not pix.is_cuda):
        raise RuntimeError("pix: expected a contiguous int64 CUDA tensor")
    if sqerr.dtype != torch.float32 or sqerr.numel() < 2 or not sqerr.is_cuda:
        raise RuntimeError("sqerr: expected a float32 CUDA tensor of 2 elements")
    L.check(L.lib().plnerf_mse_loss_grad(_p(rgb), _p(rgb0) if rgb0 is not None else None, _p(target),
                                         _p(pix) if pix is not None else None, n, float(scale), _p(g),
                                         _p(g0) if g0 is not None else None, _p(sqerr), _stream()))
    return g, g0


def adam_step(params, grads, exp_avg, exp_avg_sq, lr, step, betas=(0.9, 0.999), eps=1e-8, zero_grads=False):
    """One torch.optim.Adam update (amsgrad=False, weight_decay=0) of a flat fp32 segment, in place; ``step`` is the
    1-based count of this update."""
    for t, name in ((params, "params"), (grads, "grads"), (exp_avg, "exp_avg"), (exp_avg_sq, "exp_avg_sq")):
        if not t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous() or t.numel() != params.numel():
            raise RuntimeError(f"{name}: expected a contiguous float32 CUDA tensor of {params.numel()} elements")
    L.check(L.lib().plnerf_adam_step(_p(params), _p(grads), _p(exp_avg), _p(exp_avg_sq), params.numel(), float(lr),
                                     float(betas[0]), float(betas[1]), float(eps), int(step), int(bool(zero_grads)), _stream()))


def packed_of(net):
    """Per-module PackedNet cache stored on the module itself."""
    pk = net.__dict__.get("_plnerf_packed")
    if pk is None:
        pk = PackedNet(net)
        net.__dict__["_plnerf_packed"] = pk
    return pk


def invalidate_packed(net, params_replaced=False):
    """Forget the packed copies of a network whose parameters were modified through an alias (e.g. a flat buffer
    that the optimizer updates in place): the parameters' own version counters do not move in that case.
    ``params_replaced``: the module's Parameter OBJECTS were exchanged (not just their values or storage): also forget
    the cached parameter list."""
    pk = net.__dict__.get("_plnerf_packed")
    if pk is not None:
        pk.version.clear()
        if params_replaced:
            pk._params = None
            pk._prm = None
    cache = net.__dict__.get("_plnerf_packed_bwd")
    if cache is not None:
        cache.pop("ver", None)


def _multires_of(ch):
    if ch == 3:
        return -1
    if ch < 3 or (ch - 3) % 6:
        raise RuntimeError(f"plnerf_b200: {ch} input channels is not a NeRF positional encoding (3 + 6*L)")
    return (ch - 3) // 6


def network_query(net, rays, z_vals, precision=None):
    """run_network (run_plnerf.py:78-92) fused: rays [n, 8|11], z_vals [n,S] -> raw [n,S,C]."""
    rays, z_vals = _f32(rays, "rays"), _f32(z_vals, "z_vals")
    pk = packed_of(net)
    buf = pk.get(net, precision)
    n, S = z_vals.shape
    ch = 4 if pk.desc.use_viewdirs else pk.desc.output_ch
    raw = torch.empty((n, S, ch), device=rays.device)
    wsb = L.lib().plnerf_query_workspace_bytes(C.byref(pk.desc), n)
    ws = torch.empty(wsb, dtype=torch.uint8, device=rays.device)
    mr = _multires_of(pk.desc.input_ch)
    mrv = _multires_of(pk.desc.input_ch_views) if pk.desc.use_viewdirs else -1
    L.check(L.lib().plnerf_network_query(C.byref(pk.desc), _p(buf), _prec(precision), mr, mrv, _p(rays), n,
                                          rays.shape[1], _p(z_vals), S, _p(raw), _p(ws), wsb, _stream()))
    return raw


def mlp_forward(net, x, precision=None):
    """NeRF.forward (run_nerf_helpers.py:105-128) on embedded rows x [m, input_ch+input_ch_views]."""
    x = _f32(x, "x")
    pk = packed_of(net)
    buf = pk.get(net, precision)
    width = pk.desc.input_ch + (pk.desc.input_ch_views if pk.desc.use_viewdirs else 0)
    if x.shape[-1] < width:
        raise RuntimeError(f"NeRF.forward: expected at least {width} input channels, got {x.shape[-1]}")
    x2 = x.reshape(-1, x.shape[-1])[:, :width].contiguous()
    m = x2.shape[0]
    ch = 4 if pk.desc.use_viewdirs else pk.desc.output_ch
    out = torch.empty((m, ch), device=x.device)
    wsb = L.lib().plnerf_query_workspace_bytes(C.byref(pk.desc), m)
    ws = torch.empty(wsb, dtype=torch.uint8, device=x.device)
    L.check(L.lib().plnerf_mlp_forward(C.byref(pk.desc), _p(buf), _prec(precision), _p(x2), m, _p(out), _p(ws), wsb,
                                        _stream()))
    return out.reshape(*x.shape[:-1], ch)


def _render_cfg(pkc, N_samples, N_importance, mode, color_mode, perturb, white_bkgd, lindisp, raw_noise_std, zero_tol,
                epsilon, farcolorfix, seed, ray_id_offset, precision):
    if mode not in ("linear", "constant"):
        raise ValueError(f"mode must be 'linear' or 'constant', got {mode!r}")
    if color_mode not in ("midpoint", "left"):
        raise ValueError(f"color_mode must be 'midpoint' or 'left', got {color_mode!r}")
    cfg = L.RenderCfg()
    cfg.N_samples, cfg.N_importance = N_samples, N_importance
    cfg.mode = L.MODE_LINEAR if mode == "linear" else L.MODE_CONSTANT
    cfg.color_mode = L.COLOR_MIDPOINT if color_mode == "midpoint" else L.COLOR_LEFT
    cfg.white_bkgd, cfg.lindisp, cfg.farcolorfix = int(bool(white_bkgd)), int(bool(lindisp)), int(bool(farcolorfix))
    cfg.perturb = int(bool(perturb))
    cfg.raw_noise_std, cfg.zero_tol, cfg.epsilon = float(raw_noise_std), float(zero_tol), float(epsilon)
    cfg.multires = _multires_of(pkc.desc.input_ch)
    cfg.multires_views = _multires_of(pkc.desc.input_ch_views) if pkc.desc.use_viewdirs else -1
    cfg.precision = _prec(precision)
    cfg.seed, cfg.ray_id_offset = int(seed), int(ray_id_offset)
    return cfg


def render_rays_fwd_train(rays, net_coarse, net_fine, N_samples, N_importance, mode, color_mode, perturb=True,
                          white_bkgd=False, lindisp=False, raw_noise_std=0.0, zero_tol=1e-4, epsilon=1e-3, farcolorfix=False,
                          t_rand=None, u=None, noise0=None, noise1=None, seed=0, ray_id_offset=0, retraw=False):
    """render_rays forward that keeps what the backward needs (plnerf_render_rays_fwd_train): one C call.
    Returns (dict of outputs like render_rays_fwd, ctx) -- ctx goes to render_rays_bwd."""
    rays = _f32(rays, "rays")
    n, dev = rays.shape[0], rays.device
    pkc = packed_of(net_coarse)
    bufc = pkc.get(net_coarse, "bf16")
    fine = net_fine if (N_importance > 0 and net_fine is not None) else None
    pkf = packed_of(fine) if fine is not None else None
    buff = pkf.get(fine, "bf16") if fine is not None else None
    cfg = _render_cfg(pkc, N_samples, N_importance, mode, color_mode, perturb, white_bkgd, lindisp, raw_noise_std, zero_tol,
                      epsilon, farcolorfix, seed, ray_id_offset, "bf16")
    S_last = N_samples + N_importance
    o, ret = L.RenderOut(), {}

    def new(key, shape, dtype=torch.float32):
        t = torch.empty(shape, device=dev, dtype=dtype)
        setattr(o, key, t.data_ptr())
        ret[key] = t
    new("rgb_map", (n, 3)); new("disp_map", (n,)); new("acc_map", (n,)); new("depth_map", (n,))
    if retraw:
        last = pkf.desc if pkf else pkc.desc
        new("raw", (n, S_last, 4 if last.use_viewdirs else last.output_ch))
    if N_importance > 0:
        new("rgb0", (n, 3)); new("disp0", (n,)); new("depth0", (n,)); new("acc0", (n,)); new("z_std", (n,))
    opt = lambda t, nm: None if t is None else _f32(t, nm)
    t_rand, u, noise0, noise1 = opt(t_rand, "t_rand"), opt(u, "u"), opt(noise0, "noise0"), opt(noise1, "noise1")
    fdesc = C.byref(pkf.desc) if pkf else None
    wsb = L.lib().plnerf_render_train_workspace_bytes(C.byref(cfg), C.byref(pkc.desc), fdesc, n)
    if wsb == 0:
        L.check(-2)
    if wsb > 64 * 2 ** 30:
        raise RuntimeError(f"plnerf_b200: the training workspace for {n} rays x {S_last} samples would need {wsb / 2**30:.1f} GiB; "
                           "use a smaller ray batch / chunk for calls that require gradients")
    ws = torch.empty(wsb + 1024, dtype=torch.uint8, device=dev)
    off = (-ws.data_ptr()) % 1024
    L.check(L.lib().plnerf_render_rays_fwd_train(C.byref(cfg), C.byref(pkc.desc), _p(bufc), fdesc, _p(buff), _p(rays), n,
                                                  rays.shape[1], _p(t_rand), _p(u), _p(noise0), _p(noise1), C.byref(o),
                                                  C.c_void_p(ws.data_ptr() + off), wsb, _stream()))
    ctx = dict(cfg=cfg, rays=rays, n=n, ws=ws, off=off, wsb=wsb, net_c=net_coarse, net_f=fine, noise0=noise0, noise1=noise1)
    return ret, ctx


class PreparedGrads:
    """The gradient buffers of one network ({state_dict name: fp32 tensor}) as the plnerf_net_grads struct the backward
    entries take, built once (a training step hands over the same views of its flat gradient buffer every iteration)."""

    def __init__(self, net, grads):
        self.tensors = []
        self.struct = L.NetGrads()
        _fill_params(self.struct, packed_of(net).desc, grads, self.tensors)


def render_rays_bwd(ctx, g_fine, g_coarse, grads_c, grads_f):
    """Backward of render_rays_fwd_train (plnerf_render_rays_bwd): g_fine / g_coarse = (g_rgb, g_disp, g_acc, g_depth) of the
    fine / coarse maps (entries or the whole tuple may be None; with N_importance == 0 the maps are g_fine's).  Parameter
    gradients are ADDED into grads_c / grads_f ({state_dict name: fp32 tensor}, or a PreparedGrads of it)."""
    net_c, net_f = ctx["net_c"], ctx["net_f"]
    pkc = packed_of(net_c)
    bufc, bwdc = pkc.get(net_c, "bf16"), packed_bwd_of(net_c)
    keep = []
    gc = (grads_c if isinstance(grads_c, PreparedGrads) else PreparedGrads(net_c, grads_c)).struct
    fdesc = buff = bwdf = gf_ref = None
    if net_f is not None:
        pkf = packed_of(net_f)
        fdesc, buff, bwdf = C.byref(pkf.desc), pkf.get(net_f, "bf16"), packed_bwd_of(net_f)
        gf = (grads_f if isinstance(grads_f, PreparedGrads) else PreparedGrads(net_f, grads_f)).struct
        gf_ref = C.byref(gf)
    g = L.RenderGrads()
    c = lambda t: None if t is None else _f32(t, "upstream gradient")
    for names, tup in ((("g_rgb_map", "g_disp_map", "g_acc_map", "g_depth_map"), g_fine),
                       (("g_rgb0", "g_disp0", "g_acc0", "g_depth0"), g_coarse)):
        for nm, t in zip(names, tup if tup is not None else (None,) * 4):
            t = c(t)
            keep.append(t)
            setattr(g, nm, None if t is None else t.data_ptr())
    rays = ctx["rays"]
    L.check(L.lib().plnerf_render_rays_bwd(C.byref(ctx["cfg"]), C.byref(pkc.desc), _p(bufc), _p(bwdc), fdesc, _p(buff), _p(bwdf),
                                            _p(rays), ctx["n"], rays.shape[1], _p(ctx["noise0"]), _p(ctx["noise1"]), C.byref(g),
                                            C.byref(gc), gf_ref, C.c_void_p(ctx["ws"].data_ptr() + ctx["off"]), ctx["wsb"],
                                            _stream()))


def train_rays_mse(rays, net_coarse, net_fine, N_samples, N_importance, mode, color_mode, target, scale, sqerr, grads_c,
                   grads_f, pix=None, perturb=True, white_bkgd=False, lindisp=False, raw_noise_std=0.0, zero_tol=1e-4,
                   epsilon=1e-3, farcolorfix=False, t_rand=None, u=None, noise0=None, noise1=None, seed=0, ray_id_offset=0,
                   want_maps=False):
    """The device work of one iteration of the reference loop on a ray batch -- render_rays forward (stash mode), the two
    img2mse terms, loss.backward() -- in ONE C call (plnerf_train_rays_mse); the coarse pass's loss and backward run on a
    side stream beside the fine pass.  ``target`` / ``pix`` / ``scale`` / ``sqerr`` as in mse_loss_grad; parameter gradients
    are ADDED into grads_c / grads_f (dicts or PreparedGrads).  Returns the maps as a dict when ``want_maps``, else None."""
    rays, target = _f32(rays, "rays"), _f32(target, "target")
    n, dev = rays.shape[0], rays.device
    fine = net_fine if (N_importance > 0 and net_fine is not None) else None
    pkc = packed_of(net_coarse)
    bufc, bwdc = pkc.get(net_coarse, "bf16"), packed_bwd_of(net_coarse)
    gc = (grads_c if isinstance(grads_c, PreparedGrads) else PreparedGrads(net_coarse, grads_c)).struct
    fdesc = buff = bwdf = gf_ref = None
    if fine is not None:
        pkf = packed_of(fine)
        fdesc, buff, bwdf = C.byref(pkf.desc), pkf.get(fine, "bf16"), packed_bwd_of(fine)
        gf = (grads_f if isinstance(grads_f, PreparedGrads) else PreparedGrads(fine, grads_f)).struct
        gf_ref = C.byref(gf)
    cfg = _render_cfg(pkc, N_samples, N_importance, mode, color_mode, perturb, white_bkgd, lindisp, raw_noise_std, zero_tol,
                      epsilon, farcolorfix, seed, ray_id_offset, "bf16")
    if pix is not None and (pix.dtype != torch.int64 or not pix.is_contiguous() or not pix.is_cuda):
        raise RuntimeError("pix: expected a contiguous int64 CUDA tensor")
    if sqerr.dtype != torch.float32 or sqerr.numel() < 2 or not sqerr.is_cuda:
        raise RuntimeError("sqerr: expected a float32 CUDA tensor of 2 elements")
    o, ret = None, None
    if want_maps:
        o, ret = L.RenderOut(), {}
        keys = [("rgb_map", (n, 3)), ("disp_map", (n,)), ("acc_map", (n,)), ("depth_map", (n,))]
        if N_importance > 0:
            keys += [("rgb0", (n, 3)), ("disp0", (n,)), ("acc0", (n,)), ("depth0", (n,)), ("z_std", (n,))]
        for key, shape in keys:
            ret[key] = torch.empty(shape, device=dev)
            setattr(o, key, ret[key].data_ptr())
    opt = lambda t, nm: None if t is None else _f32(t, nm)
    t_rand, u, noise0, noise1 = opt(t_rand, "t_rand"), opt(u, "u"), opt(noise0, "noise0"), opt(noise1, "noise1")
    wsb = L.lib().plnerf_render_train_workspace_bytes(C.byref(cfg), C.byref(pkc.desc), fdesc, n)
    if wsb == 0:
        L.check(-2)
    if wsb > 64 * 2 ** 30:
        raise RuntimeError(f"plnerf_b200: the training workspace for {n} rays x {N_samples + N_importance} samples would need "
                           f"{wsb / 2**30:.1f} GiB; use a smaller ray batch / chunk")
    ws = torch.empty(wsb + 1024, dtype=torch.uint8, device=dev)
    off = (-ws.data_ptr()) % 1024
    L.check(L.lib().plnerf_train_rays_mse(C.byref(cfg), C.byref(pkc.desc), _p(bufc), _p(bwdc), fdesc, _p(buff), _p(bwdf), _p(rays),
                                           n, rays.shape[1], _p(t_rand), _p(u), _p(noise0), _p(noise1), _p(target), _p(pix),
                                           float(scale), _p(sqerr), C.byref(o) if o is not None else None, C.byref(gc), gf_ref,
                                           C.c_void_p(ws.data_ptr() + off), wsb, _stream()))
    return ret


def render_rays_fwd(rays, net_coarse, net_fine, N_samples, N_importance, mode, color_mode, perturb=True,
                    white_bkgd=False, lindisp=False, raw_noise_std=0.0, zero_tol=1e-4, epsilon=1e-3, farcolorfix=False,
                    t_rand=None, u=None, noise0=None, noise1=None, seed=0, ray_id_offset=0, retraw=False,
                    want_z=False, want_inds=False, precision=None):
    """Whole render_rays forward (run_plnerf.py:627-758) in one C call.  Returns the reference's dict."""
    rays = _f32(rays, "rays")
    n = rays.shape[0]
    dev = rays.device
    pkc = packed_of(net_coarse)
    bufc = pkc.get(net_coarse, precision)
    pkf, buff = None, None
    if N_importance > 0 and net_fine is not None:
        pkf = packed_of(net_fine)
        buff = pkf.get(net_fine, precision)
    cfg = L.RenderCfg()
    cfg.N_samples, cfg.N_importance = N_samples, N_importance
    cfg.mode = L.MODE_LINEAR if mode == "linear" else L.MODE_CONSTANT
    if mode not in ("linear", "constant"):
        raise ValueError(f"mode must be 'linear' or 'constant', got {mode!r}")
    if color_mode not in ("midpoint", "left"):
        raise ValueError(f"color_mode must be 'midpoint' or 'left', got {color_mode!r}")
    cfg.color_mode = L.COLOR_MIDPOINT if color_mode == "midpoint" else L.COLOR_LEFT
    cfg.white_bkgd, cfg.lindisp, cfg.farcolorfix = int(bool(white_bkgd)), int(bool(lindisp)), int(bool(farcolorfix))
    cfg.perturb = int(bool(perturb))
    cfg.raw_noise_std, cfg.zero_tol, cfg.epsilon = float(raw_noise_std), float(zero_tol), float(epsilon)
    cfg.multires = _multires_of(pkc.desc.input_ch)
    cfg.multires_views = _multires_of(pkc.desc.input_ch_views) if pkc.desc.use_viewdirs else -1
    cfg.precision = _prec(precision)
    cfg.seed, cfg.ray_id_offset = int(seed), int(ray_id_offset)
    S_last = N_samples + N_importance
    ch = 4 if pkc.desc.use_viewdirs else pkc.desc.output_ch
    o = L.RenderOut()
    ret = {}

    def new(key, shape, field=None, dtype=torch.float32):
        t = torch.empty(shape, device=dev, dtype=dtype)
        setattr(o, field or key, t.data_ptr())
        ret[key] = t
        return t
    new("rgb_map", (n, 3)); new("disp_map", (n,)); new("acc_map", (n,)); new("depth_map", (n,))
    if retraw:
        new("raw", (n, S_last, ch))
    if N_importance > 0:
        new("rgb0", (n, 3)); new("disp0", (n,)); new("depth0", (n,)); new("acc0", (n,)); new("z_std", (n,))
        if want_inds:
            new("inds", (n, N_importance), dtype=torch.int64)
    if want_z:
        new("z_vals", (n, S_last))
    opt = lambda t, nm: None if t is None else _f32(t, nm)
    t_rand, u, noise0, noise1 = opt(t_rand, "t_rand"), opt(u, "u"), opt(noise0, "noise0"), opt(noise1, "noise1")
    wsb = L.lib().plnerf_render_workspace_bytes(C.byref(cfg), C.byref(pkc.desc), n)
    ws = torch.empty(wsb + 256, dtype=torch.uint8, device=dev)
    off = (-ws.data_ptr()) % 256
    L.check(L.lib().plnerf_render_rays_fwd(C.byref(cfg), C.byref(pkc.desc), _p(bufc),
                                            C.byref(pkf.desc) if pkf else None, _p(buff), _p(rays), n, rays.shape[1],
                                            _p(t_rand), _p(u), _p(noise0), _p(noise1), C.byref(o),
                                            C.c_void_p(ws.data_ptr() + off), wsb, _stream()))
    return ret
