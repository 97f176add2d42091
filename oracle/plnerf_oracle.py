"""CPU oracle for the PL-NeRF ray-rendering hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A numpy (float32) restatement -- large MLP / encoding batches run through torch's CPU kernels (the BLAS and vector
math the reference itself uses, so the timed CPU baseline is not handicapped) -- of the reference's algorithm for the
path named by BASELINE.json
(render -> render_rays -> PE + coarse/fine MLP -> piecewise-linear quadrature -> inverse-CDF
sampler).  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this module; the product package never does.

Pinning: the reference has no tests or golden vectors of its own, so this oracle is
pinned against outputs of the reference itself, imported unmodified from /root/reference in the
build container by ``tests/golden/make_golden.py`` (fixtures committed under ``tests/golden``).
``tests/test_oracle_golden.py`` checks every function below against those fixtures.

Every function cites the reference file:line it follows (paths relative to /root/reference).
Arithmetic notes that matter for bit-parity of the integer outputs:
torch CPU ``cumsum``/``cumprod`` on float32 accumulate left-to-right in float64 and round each
prefix to float32 -- ``_cumsum32``/``_cumprod32`` do exactly that.
"""
import numpy as np

F32 = np.float32


def _f(x):
    return np.asarray(x, dtype=F32)


def _cumsum32(x):
    return np.cumsum(x.astype(np.float64), axis=-1).astype(F32)


def _cumprod32(x):
    return np.cumprod(x.astype(np.float64), axis=-1).astype(F32)


try:  # the reference runs its Linear layers through torch's CPU BLAS (MKL sgemm); use the same
    import torch as _torch  # library for the big GEMMs so that the timed CPU baseline is not handicapped
except Exception:  # pragma: no cover
    _torch = None
USE_TORCH_BLAS = _torch is not None


# --------------------------------------------------------------------------------------------
# Positional encoding -- run_nerf_helpers.py:24-72 (Embedder.embed / get_embedder)
# --------------------------------------------------------------------------------------------
def embed(x, multires):
    """[x, sin(2^0 x), cos(2^0 x), ..., sin(2^(L-1) x), cos(2^(L-1) x)], 3-wide blocks
    (run_nerf_helpers.py:36-54).  multires < 0 -> identity (i_embed == -1, :58-59)."""
    x = _f(x)
    if multires < 0:
        return x
    if USE_TORCH_BLAS and x.size >= 3 * 4096:
        # large batches: the same elementwise kernels the reference runs (torch CPU sin/cos, all host threads), so
        # that the timed CPU baseline is not handicapped by single-threaded numpy
        xt = _torch.from_numpy(np.ascontiguousarray(x))
        outs = [xt]
        for k in range(multires):
            xf = xt * float(2.0 ** k)
            outs.append(_torch.sin(xf))
            outs.append(_torch.cos(xf))
        return _torch.cat(outs, -1).numpy()
    outs = [x]
    for k in range(multires):
        freq = F32(2.0 ** k)
        xf = x * freq
        outs.append(np.sin(xf))
        outs.append(np.cos(xf))
    return np.concatenate(outs, -1).astype(F32)


def embed_dim(multires):
    return 3 if multires < 0 else 3 + 6 * multires


# --------------------------------------------------------------------------------------------
# MLP -- run_nerf_helpers.py:105-128 (NeRF.forward)
# --------------------------------------------------------------------------------------------


def _affine(h, w, b, relu=False):
    """(h @ w.T + b), optionally ReLU'd, in float32.  Large batches go through torch's CPU kernels
    (F.linear = MKL sgemm with fused bias, like the reference's nn.Linear); tiny ones stay numpy."""
    if USE_TORCH_BLAS and h.shape[0] >= 256:
        y = _torch.nn.functional.linear(_torch.from_numpy(np.ascontiguousarray(h)),
                                        _torch.from_numpy(np.ascontiguousarray(w)),
                                        None if b is None else _torch.from_numpy(np.ascontiguousarray(b)))
        if relu:
            y = _torch.relu_(y)
        return y.numpy()
    y = h @ w.T
    if b is not None:
        y = y + b
    return np.maximum(y, F32(0)) if relu else y


def _matmul_t(h, w):
    return _affine(h, w, None)


def _linear(h, params, name):
    return _affine(h, params[name + ".weight"], params[name + ".bias"])


def bf16_round(x):
    """Round float32 to the nearest bfloat16 (ties to even), returned as float32."""
    x = np.ascontiguousarray(x, dtype=F32)
    b = x.view(np.uint32)
    r = ((b + np.uint32(0x7FFF) + ((b >> np.uint32(16)) & np.uint32(1))) & np.uint32(0xFFFF0000)).astype(np.uint32)
    return r.view(F32).reshape(x.shape)


def nerf_forward(params, x, D=8, skips=(4,), input_ch=63, input_ch_views=27, use_viewdirs=True,
                 emulate_bf16=False):
    """x [M, input_ch+input_ch_views] -> [M,4] (viewdirs) or [M,output_ch]
    (run_nerf_helpers.py:105-128).

    emulate_bf16=True restates the arithmetic of the fast tcgen05 kernel (NOT of the reference):
    GEMM operands (activations, PE inputs, trunk/feature/views weights) rounded to bf16 with fp32
    accumulation; biases, the alpha / rgb / output_linear heads and the viewdir columns of
    views_linears stay fp32 and read the un-rounded fp32 activations."""
    x = _f(x)
    if USE_TORCH_BLAS and not emulate_bf16 and x.shape[0] >= 256:
        return _nerf_forward_large(params, x, D, skips, input_ch, input_ch_views, use_viewdirs)
    rnd = bf16_round if emulate_bf16 else (lambda a: a)
    input_pts, input_views = x[:, :input_ch], x[:, input_ch:input_ch + input_ch_views]
    pts_q = rnd(input_pts)
    h = pts_q
    h32 = None
    for i in range(D):
        w, b = params[f"pts_linears.{i}.weight"], params[f"pts_linears.{i}.bias"]
        h32 = _affine(h, rnd(w), b, relu=True)
        h = rnd(h32)
        if i in skips:
            h = np.concatenate([pts_q, h], -1)
    if use_viewdirs:
        alpha = _linear(h32, params, "alpha_linear")
        feature = rnd(_affine(h, rnd(params["feature_linear.weight"]), params["feature_linear.bias"]))
        wv, bv = params["views_linears.0.weight"], params["views_linears.0.bias"]
        W = feature.shape[1]
        hv = np.maximum(_matmul_t(feature, rnd(wv[:, :W])) + (_matmul_t(input_views, wv[:, W:]) + bv), F32(0))
        rgb = _linear(hv, params, "rgb_linear")
        return np.concatenate([rgb, alpha], -1).astype(F32)
    return _linear(h32, params, "output_linear").astype(F32)


def _nerf_forward_large(params, x, D, skips, input_ch, input_ch_views, use_viewdirs):
    """nerf_forward for large batches with every intermediate kept in torch CPU tensors (F.linear = MKL sgemm with
    fused bias, multi-threaded relu / cat): the operations and their order are those of the numpy branch above --
    and of the reference's forward (run_nerf_helpers.py:105-128) -- without the single-threaded numpy copies."""
    F_ = _torch.nn.functional
    t = lambda name: _torch.from_numpy(np.ascontiguousarray(params[name]))
    xt = _torch.from_numpy(np.ascontiguousarray(x))
    pts, views = xt[:, :input_ch], xt[:, input_ch:input_ch + input_ch_views]
    h = pts
    for i in range(D):
        h = _torch.relu_(F_.linear(h, t(f"pts_linears.{i}.weight"), t(f"pts_linears.{i}.bias")))
        if i in skips:
            h = _torch.cat([pts, h], -1)
    if not use_viewdirs:
        return F_.linear(h, t("output_linear.weight"), t("output_linear.bias")).numpy()
    alpha = F_.linear(h, t("alpha_linear.weight"), t("alpha_linear.bias"))
    feature = F_.linear(h, t("feature_linear.weight"), t("feature_linear.bias"))
    hv = _torch.relu_(F_.linear(_torch.cat([feature, views], -1), t("views_linears.0.weight"), t("views_linears.0.bias")))
    rgb = F_.linear(hv, t("rgb_linear.weight"), t("rgb_linear.bias"))
    return _torch.cat([rgb, alpha], -1).numpy()


def run_network(pts, viewdirs, params, multires=10, multires_views=4, netchunk=1024 * 64, **net_kw):
    """run_plnerf.py:68-92 (batchify + run_network): flatten, PE, broadcast+PE dirs, chunked MLP."""
    pts = _f(pts)
    flat = pts.reshape(-1, pts.shape[-1])
    embedded = embed(flat, multires)
    if viewdirs is not None:
        dirs = np.broadcast_to(_f(viewdirs)[:, None, :], pts.shape).reshape(-1, 3)
        embedded = np.concatenate([embedded, embed(dirs, multires_views)], -1)
    outs = [nerf_forward(params, embedded[i:i + netchunk], **net_kw)
            for i in range(0, embedded.shape[0], netchunk)]
    out = np.concatenate(outs, 0)
    return out.reshape(list(pts.shape[:-1]) + [out.shape[-1]])


def extract_fields(axes, params, multires=10, multires_views=4, block=64, **net_kw):
    """Density grid of the mesh extractor -- nerf_extract_mesh.py:531-562 (extract_fields) with its
    own run_network (:80-95, per-point viewdirs).

    ``axes`` = (X, Y, Z): the three ``torch.linspace(bound_min[k], bound_max[k], resolution)`` coordinate
    vectors (passed in so that the oracle, the reference and the CUDA path see identical coordinates).
    The grid is walked in ``block``-sized sub-cubes (N = 64, :532-535), each sub-cube's points are the
    'ij' meshgrid of its coordinate slices (:542-543), the view directions are all-zero rows (:545),
    and the stored value is relu(raw[..., 3]) (:555,:561).  Returns u [len(X), len(Y), len(Z)] float32."""
    X, Y, Z = (_f(a) for a in axes)
    u = np.zeros((len(X), len(Y), len(Z)), F32)
    for x0 in range(0, len(X), block):
        for y0 in range(0, len(Y), block):
            for z0 in range(0, len(Z), block):
                xs, ys, zs = X[x0:x0 + block], Y[y0:y0 + block], Z[z0:z0 + block]
                xx, yy, zz = np.meshgrid(xs, ys, zs, indexing="ij")
                pts = np.stack([xx.reshape(-1), yy.reshape(-1), zz.reshape(-1)], -1).astype(F32)
                emb = embed(pts, multires)
                if net_kw.get("use_viewdirs", True):
                    emb = np.concatenate([emb, embed(np.zeros_like(pts), multires_views)], -1)
                val = nerf_forward(params, emb, **net_kw).reshape(len(xs), len(ys), len(zs), -1)
                u[x0:x0 + len(xs), y0:y0 + len(ys), z0:z0 + len(zs)] = np.maximum(val[..., 3], F32(0))
    return u


# --------------------------------------------------------------------------------------------
# Quadrature -- run_plnerf.py:504-624
# --------------------------------------------------------------------------------------------
def _norm3(rays_d):
    return np.sqrt(np.sum(_f(rays_d) ** 2, -1, keepdims=True)).astype(F32)


def compute_weights(raw, z_vals, rays_d, noise=0.0):
    """Piecewise-constant weights, run_plnerf.py:504-513."""
    dists = z_vals[..., 1:] - z_vals[..., :-1]
    dists = np.concatenate([dists, np.full_like(dists[..., :1], 1e10)], -1)
    dists = dists * _norm3(rays_d)
    sigma = np.maximum(raw[..., 3] + _f(noise), F32(0))
    alpha = F32(1.0) - np.exp(-sigma * dists)
    trans = _cumprod32(np.concatenate([np.ones((alpha.shape[0], 1), F32),
                                       F32(1.0) - alpha + F32(1e-10)], -1))[:, :-1]
    return (alpha * trans).astype(F32)


def compute_weights_piecewise_linear(raw, z_vals, near, far, rays_d, noise=0.0):
    """PL weights, tau, T -- run_plnerf.py:516-550."""
    z = np.concatenate([near, z_vals, far], -1)
    dists = (z[..., 1:] - z[..., :-1]) * _norm3(rays_d)
    n = raw.shape[0]
    tau = np.concatenate([np.full((n, 1), 1e-10, F32), raw[..., 3] + _f(noise),
                          np.full((n, 1), 1e10, F32)], -1).astype(F32)
    tau = np.maximum(tau, F32(0))
    interval_ave_tau = F32(0.5) * (tau[..., 1:] + tau[..., :-1])
    expr = np.exp(-interval_ave_tau * dists).astype(F32)
    T = _cumprod32(np.concatenate([np.ones((n, 1), F32), expr], -1))
    weights = ((F32(1) - expr) * T[:, :-1]).astype(F32)
    return weights, tau, T


def _sigmoid(x):
    return (F32(1) / (F32(1) + np.exp(-x))).astype(F32)


def raw2outputs(raw, z_vals, near, far, rays_d, mode, color_mode, noise=0.0, white_bkgd=False,
                farcolorfix=False):
    """run_plnerf.py:553-624.  ``noise`` is the already-scaled additive density noise
    (randn*raw_noise_std in the reference, :569-576) or 0."""
    raw = _f(raw)
    z_vals = _f(z_vals)
    rgb = _sigmoid(raw[..., :3])
    if mode == "linear":
        weights, tau, T = compute_weights_piecewise_linear(raw, z_vals, near, far, rays_d, noise)
        if color_mode == "midpoint":
            last = np.zeros_like(rgb[:, -1:, :]) if farcolorfix else rgb[:, -1:, :]
            cc = np.concatenate([rgb[:, :1, :], rgb, last], 1)
            rgb_mid = F32(0.5) * (cc[:, 1:, :] + cc[:, :-1, :])
            rgb_map = np.sum(weights[..., None] * rgb_mid, -2, dtype=F32)
        elif color_mode == "left":
            cc = np.concatenate([rgb[:, :1, :], rgb], 1)
            rgb_map = np.sum(weights[..., None] * cc, -2, dtype=F32)
        else:
            raise ValueError(color_mode)
        zz = np.concatenate([near, z_vals, far], -1)
        z_mid = F32(0.5) * (zz[..., 1:] + zz[..., :-1])
        depth_map = np.sum(weights * z_mid, -1, dtype=F32)
    elif mode == "constant":
        weights = compute_weights(raw, z_vals, rays_d, noise)
        rgb_map = np.sum(weights[..., None] * rgb, -2, dtype=F32)
        depth_map = np.sum(weights * z_vals, -1, dtype=F32)
        tau = None
        T = None
    else:
        raise ValueError(mode)
    acc_map = np.sum(weights, -1, dtype=F32)
    with np.errstate(divide="ignore", invalid="ignore"):
        disp_map = (F32(1.0) / np.maximum(F32(1e-10), depth_map / acc_map)).astype(F32)
    if white_bkgd:
        rgb_map = rgb_map + (F32(1.0) - acc_map[..., None])
    return rgb_map.astype(F32), disp_map, acc_map, weights, depth_map, tau, T


# --------------------------------------------------------------------------------------------
# Samplers -- run_nerf_helpers.py:241-284 (constant) and :340-445 (PL)
# --------------------------------------------------------------------------------------------
def _searchsorted_right(cdf, u):
    """torch.searchsorted(cdf, u, right=True) per row, as ATen computes it: the upper-bound binary
    search  `mid = start + (end-start)//2; if !(cdf[mid] > u) start = mid+1 else end = mid`
    (identical to "first i with cdf[i] > u" on sorted rows, and well defined when the forced
    cdf[-1] = 1.0 makes the row non-monotone by an ulp)."""
    n, m = cdf.shape[0], cdf.shape[1]
    start = np.zeros(u.shape, np.int64)
    end = np.full(u.shape, m, np.int64)
    for _ in range(int(np.ceil(np.log2(m + 1))) + 1):
        active = start < end
        mid = start + ((end - start) >> 1)
        midc = np.minimum(mid, m - 1)
        val = np.take_along_axis(cdf, midc, -1)
        go_right = active & ~(val > u)
        go_left = active & (val > u)
        start = np.where(go_right, mid + 1, start)
        end = np.where(go_left, mid, end)
    return start


def sample_pdf(bins, weights, u):
    """Piecewise-constant inverse CDF, run_nerf_helpers.py:241-284, with the uniforms ``u``
    [N,Ni] supplied by the caller.  Returns (samples, inds)."""
    bins = _f(bins)
    u = _f(u)
    weights = _f(weights) + F32(1e-5)
    pdf = weights / np.sum(weights, -1, keepdims=True, dtype=F32)
    cdf = _cumsum32(pdf)
    cdf = np.concatenate([np.zeros_like(cdf[..., :1]), cdf], -1)
    inds = _searchsorted_right(cdf, u)
    below = np.maximum(0, inds - 1)
    above = np.minimum(cdf.shape[-1] - 1, inds)
    cdf_b = np.take_along_axis(cdf, below, -1)
    cdf_a = np.take_along_axis(cdf, above, -1)
    bins_b = np.take_along_axis(bins, below, -1)
    bins_a = np.take_along_axis(bins, above, -1)
    denom = cdf_a - cdf_b
    denom = np.where(denom < F32(1e-5), F32(1), denom)
    t = (u - cdf_b) / denom
    return (bins_b + t * (bins_a - bins_b)).astype(F32), inds


def _ln_term(T_left, u, eps):
    # run_nerf_helpers.py:341 / :353
    return -np.log(np.maximum(eps, (F32(1) - u) / np.maximum(eps, T_left)))


def pw_linear_sample_increasing(s_left, s_right, T_left, tau_left, tau_right, u, epsilon=1e-3):
    """run_nerf_helpers.py:340-349."""
    eps = F32(epsilon)
    ln_term = _ln_term(T_left, u, eps)
    disc = tau_left ** 2 + (F32(2) * (tau_right - tau_left) * ln_term) / np.maximum(eps, s_right - s_left)
    t = ((s_right - s_left) * (-tau_left + np.sqrt(np.maximum(eps, disc)))) / np.maximum(eps, tau_right - tau_left)
    t = np.minimum(np.maximum(t, eps), s_right - s_left)   # torch.clamp(min, max): max wins
    return (s_left + t).astype(F32)


def pw_linear_sample_decreasing(s_left, s_right, T_left, tau_left, tau_right, u, epsilon=1e-3):
    """run_nerf_helpers.py:352-361."""
    eps = F32(epsilon)
    ln_term = _ln_term(T_left, u, eps)
    disc = tau_left ** 2 - (F32(2) * (tau_left - tau_right) * ln_term) / np.maximum(eps, s_right - s_left)
    t = ((s_right - s_left) * (tau_left - np.sqrt(np.maximum(eps, disc)))) / np.maximum(eps, tau_left - tau_right)
    t = np.minimum(np.maximum(t, eps), s_right - s_left)
    return (s_left + t).astype(F32)


def sample_pdf_reformulation(bins, weights, tau, T, near, far, u, zero_threshold=1e-4, epsilon_=1e-3):
    """PL inverse-CDF sampler, run_nerf_helpers.py:364-445, uniforms supplied.
    Returns (samples, inds) with inds the int64 searchsorted result (:397)."""
    bins = np.concatenate([_f(near), _f(bins), _f(far)], -1)
    u = _f(u)
    cdf = _cumsum32(_f(weights))
    cdf = np.concatenate([np.zeros_like(cdf[..., :1]), cdf], -1)
    cdf[:, -1] = 1.0
    inds = _searchsorted_right(cdf, u)
    below = np.maximum(0, inds - 1)
    above = np.minimum(cdf.shape[-1] - 1, inds)
    g = lambda a, i: np.take_along_axis(a, i, -1)
    s_left, s_right = g(bins, below), g(bins, above)
    T_left = g(T, below)
    tau_left, tau_right = g(tau, below), g(tau, above)
    tau_diff = tau[..., 1:] - tau[..., :-1]
    if below.max(initial=0) >= tau_diff.shape[-1]:
        # the reference raises here too (SURVEY.md 8c caveat 3: u == 1.0 with det=True)
        raise IndexError("index out of range in tau_diff gather (u must be < 1)")
    tau_diff_g = g(tau_diff, below)
    zt = F32(zero_threshold)
    with np.errstate(invalid="ignore", divide="ignore", over="ignore"):
        inc = pw_linear_sample_increasing(s_left, s_right, T_left, tau_left, tau_right, u, epsilon_)
        dec = pw_linear_sample_decreasing(s_left, s_right, T_left, tau_left, tau_right, u, epsilon_)
    samples = np.where((tau_diff_g < zt) & (tau_diff_g > -zt), s_left, F32(-1.0))
    samples = np.where(tau_diff_g >= zt, inc, samples)
    samples = np.where(tau_diff_g <= -zt, dec, samples)
    samples = np.where(np.isnan(samples), s_left, samples)
    return samples.astype(F32), inds


def sample_pdf_return_u(bins, weights, u):
    """run_nerf_helpers.py:286-337 with the uniforms (`load_u` or the draw) supplied: (samples, u)."""
    return sample_pdf(bins, weights, u)[0], _f(u)


def sample_pdf_reformulation_return_u(bins, weights, tau, T, near, far, u, zero_threshold=1e-4, epsilon_=1e-3):
    """run_nerf_helpers.py:448-533 with the uniforms supplied: the samples of sample_pdf_reformulation plus T, tau and the
    knot gathered at the lower bracket index (:523-529) and u -> (samples, T_below, tau_below, bin_below, u)."""
    samples, inds = sample_pdf_reformulation(bins, weights, tau, T, near, far, u, zero_threshold, epsilon_)
    knots = np.concatenate([_f(near), _f(bins), _f(far)], -1)
    below = np.maximum(0, inds - 1)
    g = lambda a: np.take_along_axis(_f(a), below, -1)
    return samples, g(T), g(tau), g(knots), _f(u)


def _max_grad(c, x):
    """d max(c, x) / dx as torch.max(tensor, tensor) routes it: 1 where x > c, 1/2 at ties, 0 below."""
    return np.where(x > c, F32(1), np.where(x == c, F32(0.5), F32(0))).astype(F32)


def _scatter_rows(n_cols, idx, vals):
    out = np.zeros((idx.shape[0], n_cols), np.float64)
    rows = np.broadcast_to(np.arange(idx.shape[0])[:, None], idx.shape)
    np.add.at(out, (rows, idx), vals.astype(np.float64))
    return out


def sample_pdf_reformulation_return_u_bwd(bins, weights, tau, T, near, far, u, g_samples=None, g_T_below=None,
                                          g_tau_below=None, g_bin_below=None, zero_threshold=1e-4, epsilon_=1e-3):
    """What torch autograd computes through run_nerf_helpers.py:448-533 (the depth experiments back-propagate through the
    samples, depth_supervised_exps/run_nerf_sample_based_depth.py:881-932): cotangents [N,Ni] of (samples, T_below, tau_below,
    bin_below) -> (g_bins [N,S], g_near [N,1], g_far [N,1], g_tau [N,S+2], g_T [N,S+2]).  searchsorted has no gradient, so
    the weights get none; per sample the chain rule of pw_linear_sample_increasing / _decreasing (:340-361) with torch's
    rules at the kinks: max(eps, x) splits ties, clamp(t, eps, ds) sends the gradient to the bound it returns (ds when
    ds < eps or t > ds, nothing when t < eps), a NaN sample's gradient goes to s_left (:514)."""
    knots = np.concatenate([_f(near), _f(bins), _f(far)], -1)
    tau, T, u = _f(tau), _f(T), _f(u)
    nk = knots.shape[-1]
    z = lambda g: np.zeros_like(u) if g is None else _f(g)
    gx, gTb, gtb, gbb = z(g_samples), z(g_T_below), z(g_tau_below), z(g_bin_below)
    _, inds = sample_pdf_reformulation(bins, weights, tau, T, near, far, u, zero_threshold, epsilon_)
    below = np.maximum(0, inds - 1)
    above = np.minimum(nk - 1, inds)
    g = lambda a, i: np.take_along_axis(a, i, -1)
    s_l, s_r, T_l, tau_l, tau_r = g(knots, below), g(knots, above), g(T, below), g(tau, below), g(tau, above)
    dtau = g(tau[..., 1:] - tau[..., :-1], below)
    eps, zt = F32(epsilon_), F32(zero_threshold)
    const = (dtau < zt) & (dtau > -zt)
    inc = dtau >= zt
    sgn = np.where(inc, F32(1), F32(-1))
    with np.errstate(invalid="ignore", divide="ignore", over="ignore"):
        Tm = np.maximum(eps, T_l); c1 = (F32(1) - u) / Tm; m1 = np.maximum(eps, c1); L = -np.log(m1)
        dsr = s_r - s_l; dsm = np.maximum(eps, dsr)
        A = np.where(inc, tau_r - tau_l, tau_l - tau_r)
        D = tau_l * tau_l + sgn * (F32(2) * A * L) / dsm
        sq = np.sqrt(np.maximum(eps, D)); dt = np.maximum(eps, A)
        num = np.where(inc, -tau_l + sq, tau_l - sq)
        t0 = dsr * num / dt
        t = np.minimum(np.maximum(t0, eps), dsr)
        x = s_l + t
        live = ~const & ~np.isnan(x)
        to_bound = (dsr < eps) | (t0 > dsr)
        g_dsr = np.where(to_bound, gx, F32(0))
        g_t0 = np.where(~to_bound & (t0 >= eps), gx, F32(0))
        g_dsr = g_dsr + g_t0 * (num / dt)
        g_num = g_t0 * (dsr / dt)
        g_dt = -g_t0 * (t0 / dt)
        g_tl = np.where(inc, -g_num, g_num)
        g_sq = np.where(inc, g_num, -g_num)
        g_D = g_sq * (F32(0.5) / sq) * _max_grad(eps, D)
        g_tl = g_tl + F32(2) * tau_l * g_D
        g_A = sgn * (F32(2) * L / dsm) * g_D + _max_grad(eps, A) * g_dt
        g_L = sgn * (F32(2) * A / dsm) * g_D
        g_dsm = -sgn * (F32(2) * A * L) / (dsm * dsm) * g_D
        g_tr = np.where(inc, g_A, -g_A)
        g_tl = g_tl + np.where(inc, -g_A, g_A)
        g_dsr = g_dsr + _max_grad(eps, dsr) * g_dsm
        g_c1 = _max_grad(eps, c1) * (-g_L / m1)
        g_Tl = _max_grad(eps, T_l) * (-g_c1 * (c1 / Tm))
    zero = F32(0)
    g_sl = gx + np.where(live, -g_dsr, zero) + gbb
    g_sr = np.where(live, g_dsr, zero)
    g_Tl = np.where(live, g_Tl, zero) + gTb
    g_tl = np.where(live, g_tl, zero) + gtb
    g_tr = np.where(live, g_tr, zero)
    g_knots = _scatter_rows(nk, below, g_sl) + _scatter_rows(nk, above, g_sr)
    g_tau = _scatter_rows(nk, below, g_tl) + _scatter_rows(nk, above, g_tr)
    g_T = _scatter_rows(nk, below, g_Tl)
    return (g_knots[:, 1:-1].astype(F32), g_knots[:, :1].astype(F32), g_knots[:, -1:].astype(F32), g_tau.astype(F32),
            g_T.astype(F32))


def sample_pdf_return_u_bwd(bins, weights, u, g_samples):
    """Autograd through run_nerf_helpers.py:286-337: x = bins_b + t (bins_a - bins_b), t = (u - cdf_b) / denom (denom
    replaced by 1 below 1e-5, :331: no gradient through it then), cdf = [0, cumsum(pdf)], pdf = (w + 1e-5) / sum(w + 1e-5)
    -> (g_bins [N,nb], g_weights [N,nb-1])."""
    bins, u, gx = _f(bins), _f(u), _f(g_samples)
    wt = _f(weights) + F32(1e-5)
    W = np.sum(wt, -1, keepdims=True, dtype=F32)
    pdf = wt / W
    cdf = _cumsum32(pdf)
    cdf = np.concatenate([np.zeros_like(cdf[..., :1]), cdf], -1)
    nb = cdf.shape[-1]
    inds = _searchsorted_right(cdf, u)
    below = np.maximum(0, inds - 1)
    above = np.minimum(nb - 1, inds)
    g = lambda a, i: np.take_along_axis(a, i, -1)
    denom0 = g(cdf, above) - g(cdf, below)
    small = denom0 < F32(1e-5)
    denom = np.where(small, F32(1), denom0)
    t = (u - g(cdf, below)) / denom
    g_t = gx * (g(bins, above) - g(bins, below))
    g_den = np.where(small, F32(0), -g_t * t / denom)
    g_bins = _scatter_rows(nb, below, gx * (F32(1) - t)) + _scatter_rows(nb, above, gx * t)
    g_cdf = _scatter_rows(nb, below, -g_t / denom - g_den) + _scatter_rows(nb, above, g_den)
    g_pdf = np.cumsum(g_cdf[:, ::-1], -1)[:, ::-1][:, 1:]            # cdf[i] = sum_{j<i} pdf[j]
    g_w = (g_pdf - np.sum(g_pdf * pdf, -1, keepdims=True)) / W
    return g_bins.astype(F32), g_w.astype(F32)


# --------------------------------------------------------------------------------------------
# render_rays / render -- run_plnerf.py:95-175, 627-758
# --------------------------------------------------------------------------------------------
def stratified_z(near, far, N_samples, t_rand=None, lindisp=False):
    """run_plnerf.py:683-705.  t_rand None -> perturb == 0."""
    # torch.linspace(0,1,steps) in fp32 (ATen RangeFactories): step = fl32(1/(n-1)); first half
    # start + step*i, second half end - step*(n-1-i), each evaluated with ONE rounding (fused
    # multiply-add: verified bit-exact against torch 2.11 CPU; the CUDA kernel contracts the same
    # way).  The product step32*k is exact in float64, so float64 arithmetic + one cast is an fma.
    step = np.float64(F32(1.0) / F32(N_samples - 1)) if N_samples > 1 else np.float64(0)
    idx = np.arange(N_samples)
    half = N_samples // 2
    t_vals = np.where(idx < half, step * idx, 1.0 - step * (N_samples - 1 - idx)).astype(F32)
    if not lindisp:
        z = near * (F32(1) - t_vals) + far * t_vals
    else:
        z = F32(1) / (F32(1) / near * (F32(1) - t_vals) + F32(1) / far * t_vals)
    z = np.broadcast_to(z, (near.shape[0], N_samples)).astype(F32)
    if t_rand is not None:
        mids = F32(0.5) * (z[..., 1:] + z[..., :-1])
        upper = np.concatenate([mids, z[..., -1:]], -1)
        lower = np.concatenate([z[..., :1], mids], -1)
        z = lower + (upper - lower) * _f(t_rand)
    return z.astype(F32)


def render_rays(ray_batch, params_coarse, params_fine, N_samples, mode, color_mode, N_importance=0,
                t_rand=None, u=None, noise0=0.0, noise1=0.0, lindisp=False, white_bkgd=False,
                zero_tol=1e-4, epsilon=1e-3, farcolorfix=False, constant_init=False, retraw=False,
                multires=10, multires_views=4, net_kw=None):
    """run_plnerf.py:627-758 with the random draws (t_rand, u, noise) made explicit."""
    net_kw = net_kw or {}
    ray_batch = _f(ray_batch)
    rays_o, rays_d = ray_batch[:, 0:3], ray_batch[:, 3:6]
    viewdirs = ray_batch[:, -3:] if ray_batch.shape[-1] > 8 else None
    near, far = ray_batch[:, 6:7], ray_batch[:, 7:8]
    z_vals = stratified_z(near, far, N_samples, t_rand, lindisp)
    pts = rays_o[:, None, :] + rays_d[:, None, :] * z_vals[:, :, None]
    if constant_init:
        mode = "constant"
    q = lambda p, prm: run_network(p, viewdirs, prm, multires, multires_views, **net_kw)
    raw = q(pts, params_coarse)
    rgb_map, disp_map, acc_map, weights, depth_map, tau, T = raw2outputs(
        raw, z_vals, near, far, rays_d, mode, color_mode, noise0, white_bkgd, farcolorfix)
    ret = {}
    if N_importance > 0:
        ret.update(rgb0=rgb_map, disp0=disp_map, depth0=depth_map, acc0=acc_map)
        ret.update(z_vals0=z_vals, weights0=weights, raw0=raw)
        if mode == "linear":
            ret.update(tau0=tau, T0=T)
            z_samples, inds = sample_pdf_reformulation(z_vals, weights, tau, T, near, far, u,
                                                       zero_tol, epsilon)
        else:
            z_mid = F32(0.5) * (z_vals[..., 1:] + z_vals[..., :-1])
            z_samples, inds = sample_pdf(z_mid, weights[..., 1:-1], u)
        ret.update(inds=inds, z_samples_raw=z_samples)
        z_samples = np.minimum(np.maximum(z_samples, near), far)
        z_vals = np.sort(np.concatenate([z_vals, z_samples], -1), -1)
        pts = rays_o[:, None, :] + rays_d[:, None, :] * z_vals[:, :, None]
        raw = q(pts, params_fine if params_fine is not None else params_coarse)
        rgb_map, disp_map, acc_map, weights, depth_map, tau, T = raw2outputs(
            raw, z_vals, near, far, rays_d, mode, color_mode, noise1, white_bkgd, farcolorfix)
        ret["z_std"] = np.std(z_samples.astype(F32), -1).astype(F32)
        ret["z_vals"] = z_vals
    ret.update(rgb_map=rgb_map, disp_map=disp_map, acc_map=acc_map, depth_map=depth_map)
    if retraw:
        ret["raw"] = raw
    return ret


def ndc_rays(H, W, focal, near, rays_o, rays_d):
    """run_nerf_helpers.py:184-201."""
    rays_o, rays_d = _f(rays_o), _f(rays_d)
    t = -(F32(near) + rays_o[..., 2]) / rays_d[..., 2]
    rays_o = rays_o + t[..., None] * rays_d
    sx = F32(-1.0 / (W / (2.0 * focal)))
    sy = F32(-1.0 / (H / (2.0 * focal)))
    o0 = sx * rays_o[..., 0] / rays_o[..., 2]
    o1 = sy * rays_o[..., 1] / rays_o[..., 2]
    o2 = F32(1.0) + F32(2.0 * near) / rays_o[..., 2]
    d0 = sx * (rays_d[..., 0] / rays_d[..., 2] - rays_o[..., 0] / rays_o[..., 2])
    d1 = sy * (rays_d[..., 1] / rays_d[..., 2] - rays_o[..., 1] / rays_o[..., 2])
    d2 = F32(-2.0 * near) / rays_o[..., 2]
    return np.stack([o0, o1, o2], -1).astype(F32), np.stack([d0, d1, d2], -1).astype(F32)


def pack_rays(H, W, K, rays_o, rays_d, near, far, use_viewdirs, ndc):
    """The ray-batch packing done by render(), run_plnerf.py:140-164 -> [N, 8|11]."""
    rays_o, rays_d = _f(rays_o).reshape(-1, 3), _f(rays_d).reshape(-1, 3)
    viewdirs = None
    if use_viewdirs:
        viewdirs = rays_d / np.sqrt(np.sum(rays_d ** 2, -1, keepdims=True))
    if ndc:
        rays_o, rays_d = ndc_rays(H, W, K[0][0], 1.0, rays_o, rays_d)
    nearv = F32(near) * np.ones_like(rays_d[..., :1])
    farv = F32(far) * np.ones_like(rays_d[..., :1])
    cols = [rays_o, rays_d, nearv, farv] + ([viewdirs] if use_viewdirs else [])
    return np.concatenate(cols, -1).astype(F32)


def render(H, W, K, rays_o, rays_d, chunk=1024 * 32, ndc=True, near=0.0, far=1.0, use_viewdirs=False,
           t_rand=None, u=None, noise0=None, noise1=None, **kwargs):
    """run_plnerf.py:110-175 + batchify_rays :95-107 (chunked over rays, dict of concatenated
    outputs).  Per-ray random tensors are sliced per chunk."""
    rays = pack_rays(H, W, K, rays_o, rays_d, near, far, use_viewdirs, ndc)
    outs = {}
    sl = lambda a, i: None if a is None else a[i:i + chunk]
    for i in range(0, rays.shape[0], chunk):
        kw = dict(kwargs)
        if noise0 is not None:
            kw["noise0"] = noise0[i:i + chunk]
        if noise1 is not None:
            kw["noise1"] = noise1[i:i + chunk]
        r = render_rays(rays[i:i + chunk], t_rand=sl(t_rand, i), u=sl(u, i), **kw)
        for k, v in r.items():
            outs.setdefault(k, []).append(v)
    return {k: np.concatenate(v, 0) for k, v in outs.items()}
