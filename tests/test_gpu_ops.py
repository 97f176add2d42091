"""GPU parity tests (run with -m gpu on a B200): every CUDA operator is called through the C ABI
(plnerf_b200.ops -> ctypes -> libplnerf_b200.so) and compared with the numpy oracle and with the
golden vectors of the unmodified reference.

Tolerances (stated per test): integer outputs (searchsorted indices) and pure data movement
(sort-merge) are bit-exact; fp32 kernels <= 1e-5..1e-4 relative as noted.
"""
import numpy as np
import pytest
import torch

import plnerf_oracle as O
from util import CASES, case_params, load_golden, max_rel, oracle_net_kw, synth

pytestmark = pytest.mark.gpu

ALL = list(CASES)
FINE = [c for c in ALL if CASES[c]["Ni"] > 0]


@pytest.fixture(scope="module")
def P():
    import plnerf_b200
    from plnerf_b200 import ops
    assert torch.cuda.is_available()
    return ops


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def host(t):
    return t.detach().cpu().numpy()


# ------------------------------------------------------------------------------------------------
def test_umma_debug_gemm(P):
    """Pins the tcgen05 descriptor encodings: D = A B^T with bf16-rounded operands, fp32 accumulate,
    A from shared memory (SS) and from tensor memory (TS).  The single-tile GEMM lives in the developer library
    (libplnerf_b200_debug.so, -DPLNERF_DEBUG): the product library exports no debug entry points."""
    import ctypes as C
    from plnerf_b200 import _lib as L
    rs = np.random.RandomState(0)
    for a_mode in (0, 1):
        for N, K in ((128, 16), (128, 64), (256, 256), (128, 256)):
            A = rs.randn(128, K).astype(np.float32)
            B = rs.randn(N, K).astype(np.float32)
            ref = O.bf16_round(A).astype(np.float64) @ O.bf16_round(B).astype(np.float64).T
            dA, dB = dev(A), dev(B)
            D = torch.zeros((128, N), device="cuda")
            L.check(L.debug_lib().plnerf_debug_umma_gemm_ex(dA.data_ptr(), dB.data_ptr(), N, K, a_mode, 2048, 128,
                                                            D.data_ptr(), None))
            torch.cuda.synchronize()
            err = np.abs(host(D) - ref).max() / np.abs(ref).max()
            assert err < 1e-5, (a_mode, N, K, err)


@pytest.mark.parametrize("name", ALL)
def test_encode(P, name):
    g = load_golden(name)
    out = host(P.encode(dev(g["pts0"][0]), 10))
    assert max_rel(out, g["embed_pts0"], 1.0) < 2e-6       # |sin|,|cos| <= 1: absolute 2e-6
    if "embed_dirs" in g:
        out = host(P.encode(dev(g["ray_batch"][:, -3:]), 4))
        assert max_rel(out, g["embed_dirs"], 1.0) < 1e-6


@pytest.mark.parametrize("name", ALL)
def test_stratified_z_bit_exact(P, name):
    g = load_golden(name)
    cfg = CASES[name]
    z = host(P.stratified_z(dev(g["ray_batch"]), cfg["Ns"], cfg["lindisp"], True, dev(g["t_rand"])))
    np.testing.assert_array_equal(z, g["z_vals0"])


@pytest.mark.parametrize("name", ALL)
def test_raw2outputs_vs_golden(P, name):
    """Quadrature on the reference's own raw: weights/tau/T and the composited maps, <= 2e-5 rel."""
    g = load_golden(name)
    cfg = CASES[name]
    mode = "constant" if cfg["constant_init"] else cfg["mode"]
    noise = dev(g["noise0"]) if "noise0" in g else None
    rgb, disp, acc, w, depth, tau, T = P.raw2outputs(dev(g["raw0"]), dev(g["z_vals0"]), dev(g["ray_batch"]), mode,
                                                     cfg["color_mode"], noise=noise, white_bkgd=cfg["white_bkgd"])
    sfx = "0" if cfg["Ni"] > 0 else "_map"
    assert max_rel(host(w), g["weights0"], 1e-2) < 2e-5
    assert max_rel(host(rgb), g["rgb0" if cfg["Ni"] > 0 else "rgb_map"]) < 2e-5
    assert max_rel(host(depth), g["depth" + sfx]) < 2e-5
    assert max_rel(host(acc), g["acc" + sfx]) < 2e-5
    assert max_rel(host(disp), g["disp" + sfx]) < 2e-5
    if mode == "linear":
        np.testing.assert_array_equal(host(tau), g["tau0"])
        assert max_rel(host(T), g["T0"], 1e-6) < 2e-5


@pytest.mark.parametrize("name", ALL)
def test_raw2outputs_backward_vs_reference_autograd(P, name):
    """d(loss)/d(raw) through raw2outputs vs torch autograd on the reference's own function (golden):
    <= 1e-4 of the gradient scale (fp32; (1-e) cancellations and exp make this looser than forward)."""
    g = load_golden(name)
    cfg = CASES[name]
    mode = "constant" if cfg["constant_init"] else cfg["mode"]
    noise = dev(g["noise0"]) if "noise0" in g else None
    graw = host(P.raw2outputs_bwd(dev(g["raw0"]), dev(g["z_vals0"]), dev(g["ray_batch"]), mode, cfg["color_mode"],
                                  g_rgb=dev(g["up_rgb"]), g_depth=dev(g["up_depth"]), g_acc=dev(g["up_acc"]),
                                  g_disp=dev(g["up_disp"]), noise=noise, white_bkgd=cfg["white_bkgd"]))
    ref = g["g_raw0"]
    for c in range(4):
        scale = np.abs(ref[..., c]).max() + 1e-12
        err = np.abs(graw[..., c] - ref[..., c]).max() / scale
        assert err < 1e-4, (c, err)
    assert np.all(graw[..., 4:] == 0)


@pytest.mark.parametrize("name", FINE)
def test_sampler_indices_bit_exact(P, name):
    """Given the reference's own (z, weights, tau, T, u): searchsorted indices identical, samples 1e-5."""
    g = load_golden(name)
    cfg = CASES[name]
    mode = "constant" if cfg["constant_init"] else cfg["mode"]
    rays = dev(g["ray_batch"])
    if mode == "linear":
        zs, inds = P.sample_pdf_pl(dev(g["z_vals0"]), dev(g["weights0"]), dev(g["tau0"]), dev(g["T0"]), rays,
                                   cfg["Ni"], u=dev(g["u"]), return_inds=True)
    else:
        z = g["z_vals0"]
        zs, inds = P.sample_pdf(dev(0.5 * (z[..., 1:] + z[..., :-1])), dev(g["weights0"][..., 1:-1]), cfg["Ni"],
                                u=dev(g["u"]), return_inds=True)
    np.testing.assert_array_equal(host(inds), g["inds"])
    assert max_rel(host(zs), g["z_samples_raw"], 1e-2) < 1e-5


@pytest.mark.parametrize("name", FINE)
def test_merge_bit_exact(P, name):
    g = load_golden(name)
    rb = g["ray_batch"]
    zm, zstd = P.merge_samples(dev(g["z_vals0"]), dev(g["z_samples_raw"]), dev(rb))
    np.testing.assert_array_equal(host(zm), g["z_vals"])
    assert max_rel(host(zstd), g["z_std"], 1e-3) < 1e-4


def test_sampler_edge_cases(P):
    """Ragged / degenerate rows: all-zero weights, a single spike, u at 0 and just below 1,
    equal depths, N_importance not a multiple of 32."""
    rs = np.random.RandomState(3)
    n, S, Ni = 7, 64, 45
    z = np.sort(rs.uniform(2, 6, (n, S)).astype(np.float32), -1)
    z[1] = z[1, 0]                               # all depths equal
    near = np.full((n, 1), 2.0, np.float32); far = np.full((n, 1), 6.0, np.float32)
    raw = rs.randn(n, S, 4).astype(np.float32) * 3
    raw[2, :, 3] = -5.0                          # empty space: tau = 0 everywhere
    raw[3, :, 3] = 0.0; raw[3, 20, 3] = 500.0    # one spike
    rays = np.zeros((n, 8), np.float32); rays[:, 3:6] = rs.randn(n, 3); rays[:, 6:7] = near; rays[:, 7:8] = far
    rgb, disp, acc, w, depth, tau, T = O.raw2outputs(raw, z, near, far, rays[:, 3:6], "linear", "midpoint")
    u = rs.uniform(0, 1, (n, Ni)).astype(np.float32)
    u[:, 0] = 0.0; u[:, 1] = np.float32(1.0) - np.float32(2 ** -24)
    zs_o, inds_o = O.sample_pdf_reformulation(z, w, tau, T, near, far, u)
    zs, inds = P.sample_pdf_pl(dev(z), dev(w), dev(tau), dev(T), dev(rays), Ni, u=dev(u), return_inds=True)
    np.testing.assert_array_equal(host(inds), inds_o)
    assert max_rel(host(zs), zs_o, 1e-2) < 1e-5
    zm, _ = P.merge_samples(dev(z), dev(zs_o), dev(rays))
    np.testing.assert_array_equal(host(zm), np.sort(np.concatenate([z, np.clip(zs_o, near, far)], -1), -1))


@pytest.mark.parametrize("Ni", [1, 5, 31, 32, 33, 64, 100, 128, 200, 256, 300])
@pytest.mark.parametrize("with_nan", [False, True])
def test_merge_sort_paths_vs_torch_sort(P, Ni, with_nan):
    """clamp + sort(cat) (run_plnerf.py:728-734) over every sorting path of merge_ray: keys in registers (1, 2, 4 or 8 per
    lane for N_importance <= 32, 64, 128, 256; pads when it is not a power of two), the shared-memory network for
    N_importance > 256 and for rows containing NaN -- bit for bit against torch.sort, with repeated values, samples outside
    [near, far] (clamped) and samples equal to coarse depths."""
    g = torch.Generator(device="cuda")
    g.manual_seed(100 + Ni)
    n, S = 37, 64
    z = torch.sort(torch.rand(n, S, device="cuda", generator=g) * 4 + 2, -1)[0]
    x = torch.rand(n, Ni, device="cuda", generator=g) * 5 + 1.5          # some outside [2, 6]
    if Ni >= 5:
        x[:, 3] = x[:, 0]                                                  # repeats
        x[:, 4] = z[:, 10]                                                 # ties with a coarse depth
    if with_nan:
        x[::3, Ni // 2] = float("nan")
    rays = torch.zeros(n, 8, device="cuda")
    rays[:, 3:6] = 1.0
    rays[:, 6], rays[:, 7] = 2.0, 6.0
    zm, zstd = P.merge_samples(z, x, rays)
    ref = torch.sort(torch.cat([z, torch.clamp(x, 2.0, 6.0)], -1), -1)[0]
    assert torch.equal(zm.view(torch.int32), ref.view(torch.int32))
    clean = ~torch.isnan(x).any(-1)
    std_ref = torch.std(torch.clamp(x, 2.0, 6.0), -1, unbiased=False)
    assert torch.allclose(zstd[clean], std_ref[clean], rtol=1e-4, atol=1e-6)


def test_empty_batches(P):
    rays = torch.zeros((0, 11), device="cuda")
    assert P.stratified_z(rays, 64).shape == (0, 64)
    assert P.encode(torch.zeros((0, 3), device="cuda"), 10).shape == (0, 63)


def test_cpu_tensor_rejected(P):
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        P.encode(torch.zeros((4, 3)), 10)


def test_tensor_on_another_device_rejected(P):
    """The wrappers enqueue on the current device's current stream: a tensor of another GPU is refused, not mis-launched."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    with pytest.raises(RuntimeError, match="current CUDA device"):
        P.encode(torch.zeros((4, 3), device="cuda:1"), 10)
    with torch.cuda.device(1):
        assert P.encode(torch.zeros((4, 3), device="cuda:1"), 10).shape == (4, 63)


# ------------------------------------------------------------------------------------------------
def make_net(kw, params):
    from plnerf_b200.run_nerf_helpers import NeRF
    net = NeRF(D=kw["D"], W=kw["W"], input_ch=kw["input_ch"], input_ch_views=kw["input_ch_views"],
               output_ch=kw["output_ch"], skips=list(kw["skips"]), use_viewdirs=kw["use_viewdirs"])
    net.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in params.items()})
    return net.cuda()


@pytest.mark.parametrize("name", ["lego_linear_mid", "lego_left_noise_lindisp"])
@pytest.mark.parametrize("precision,tol", [("bf16x3", 1e-4), ("bf16", 2e-2)])
def test_mlp_forward_vs_reference(P, name, precision, tol):
    """NeRF.forward on embedded rows vs the reference's raw (golden): relative to the output scale.
    bf16x3 must meet the 1e-4 north-star tolerance; plain bf16 is only sanity-bounded here (its
    gate is the bf16-emulating oracle below)."""
    g = load_golden(name)
    cfg, kw, pc, pf = case_params(name)
    net = make_net(kw, pc)
    pts = g["pts0"].reshape(-1, 3)
    emb = O.embed(pts, 10)
    if cfg["use_viewdirs"]:
        vd = np.broadcast_to(g["ray_batch"][:, None, -3:], g["pts0"].shape).reshape(-1, 3)
        emb = np.concatenate([emb, O.embed(vd, 4)], -1)
    with torch.no_grad():
        out = host(P.mlp_forward(net, dev(emb), precision=precision))
        P.set_precision(precision)          # NeRF.forward itself uses the process-wide default
        out2 = host(net(dev(emb)))
        P.set_precision("bf16")
    np.testing.assert_array_equal(out, out2)
    ref = g["raw0"].reshape(out.shape[0], -1)[:, :out.shape[1]]
    scale = np.abs(ref).max(0, keepdims=True)
    err = np.abs(out - ref) / scale
    assert err.max() < tol, err.max()


@pytest.mark.parametrize("name", ["lego_linear_mid", "lego_left_noise_lindisp"])
def test_mlp_forward_bf16_vs_emulation(P, name):
    """Fast mode vs an oracle that rounds the same operands to bf16.  Agreement is ~1e-7 of the output
    scale wherever both sides round every activation to the same bf16 value; fp32 summation-order
    differences flip an occasional bf16 rounding (1 ulp = 2^-8 relative on that activation), which
    shows up as a sparse tail (~1-2% of outputs) -- so the gate is on quantiles: 90% within 1e-6,
    99% within 5e-4, max 5e-3."""
    g = load_golden(name)
    cfg, kw, pc, pf = case_params(name)
    net = make_net(kw, pc)
    pts = g["pts0"].reshape(-1, 3)
    emb = O.embed(pts, 10)
    if cfg["use_viewdirs"]:
        vd = np.broadcast_to(g["ray_batch"][:, None, -3:], g["pts0"].shape).reshape(-1, 3)
        emb = np.concatenate([emb, O.embed(vd, 4)], -1)
    with torch.no_grad():
        out = host(P.mlp_forward(net, dev(emb), precision="bf16"))
    ref = O.nerf_forward(pc, emb, emulate_bf16=True, **oracle_net_kw(kw))[:, :out.shape[1]]
    scale = np.abs(ref).max(0, keepdims=True)
    err = np.abs(out - ref) / scale
    qs = np.quantile(err, [0.5, 0.9, 0.99, 0.999, 1.0])
    assert qs[1] < 1e-6 and qs[2] < 5e-4 and qs[4] < 5e-3, qs


@pytest.mark.parametrize("D,viewdirs", [(11, True), (11, False), (3, True), (2, False)])
def test_mlp_other_depths(P, D, viewdirs):
    """Depths other than 8: D=11 with view directions (50 weight stages per tile) does not fit k_mlp3's stage program and
    runs on the single-tile kernel k_mlp_fwd (the fallback must stay correct); the others run on k_mlp3 with a different
    layer program.  (D >= 12 exceeds the kernels' shared-memory constant block and is refused with PLNERF_E_UNSUPPORTED.)  Same bf16-emulating oracle
    and quantile gate as above, through the fused-PE query (194 rays x 33 samples: ragged tiles)."""
    from plnerf_b200.run_nerf_helpers import NeRF
    kw = dict(D=D, W=256, input_ch=63, input_ch_views=27 if viewdirs else 0, output_ch=4, skips=(1,) if D > 2 else (),
              use_viewdirs=viewdirs)
    prm = synth.nerf_params(5, **kw)
    net = NeRF(D=D, W=256, input_ch=63, input_ch_views=kw["input_ch_views"], output_ch=4, skips=list(kw["skips"]),
               use_viewdirs=viewdirs)
    net.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in prm.items()})
    net = net.cuda()
    n, S = 194, 33
    ro, rd, K, _ = synth.lego_rays(n, seed=2)
    vd = rd / np.linalg.norm(rd, axis=-1, keepdims=True)
    rays = np.concatenate([ro, rd, np.full((n, 1), 2, np.float32), np.full((n, 1), 6, np.float32), vd], -1).astype(np.float32)
    z = np.sort(np.random.RandomState(1).rand(n, S).astype(np.float32) * 4 + 2, -1)
    pts = (ro[:, None] + rd[:, None] * z[..., None]).astype(np.float32).reshape(-1, 3)
    emb = O.embed(pts, 10)
    if viewdirs:
        emb = np.concatenate([emb, O.embed(np.broadcast_to(vd[:, None], (n, S, 3)).reshape(-1, 3).astype(np.float32), 4)], -1)
    ref = O.nerf_forward(prm, emb, emulate_bf16=True, D=D, skips=kw["skips"], input_ch=63,
                         input_ch_views=kw["input_ch_views"], use_viewdirs=viewdirs)
    with torch.no_grad():
        raw = host(P.network_query(net, dev(rays if viewdirs else rays[:, :8]), dev(z), precision="bf16"))
    ref = ref[:, :raw.shape[-1]]
    scale = np.abs(ref).max(0, keepdims=True)
    qs = np.quantile(np.abs(raw.reshape(ref.shape) - ref) / scale, [0.5, 0.9, 0.99, 0.999, 1.0])
    assert qs[1] < 2e-6 and qs[2] < 1e-3 and qs[4] < 1e-2, qs


@pytest.mark.parametrize("name", ALL)
def test_network_query_fused_pe(P, name):
    """Fused query (PE computed in-kernel from rays and depths) vs the reference's raw0, bf16x3."""
    g = load_golden(name)
    cfg, kw, pc, pf = case_params(name)
    net = make_net(kw, pc)
    with torch.no_grad():
        raw = host(P.network_query(net, dev(g["ray_batch"]), dev(g["z_vals0"]), precision="bf16x3"))
    ref = g["raw0"][..., :raw.shape[-1]]
    scale = np.abs(ref).reshape(-1, ref.shape[-1]).max(0)
    err = np.abs(raw - ref) / scale
    assert err.max() < 1e-4, err.max()


def test_mlp_ragged_rows(P):
    """Row counts that are not multiples of the 128-row tile, including a single row."""
    cfg, kw, pc, pf = case_params("lego_linear_mid")
    net = make_net(kw, pc)
    rs = np.random.RandomState(1)
    for m in (1, 127, 129, 300):
        x = rs.uniform(-1, 1, (m, 90)).astype(np.float32)
        with torch.no_grad():
            out = host(P.mlp_forward(net, dev(x), precision="bf16x3"))
        ref = O.nerf_forward(pc, x, **oracle_net_kw(kw))
        assert np.abs(out - ref).max() / np.abs(ref).max() < 1e-4, m


# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ALL)
def test_render_end_to_end_vs_reference(P, name):
    """render() through the public API with the reference's pytest draws vs the reference's outputs:
    rgb/depth/acc/disp within 1e-4 relative (north-star tolerance), bf16x3 precision."""
    from plnerf_b200 import run_plnerf as RP
    g = load_golden(name)
    cfg, kw, pc, pf = case_params(name)
    net_c = make_net(kw, pc)
    net_f = make_net(kw, pf) if cfg["Ni"] > 0 else None
    Hh, Ww, focal = g["hwf"]
    rays = torch.stack([dev(g["rays_o"]), dev(g["rays_d"])])
    with torch.no_grad():
        rgb, disp, acc, extras = RP.render(int(Hh), int(Ww), g["K"], chunk=1024 * 32, rays=rays, ndc=cfg["ndc"],
                                           near=cfg["near"], far=cfg["far"], use_viewdirs=cfg["use_viewdirs"],
                                           network_query_fn=None, network_fn=net_c, network_fine=net_f,
                                           N_samples=cfg["Ns"], N_importance=cfg["Ni"], perturb=1.0,
                                           raw_noise_std=cfg["raw_noise_std"], white_bkgd=cfg["white_bkgd"],
                                           mode=cfg["mode"], color_mode=cfg["color_mode"], lindisp=cfg["lindisp"],
                                           pytest=True, retraw=True, constant_init=cfg["constant_init"],
                                           precision="bf16x3")
    tol = 1e-4
    assert max_rel(host(rgb), g["rgb_map"]) < tol
    assert max_rel(host(acc), g["acc_map"]) < tol
    assert max_rel(host(disp), g["disp_map"]) < tol
    assert max_rel(host(extras["depth_map"]), g["depth_map"]) < tol
    if cfg["Ni"] > 0:
        for k in ("rgb0", "depth0", "acc0", "disp0"):
            assert max_rel(host(extras[k]), g[k]) < tol, k
        assert max_rel(host(extras["z_std"]), g["z_std"], 1e-3) < 1e-3
    assert extras["raw"].shape == g["raw"].shape[:2] + (extras["raw"].shape[-1],)


def test_render_bf16_fast_mode_close(P):
    """Fast bf16 mode end to end vs the reference: documented looser bound (bf16 operands)."""
    from plnerf_b200 import run_plnerf as RP
    name = "lego_linear_mid"
    g = load_golden(name)
    cfg, kw, pc, pf = case_params(name)
    net_c, net_f = make_net(kw, pc), make_net(kw, pf)
    rays = torch.stack([dev(g["rays_o"]), dev(g["rays_d"])])
    Hh, Ww, focal = g["hwf"]
    with torch.no_grad():
        rgb, disp, acc, extras = RP.render(int(Hh), int(Ww), g["K"], rays=rays, ndc=False, near=2., far=6.,
                                           use_viewdirs=True, network_query_fn=None, network_fn=net_c,
                                           network_fine=net_f, N_samples=64, N_importance=128, perturb=1.0,
                                           white_bkgd=True, mode="linear", color_mode="midpoint", pytest=True,
                                           precision="bf16")
    assert max_rel(host(rgb), g["rgb_map"]) < 2e-2
    assert max_rel(host(extras["depth_map"]), g["depth_map"]) < 5e-2


def test_philox_draws_invariant_to_chunking(P):
    """Device-side draws are keyed by global ray id: chunked and unchunked renders are identical."""
    from plnerf_b200 import run_plnerf as RP
    cfg, kw, pc, pf = case_params("lego_linear_mid")
    net_c, net_f = make_net(kw, pc), make_net(kw, pf)
    g = load_golden("lego_linear_mid")
    rays = dev(g["ray_batch"])
    common = dict(network_fn=net_c, network_query_fn=None, network_fine=net_f, N_samples=64, N_importance=128,
                  mode="linear", color_mode="midpoint", perturb=1.0, raw_noise_std=1.0, white_bkgd=True, seed=1234)
    with torch.no_grad():
        a = RP.batchify_rays(rays, chunk=1024, **common)
        b = RP.batchify_rays(rays, chunk=7, **common)
    for k in a:
        assert torch.equal(a[k], b[k]), k
    zs = host(P.stratified_z(rays, 64, seed=9))
    z0 = O.stratified_z(g["ray_batch"][:, 6:7], g["ray_batch"][:, 7:8], 64, None)
    assert np.all(zs >= 2.0) and np.all(zs <= 6.0) and np.all(np.diff(zs, axis=-1) >= 0)
    assert np.abs(zs - z0).max() < (6 - 2) / 63


# ------------------------------------------------------------------------------------------------
# f-1: ray generation + packing (render(), run_plnerf.py:138-164)
# ------------------------------------------------------------------------------------------------
def _torch_pack(H, W, K, c2w=None, rays=None, ndc=True, near=0., far=1., use_viewdirs=False, c2w_staticcam=None):
    """render()'s packing restated with the reference's own torch ops on the CPU."""
    from plnerf_b200.run_nerf_helpers import get_rays, ndc_rays
    if c2w is not None:
        rays_o, rays_d = get_rays(H, W, K, c2w)
    else:
        rays_o, rays_d = rays
    viewdirs = None
    if use_viewdirs:
        viewdirs = rays_d
        if c2w_staticcam is not None:
            rays_o, rays_d = get_rays(H, W, K, c2w_staticcam)
        viewdirs = viewdirs / torch.norm(viewdirs, dim=-1, keepdim=True)
        viewdirs = torch.reshape(viewdirs, [-1, 3]).float()
    if ndc:
        rays_o, rays_d = ndc_rays(H, W, K[0][0], 1., rays_o, rays_d)
    rays_o, rays_d = torch.reshape(rays_o, [-1, 3]).float(), torch.reshape(rays_d, [-1, 3]).float()
    nr, fr = near * torch.ones_like(rays_d[..., :1]), far * torch.ones_like(rays_d[..., :1])
    out = torch.cat([rays_o, rays_d, nr, fr], -1)
    return torch.cat([out, viewdirs], -1) if use_viewdirs else out


@pytest.mark.parametrize("case", ["lego_pose", "llff_pose_ndc", "given_rays", "given_rays_ndc_noview", "staticcam"])
def test_pack_rays_vs_torch(P, case):
    """plnerf_pack_rays against the reference's get_rays / viewdir normalisation / ndc_rays / cat sequence (torch, CPU):
    every fp32 operation is rounded separately in the same order, so the packed rows agree to the last bit or two."""
    rs = np.random.RandomState(5)
    if case.startswith("llff") or "ndc" in case:
        H, W, focal = 30, 40, 32.5
    else:
        H, W, focal = 25, 33, 41.25
    K = np.array([[focal, 0, 0.5 * W], [0, focal, 0.5 * H], [0, 0, 1]])
    pose = lambda th: torch.from_numpy(synth.pose_spherical(th, -30.0, 4.0)[:3, :4].astype(np.float32).copy())
    kw = dict(ndc=False, near=2.0, far=6.0, use_viewdirs=True)
    c2w = rays = static = None
    if case == "lego_pose":
        c2w = pose(40.0)
    elif case == "llff_pose_ndc":
        c2w = torch.from_numpy(np.concatenate([np.eye(3, dtype=np.float32) + 0.05 * rs.randn(3, 3).astype(np.float32),
                                               0.1 * rs.randn(3, 1).astype(np.float32)], 1))
        kw = dict(ndc=True, near=0.0, far=1.0, use_viewdirs=True)
    elif case.startswith("given_rays"):
        o = rs.randn(77, 3).astype(np.float32)
        d = rs.randn(77, 3).astype(np.float32); d[:, 2] = -np.abs(d[:, 2]) - 0.5
        rays = (torch.from_numpy(o), torch.from_numpy(d))
        if case == "given_rays_ndc_noview":
            kw = dict(ndc=True, near=0.0, far=1.0, use_viewdirs=False)
    else:
        c2w, static = pose(40.0), pose(-75.0)
    ref = _torch_pack(H, W, K, c2w=c2w, rays=rays, c2w_staticcam=static, **kw).numpy()
    got, sh = P.pack_rays(H, W, K, c2w=None if c2w is None else c2w.cuda(), rays=None if rays is None else (rays[0].cuda(), rays[1].cuda()),
                          c2w_staticcam=None if static is None else static.cuda(), **kw)
    got = host(got)
    assert got.shape == ref.shape and tuple(sh)[-1] == 3
    np.testing.assert_allclose(got, ref, rtol=3e-7, atol=1e-7)
    assert (got == ref).mean() > 0.95          # almost everything is bit-identical (torch.norm's reduction order is the exception)


# ------------------------------------------------------------------------------------------------
# Full-size, size-independent properties (BASELINE.json configs[1]: one 32 768-ray chunk at 64 + 128 samples)
# ------------------------------------------------------------------------------------------------
def test_full_size_properties(P):
    """At the bench shape the oracle is too slow to be the checker, so the path is held to properties that do not depend
    on size: merged depths sorted, inside [near, far] and a superset of the coarse depths; outputs in range; the
    composited weights of `retraw` raw re-derived by the op-level quadrature give back the same maps; bitwise
    determinism; invariance to the chunk split (Philox draws are keyed by global ray id); equivariance to a permutation
    of the rays (no row depends on its tile neighbours)."""
    n, Ns, Ni = 32768, 64, 128
    kw = dict(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=(4,), use_viewdirs=True)
    net_c, net_f = make_net(kw, synth.nerf_params(1, **kw)), make_net(kw, synth.nerf_params(2, **kw))
    ro, rd, K, _ = synth.lego_rays(n, seed=7)
    vd = rd / np.linalg.norm(rd, axis=-1, keepdims=True)
    rays = dev(np.concatenate([ro, rd, np.full((n, 1), 2, np.float32), np.full((n, 1), 6, np.float32), vd], -1).astype(np.float32))
    common = dict(N_samples=Ns, N_importance=Ni, mode="linear", color_mode="midpoint", perturb=True, white_bkgd=True,
                  seed=77, precision="bf16")
    with torch.no_grad():
        a = P.render_rays_fwd(rays, net_c, net_f, retraw=True, want_z=True, **common)
        b = P.render_rays_fwd(rays, net_c, net_f, retraw=True, want_z=True, **common)
        # two half chunks with the global ray offset
        h1 = P.render_rays_fwd(rays[: n // 2], net_c, net_f, ray_id_offset=0, **common)
        h2 = P.render_rays_fwd(rays[n // 2:], net_c, net_f, ray_id_offset=n // 2, **common)
        # permuted rays with explicit (permuted) draws
        g = torch.Generator(device="cuda"); g.manual_seed(3)
        t_rand = torch.rand(n, Ns, device="cuda", generator=g)
        u = torch.rand(n, Ni, device="cuda", generator=g)
        perm = torch.randperm(n, device="cuda", generator=g)
        kw2 = {k: v for k, v in common.items() if k != "seed"}
        c = P.render_rays_fwd(rays, net_c, net_f, t_rand=t_rand, u=u, **kw2)
        d = P.render_rays_fwd(rays[perm].contiguous(), net_c, net_f, t_rand=t_rand[perm].contiguous(), u=u[perm].contiguous(), **kw2)
        # the op-level quadrature on the fine raw gives the same maps
        rgb2, disp2, acc2, w2, depth2, _, _ = P.raw2outputs(a["raw"], a["z_vals"], rays, "linear", "midpoint", white_bkgd=True)
    z = a["z_vals"]
    assert bool((z[:, 1:] >= z[:, :-1]).all()) and float(z.min()) >= 2.0 and float(z.max()) <= 6.0
    for k in ("rgb_map", "acc_map", "disp_map", "depth_map", "rgb0", "acc0", "z_std"):
        assert bool(torch.isfinite(a[k]).all()), k
    assert float(a["acc_map"].min()) >= 0.0 and float(a["acc_map"].max()) <= 1.0 + 1e-5
    assert float(a["rgb_map"].min()) >= -1e-6 and float(a["rgb_map"].max()) <= 1.0 + 1e-5
    assert float(a["depth_map"].min()) >= 0.0 and float(a["depth_map"].max()) <= 6.0 * (1 + 1e-5)
    assert float(a["z_std"].min()) >= 0.0
    for k in a:                                   # determinism, bit for bit
        assert torch.equal(a[k], b[k]), k
    for k in h1:                                  # chunk-split invariance, bit for bit
        assert torch.equal(torch.cat([h1[k], h2[k]], 0), a[k]), k
    for k in c:                                   # permutation equivariance, bit for bit
        assert torch.equal(c[k][perm], d[k]), k
    assert torch.equal(rgb2, a["rgb_map"]) and torch.equal(acc2, a["acc_map"]) and torch.equal(depth2, a["depth_map"])
    assert float((w2.sum(-1) - a["acc_map"]).abs().max()) < 1e-5      # checksum: the weights add up to the opacity


@pytest.mark.parametrize("Ns,Ni,mode,two_nets", [(64, 128, "linear", True), (64, 128, "constant", True), (32, 64, "linear", False),
                                                 (64, 256, "linear", True), (96, 300, "constant", True)])
def test_fused_quadrature_paths_agree(P, Ns, Ni, mode, two_nets):
    """render_rays runs the quadrature inside k_mlp3 when a ray's samples fit the output ring (S <= 256) and as the separate
    kernel otherwise; `raw` reaches HBM only with retraw.  With and without retraw the maps are bit-identical, and the
    op-level quadrature on the returned raw reproduces them bit for bit -- for both routes, both modes, one or two networks,
    ray counts that leave partial tiles and partial CTA ranges."""
    n = 1000
    kw = dict(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=(4,), use_viewdirs=True)
    net_c = make_net(kw, synth.nerf_params(1, **kw))
    net_f = make_net(kw, synth.nerf_params(2, **kw)) if two_nets else None
    ro, rd, K, _ = synth.lego_rays(n, seed=11)
    vd = rd / np.linalg.norm(rd, axis=-1, keepdims=True)
    rays = dev(np.concatenate([ro, rd, np.full((n, 1), 2, np.float32), np.full((n, 1), 6, np.float32), vd], -1).astype(np.float32))
    common = dict(N_samples=Ns, N_importance=Ni, mode=mode, color_mode="midpoint", perturb=True, white_bkgd=True, seed=5,
                  raw_noise_std=0.5, precision="bf16")
    with torch.no_grad():
        a = P.render_rays_fwd(rays, net_c, net_f, retraw=True, want_z=True, **common)
        b = P.render_rays_fwd(rays, net_c, net_f, retraw=False, **common)
    for k in b:
        assert torch.equal(a[k], b[k]), k
    assert bool(torch.isfinite(a["rgb_map"]).all()) and bool(torch.isfinite(a["depth_map"]).all())
    z = a["z_vals"]
    assert z.shape == (n, Ns + Ni) and bool((z[:, 1:] >= z[:, :-1]).all())


def test_packed_weights_follow_the_parameters():
    """The packed bf16 copy is a cache of the module's parameters: an in-place update (optimizer.step: version counters
    move), load_state_dict, an alias update + invalidate_packed, and a storage move (.data re-homed into another buffer)
    are all seen by the next query; an untouched module is not repacked."""
    import plnerf_b200
    from plnerf_b200 import ops
    cfg, kw, pc, pf = case_params("lego_linear_mid")
    net = make_net(kw, pc)
    g = load_golden("lego_linear_mid")
    rays, z = dev(g["ray_batch"]), dev(g["z_vals0"])
    with torch.no_grad():
        base = ops.network_query(net, rays, z).clone()
        n0 = ops.launch_count()
        again = ops.network_query(net, rays, z)
        assert torch.equal(again, base) and ops.launch_count() - n0 <= 2          # no pack kernels: view bias + MLP only
        ref_sd = {k: v.clone() for k, v in net.state_dict().items()}
        net.alpha_linear.bias.add_(0.25)                                           # in place: the version counter moves
        moved = ops.network_query(net, rays, z)
        assert (moved[..., 3] - base[..., 3]).abs().max().item() > 0.2
        net.load_state_dict(ref_sd)
        assert torch.equal(ops.network_query(net, rays, z), base)
        flat = torch.cat([p.data.flatten() for p in net.parameters()])             # re-home the storage (what TrainStep does)
        off = 0
        for p in net.parameters():
            p.data = flat[off:off + p.numel()].view_as(p)
            off += p.numel()
        assert torch.equal(ops.network_query(net, rays, z), base)
        flat[-net.rgb_linear.bias.numel() - net.rgb_linear.weight.numel():] *= 0.5   # alias update: no counter moves ...
        assert torch.equal(ops.network_query(net, rays, z), base)                  # ... so the cache is (documentedly) stale
        ops.invalidate_packed(net)
        changed = ops.network_query(net, rays, z)
        assert (changed[..., :3] - base[..., :3]).abs().max().item() > 1e-3 and torch.equal(changed[..., 3], base[..., 3])
