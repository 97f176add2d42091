"""GPU tests of the training path (autograd through render_rays): compositing backward, the
input-gradient chain, the weight-gradient GEMMs and the heads, all through the C ABI.

Gradient GEMMs use bf16 tensor-core operands with fp32 accumulation (activations and output gradients
are rounded to bf16 once), so parameter gradients are compared tensor-wise: relative Frobenius error
<= 3e-2 and cosine similarity >= 0.999 against fp32 PyTorch autograd / the reference's own backward."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import plnerf_oracle as O
from util import CASES, case_params, load_golden, max_rel

pytestmark = pytest.mark.gpu


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def make_net(kw, params, requires_grad=True):
    from plnerf_b200.run_nerf_helpers import NeRF
    net = NeRF(D=kw["D"], W=kw["W"], input_ch=kw["input_ch"], input_ch_views=kw["input_ch_views"],
               output_ch=kw["output_ch"], skips=list(kw["skips"]), use_viewdirs=kw["use_viewdirs"])
    net.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in params.items()})
    net = net.cuda()
    for p in net.parameters():
        p.requires_grad_(requires_grad)
    return net


def torch_ref_query(net, rays, z):
    """fp32 PyTorch restatement of run_network + NeRF.forward (run_plnerf.py:78-92,
    run_nerf_helpers.py:105-128) on the module's own nn.Linear layers -- test-side checker only."""
    pts = rays[:, None, 0:3] + rays[:, None, 3:6] * z[:, :, None]
    def emb(x, L):
        outs = [x]
        for k in range(L):
            outs += [torch.sin(x * 2.0 ** k), torch.cos(x * 2.0 ** k)]
        return torch.cat(outs, -1)
    x = emb(pts.reshape(-1, 3), 10)
    vd = emb(rays[:, None, -3:].expand(pts.shape).reshape(-1, 3), 4)
    h = x
    for i, l in enumerate(net.pts_linears):
        h = F.relu(l(h))
        if i in net.skips:
            h = torch.cat([x, h], -1)
    alpha = net.alpha_linear(h)
    feat = net.feature_linear(h)
    hv = F.relu(net.views_linears[0](torch.cat([feat, vd], -1)))
    rgb = net.rgb_linear(hv)
    return torch.cat([rgb, alpha], -1).reshape(z.shape[0], z.shape[1], 4)


def bf16(t):
    return t.to(torch.bfloat16).to(torch.float32)


def emulated_backward(net, rays, z, g_raw):
    """The kernels' arithmetic restated in PyTorch (fp32 matmuls on bf16-ROUNDED operands): forward
    activations, output gradients and GEMM weights are rounded to bf16 exactly where the sm_100a path
    rounds them; biases, heads and the viewdir columns stay fp32.  Returns {param name: grad}."""
    D = net.D
    pts = rays[:, None, 0:3] + rays[:, None, 3:6] * z[:, :, None]
    def emb(x, L):
        outs = [x]
        for k in range(L):
            outs += [torch.sin(x * 2.0 ** k), torch.cos(x * 2.0 ** k)]
        return torch.cat(outs, -1)
    x = bf16(emb(pts.reshape(-1, 3), 10))
    vd = emb(rays[:, None, -3:].expand(pts.shape).reshape(-1, 3), 4)
    W = {k: v.detach() for k, v in net.named_parameters()}
    ic = x.shape[1]
    ins, masks = [], []
    h = x
    for i in range(D):
        w, b = W[f"pts_linears.{i}.weight"], W[f"pts_linears.{i}.bias"]
        ins.append(h)
        h32 = F.relu(h @ bf16(w).t() + b)
        masks.append(h32 > 0)
        a = bf16(h32)
        h = torch.cat([x, a], -1) if i in net.skips else a
        last = a
    feat = bf16(last @ bf16(W["feature_linear.weight"]).t() + W["feature_linear.bias"])
    wv, bv = W["views_linears.0.weight"], W["views_linears.0.bias"]
    hv32 = F.relu(feat @ bf16(wv[:, :256]).t() + (vd @ wv[:, 256:].t() + bv))
    mask_v = hv32 > 0
    hv = bf16(hv32)
    g = g_raw.reshape(-1, 4)
    g_rgb, g_alpha = g[:, :3], g[:, 3:4]
    grads = {}
    grads["rgb_linear.weight"] = g_rgb.t() @ hv
    grads["rgb_linear.bias"] = g_rgb.sum(0)
    grads["alpha_linear.weight"] = g_alpha.t() @ last
    grads["alpha_linear.bias"] = g_alpha.sum(0)
    d_hv = bf16((g_rgb @ W["rgb_linear.weight"]) * mask_v)
    grads["views_linears.0.weight"] = torch.cat([d_hv.t() @ feat, d_hv.t() @ bf16(vd)], -1)
    grads["views_linears.0.bias"] = d_hv.sum(0)
    d_feat = bf16(d_hv @ bf16(wv[:, :256]))
    grads["feature_linear.weight"] = d_feat.t() @ last
    grads["feature_linear.bias"] = d_feat.sum(0)
    dy = bf16((d_feat @ bf16(W["feature_linear.weight"]) + g_alpha * W["alpha_linear.weight"]) * masks[D - 1])
    for i in range(D - 1, -1, -1):
        grads[f"pts_linears.{i}.weight"] = dy.t() @ ins[i]
        grads[f"pts_linears.{i}.bias"] = dy.sum(0)
        if i > 0:
            w = W[f"pts_linears.{i}.weight"]
            wh = w[:, ic:] if (i - 1) in net.skips else w
            dy = bf16((dy @ bf16(wh)) * masks[i - 1])
    return grads


def cmp_grads(got, ref, tol=3e-2, what=""):
    for k in ref:
        a, b = got[k].double().flatten(), ref[k].double().flatten()
        nb = b.norm().item()
        if nb == 0:
            assert a.norm().item() == 0, (what, k)
            continue
        rel = (a - b).norm().item() / nb
        cos = torch.dot(a, b).item() / (a.norm().item() * nb + 1e-300)
        assert rel < tol and cos > 0.999, (what, k, rel, cos)


@pytest.mark.parametrize("n,S", [(40, 96), (3, 64), (257, 192)])
def test_network_query_backward_vs_torch_autograd(n, S):
    from plnerf_b200 import ops
    cfg, kw, pc, pf = case_params("lego_linear_mid")
    net = make_net(kw, pc)
    rs = np.random.RandomState(n)
    g = load_golden("lego_linear_mid")
    rb = np.tile(g["ray_batch"], (n // g["ray_batch"].shape[0] + 1, 1))[:n]
    rays = dev(rb)
    z = dev(np.sort(rs.rand(n, S).astype(np.float32) * 4 + 2, -1))      # seeded: the tolerances below are per-sample-set
    g_raw = dev(rs.randn(n, S, 4).astype(np.float32))
    with torch.no_grad():
        raw, stash = ops.network_query_train(net, rays, z)
        raw_nograd = ops.network_query(net, rays, z, precision="bf16")
        grads = ops.network_query_bwd(net, g_raw, stash, n, S)
    # the stash-mode forward (k_mlp_fwd<2>) and the inference forward (k_mlp3) are different kernels on the same bf16
    # operands with the same fp32 bias / ReLU / pack order: they differ by fp32 summation order in the small heads only
    # (alpha: 256 -> 1, rgb: 128 -> 3), i.e. by a few fp32 ulps of the output scale
    scale = raw_nograd.abs().amax(dim=(0, 1))
    assert ((raw - raw_nograd).abs() / scale).max().item() < 2e-6
    # (a) against the same arithmetic restated in PyTorch (bf16-rounded operands): tight
    emu = emulated_backward(net, rays, z, g_raw)
    cmp_grads(grads, emu, tol=2e-2, what=f"emulation n={n},S={S}")   # residual = isolated bf16 rounding flips
    # (b) against pure fp32 autograd: bf16 operand rounding accumulates over the 9 chained GEMMs of the
    # gradient chain (measured ~2% at the heads' side, ~10% at layer 0 for N(0,1) upstream gradients)
    ref_raw = torch_ref_query(net, rays, z)
    (ref_raw * g_raw).sum().backward()
    ref = {k: p.grad for k, p in net.named_parameters()}
    for k in ref:
        a, b = grads[k].double().flatten(), ref[k].double().flatten()
        cos = torch.dot(a, b).item() / (a.norm().item() * b.norm().item() + 1e-300)
        assert cos > 0.985 and abs(a.norm().item() / b.norm().item() - 1) < 0.08, (k, cos, a.norm().item() / b.norm().item())


@pytest.mark.parametrize("name", ["lego_linear_mid", "lego_constant", "llff_ndc_linear", "llff_ndc_constant"])
def test_training_loss_gradients_vs_reference(name):
    """loss = mse(rgb, tgt) + mse(rgb0, tgt) through render() with the reference's pytest draws:
    loss value and every parameter gradient (norm + first 256 entries) vs the unmodified reference."""
    from plnerf_b200 import run_plnerf as RP
    g = load_golden(name)
    cfg, kw, pc, pf = case_params(name)
    net_c, net_f = make_net(kw, pc), make_net(kw, pf)
    Hh, Ww, focal = g["hwf"]
    rays = torch.stack([dev(g["rays_o"]), dev(g["rays_d"])])
    rgb, disp, acc, extras = RP.render(int(Hh), int(Ww), g["K"], chunk=1024 * 32, rays=rays, ndc=cfg["ndc"],
                                       near=cfg["near"], far=cfg["far"], use_viewdirs=True, network_query_fn=None,
                                       network_fn=net_c, network_fine=net_f, N_samples=cfg["Ns"],
                                       N_importance=cfg["Ni"], perturb=1.0, raw_noise_std=cfg["raw_noise_std"],
                                       white_bkgd=cfg["white_bkgd"], mode=cfg["mode"], color_mode=cfg["color_mode"],
                                       lindisp=cfg["lindisp"], pytest=True, retraw=True,
                                       constant_init=cfg["constant_init"])
    assert rgb.requires_grad and extras["rgb0"].requires_grad and not extras["raw"].requires_grad
    tgt = dev(g["train_target"])
    loss = torch.mean((rgb - tgt) ** 2) + torch.mean((extras["rgb0"] - tgt) ** 2)
    assert abs(loss.item() - float(g["train_loss"])) / float(g["train_loss"]) < 1e-2     # bf16 forward
    loss.backward()
    for tag, net in (("c", net_c), ("f", net_f)):
        for k, p in net.named_parameters():
            # bf16 operand rounding accumulates along the 9 chained gradient GEMMs: the heads / views /
            # feature layers agree with the fp32 reference to a few %, the deepest trunk layers to ~10-15%
            # (the tight check of the kernels is the bf16-emulating restatement above; here the bf16 FORWARD also
            # moves the importance samples slightly, i.e. the fine network is evaluated at slightly different points)
            deep = k.startswith("pts_linears")
            ref_norm = float(g[f"gnorm_{tag}.{k}"])
            got = p.grad.flatten()
            assert abs(got.norm().item() - ref_norm) / ref_norm < (0.10 if deep else 0.05), (tag, k, got.norm().item(), ref_norm)
            head = dev(g[f"ghead_{tag}.{k}"]).double()
            a = got[:head.numel()].double()
            rel = (a - head).norm().item() / (head.norm().item() + 1e-300)
            if not deep:   # 256-entry slices of the 1e-6-sized trunk gradients (24 rays) are too noisy to gate on
                assert rel < 0.06, (tag, k, rel)


def test_gradients_need_bf16_and_reach_networks_without_viewdirs():
    """precision='bf16x3' has no backward (raises); a network without view directions (coarse only, output_ch = 4) does:
    loss.backward() through render_rays fills output_linear and the trunk, leaves the unused views_linears untouched
    (the parity of these gradients against the reference: tests/test_gpu_train_parity.py)."""
    from plnerf_b200 import run_plnerf as RP
    cfg, kw, pc, pf = case_params("coarse_only")
    net = make_net(kw, pc)
    g = load_golden("coarse_only")
    with pytest.raises(NotImplementedError):
        RP.render_rays(dev(g["ray_batch"]), net, None, 64, "linear", "midpoint", perturb=1.0, precision="bf16x3")
    assert not net.use_viewdirs
    ret = RP.render_rays(dev(g["ray_batch"]), net, None, 64, "linear", "midpoint", perturb=1.0, white_bkgd=True)
    assert ret["rgb_map"].requires_grad
    ((ret["rgb_map"] - 0.25) ** 2).mean().backward()
    for k, p in net.named_parameters():
        if k.startswith("views_linears"):
            assert p.grad is None or not bool(p.grad.any())
        else:
            assert p.grad is not None and bool(torch.isfinite(p.grad).all()) and float(p.grad.abs().max()) > 0, k
