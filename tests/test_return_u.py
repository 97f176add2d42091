"""SURVEY.md 8 f-4, forward variants: sample_pdf_reformulation_return_u / sample_pdf_return_u (run_nerf_helpers.py:448-533,
286-337) against goldens of the unmodified reference functions (tests/golden/make_golden_return_u.py).  CPU: the oracle.
GPU: the kernels through the C ABI and the helper-module mirrors -- samples <= 1e-5 relative, the gathered T / tau / knot
and u bit-exact (pure gathers of the inputs)."""
import numpy as np
import pytest
import torch

import plnerf_oracle as O
from util import load_golden, max_rel

R = None


def G():
    global R
    if R is None:
        R = load_golden("return_u")
    return R


def pl_inputs(name):
    g = load_golden(name)
    rb = g["ray_batch"]
    return g, g["z_vals0"], g["weights0"], g["tau0"], g["T0"], rb[:, 6:7], rb[:, 7:8]


def pytest_u(n, Ni):
    np.random.seed(0)
    return np.random.rand(n, Ni).astype(np.float32)


@pytest.mark.parametrize("name", ["lego_linear_mid", "llff_ndc_linear"])
@pytest.mark.parametrize("tag", ["load", "pytest"])
def test_oracle_pl_return_u(name, tag):
    g, z, w, tau, T, near, far = pl_inputs(name)
    u = g["u"] if tag == "load" else pytest_u(*g["u"].shape)
    s, Tb, taub, binb, uu = O.sample_pdf_reformulation_return_u(z, w, tau, T, near, far, u)
    ref = {k: G()[f"{name}.pl.{tag}.{k}"] for k in ("samples", "T_below", "tau_below", "bin_below", "u")}
    assert max_rel(s, ref["samples"]) < 1e-5
    for got, k in ((Tb, "T_below"), (taub, "tau_below"), (binb, "bin_below"), (uu, "u")):
        np.testing.assert_array_equal(got, ref[k])


@pytest.mark.parametrize("tag", ["load", "pytest"])
def test_oracle_const_return_u(tag):
    g = load_golden("llff_ndc_constant")
    z, w = g["z_vals0"], g["weights0"]
    z_mid = np.float32(0.5) * (z[..., 1:] + z[..., :-1])
    u = g["u"] if tag == "load" else pytest_u(*g["u"].shape)
    s, uu = O.sample_pdf_return_u(z_mid, w[..., 1:-1], u)
    assert max_rel(s, G()[f"llff_ndc_constant.const.{tag}.samples"]) < 1e-5
    np.testing.assert_array_equal(uu, G()[f"llff_ndc_constant.const.{tag}.u"])


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["lego_linear_mid", "llff_ndc_linear"])
@pytest.mark.parametrize("tag", ["load", "pytest"])
def test_gpu_pl_return_u(name, tag):
    from plnerf_b200 import run_nerf_helpers as HP
    g, z, w, tau, T, near, far = pl_inputs(name)
    Ni = g["u"].shape[1]
    kw = dict(load_u=dev(g["u"])) if tag == "load" else dict(pytest=True)
    r = HP.sample_pdf_reformulation_return_u(dev(z), dev(w), dev(tau), dev(T), dev(near), dev(far), Ni, **kw)
    r = [t.cpu().numpy() for t in r]
    ref = {k: G()[f"{name}.pl.{tag}.{k}"] for k in ("samples", "T_below", "tau_below", "bin_below", "u")}
    assert max_rel(r[0], ref["samples"]) < 1e-5
    for got, k in zip(r[1:], ("T_below", "tau_below", "bin_below", "u")):
        np.testing.assert_array_equal(got, ref[k])


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["load", "pytest"])
def test_gpu_const_return_u(tag):
    from plnerf_b200 import run_nerf_helpers as HP
    g = load_golden("llff_ndc_constant")
    z, w = g["z_vals0"], g["weights0"]
    z_mid = np.float32(0.5) * (z[..., 1:] + z[..., :-1])
    kw = dict(load_u=dev(g["u"])) if tag == "load" else dict(pytest=True)
    s, u = HP.sample_pdf_return_u(dev(z_mid), dev(np.ascontiguousarray(w[..., 1:-1])), g["u"].shape[1], **kw)
    assert max_rel(s.cpu().numpy(), G()[f"llff_ndc_constant.const.{tag}.samples"]) < 1e-5
    np.testing.assert_array_equal(u.cpu().numpy(), G()[f"llff_ndc_constant.const.{tag}.u"])


@pytest.mark.gpu
def test_gpu_return_u_device_draws():
    """load_u=None without the pytest hook: u is drawn on the device (Philox) and returned; feeding the returned u back
    through load_u reproduces samples and gathers bit for bit (what the depth experiments rely on)."""
    from plnerf_b200 import run_nerf_helpers as HP
    g, z, w, tau, T, near, far = pl_inputs("lego_linear_mid")
    a = HP.sample_pdf_reformulation_return_u(dev(z), dev(w), dev(tau), dev(T), dev(near), dev(far), 128)
    u = a[4]
    assert float(u.min()) >= 0.0 and float(u.max()) < 1.0 and 0.45 < float(u.mean()) < 0.55
    b = HP.sample_pdf_reformulation_return_u(dev(z), dev(w), dev(tau), dev(T), dev(near), dev(far), 128, load_u=u)
    for x, y in zip(a, b):
        assert torch.equal(x, y)


# ---- f-4, differentiable form: the gradients torch autograd computes through the unmodified reference functions
# (tests/golden/make_golden_return_u_grad.py).  Gate: every gradient array within 1e-4 of its own largest entry (the scatter
# sums run in a different order than autograd's index_add).
GG = None


def grad_golden():
    global GG
    if GG is None:
        GG = load_golden("return_u_grad")
    return GG


def scaled_err(got, want):
    want = np.asarray(want, np.float64)
    return float(np.abs(np.asarray(got, np.float64).reshape(want.shape) - want).max() / max(np.abs(want).max(), 1e-6))


@pytest.mark.parametrize("name", ["lego_linear_mid", "llff_ndc_linear"])
@pytest.mark.parametrize("which", ["grad", "grad_samples_only"])
def test_oracle_pl_return_u_gradients(name, which):
    g, z, w, tau, T, near, far = pl_inputs(name)
    R = grad_golden()
    cot = [R[f"{name}.pl.cot.{k}"] for k in ("samples", "T_below", "tau_below", "bin_below")]
    if which == "grad_samples_only":
        cot = [cot[0], None, None, None]
    got = O.sample_pdf_reformulation_return_u_bwd(z, w, tau, T, near, far, g["u"], *cot)
    for arr, key in zip(got, ("z", "near", "far", "tau", "T")):
        assert scaled_err(arr, R[f"{name}.pl.{which}.{key}"]) < 1e-4, key


def test_oracle_const_return_u_gradients():
    g = load_golden("llff_ndc_constant")
    z, w = g["z_vals0"], g["weights0"]
    z_mid = np.float32(0.5) * (z[..., 1:] + z[..., :-1])
    R = grad_golden()
    g_bins, g_w = O.sample_pdf_return_u_bwd(z_mid, w[..., 1:-1], g["u"], R["llff_ndc_constant.const.cot.samples"])
    assert scaled_err(g_bins, R["llff_ndc_constant.const.grad.bins"]) < 1e-4
    assert scaled_err(g_w, R["llff_ndc_constant.const.grad.weights"]) < 1e-4


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["lego_linear_mid", "llff_ndc_linear"])
@pytest.mark.parametrize("which", ["grad", "grad_samples_only"])
def test_gpu_pl_return_u_gradients(name, which):
    """loss.backward() through the helper mirror (autograd.Function over plnerf_sample_pdf_pl_return_u_bwd) against the
    reference's autograd; the weights get no gradient; a second backward gives the same bits (no atomics)."""
    from plnerf_b200 import run_nerf_helpers as HP
    g, z, w, tau, T, near, far = pl_inputs(name)
    R = grad_golden()
    leaves = [dev(a).requires_grad_(True) for a in (z, w, tau, T, near, far)]
    runs = []
    for _ in range(2):
        for t in leaves:
            t.grad = None
        outs = HP.sample_pdf_reformulation_return_u(*leaves, g["u"].shape[1], load_u=dev(g["u"]))
        keys = ("samples", "T_below", "tau_below", "bin_below") if which == "grad" else ("samples",)
        loss = sum((o * dev(R[f"{name}.pl.cot.{k}"])).sum() for o, k in zip(outs, keys))
        loss.backward()
        runs.append([None if t.grad is None else t.grad.clone() for t in leaves])
    zg, wg, taug, Tg, ng, fg = runs[0]
    assert wg is None or not bool(wg.any())
    for arr, key in ((zg, "z"), (ng, "near"), (fg, "far"), (taug, "tau"), (Tg, "T")):
        assert scaled_err(arr.cpu().numpy(), R[f"{name}.pl.{which}.{key}"]) < 1e-4, key
    for a, b in zip(runs[0], runs[1]):
        assert (a is None and b is None) or torch.equal(a, b)
    # and against the oracle restatement on the same inputs
    want = O.sample_pdf_reformulation_return_u_bwd(z, w, tau, T, near, far, g["u"], *(
        [R[f"{name}.pl.cot.{k}"] for k in ("samples", "T_below", "tau_below", "bin_below")] if which == "grad"
        else [R[f"{name}.pl.cot.samples"], None, None, None]))
    for arr, ref in zip((zg, ng, fg, taug, Tg), want):
        assert scaled_err(arr.cpu().numpy(), ref) < 1e-4


@pytest.mark.gpu
def test_gpu_const_return_u_gradients():
    from plnerf_b200 import run_nerf_helpers as HP
    g = load_golden("llff_ndc_constant")
    z, w = g["z_vals0"], g["weights0"]
    z_mid = np.float32(0.5) * (z[..., 1:] + z[..., :-1])
    R = grad_golden()
    bins = dev(z_mid).requires_grad_(True)
    wt = dev(np.ascontiguousarray(w[..., 1:-1])).requires_grad_(True)
    s, u = HP.sample_pdf_return_u(bins, wt, g["u"].shape[1], load_u=dev(g["u"]))
    assert not u.requires_grad
    (s * dev(R["llff_ndc_constant.const.cot.samples"])).sum().backward()
    assert scaled_err(bins.grad.cpu().numpy(), R["llff_ndc_constant.const.grad.bins"]) < 1e-4
    assert scaled_err(wt.grad.cpu().numpy(), R["llff_ndc_constant.const.grad.weights"]) < 1e-4
    # no gradient requested: the plain forward path, same samples
    s2, _ = HP.sample_pdf_return_u(bins.detach(), wt.detach(), g["u"].shape[1], load_u=dev(g["u"]))
    assert torch.equal(s2, s.detach())
