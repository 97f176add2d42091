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
