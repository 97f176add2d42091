"""Density grid of the mesh extractor (nerf_extract_mesh.py:531-562).

CPU: the oracle's ``extract_fields`` against grids produced by the unmodified reference
(tests/golden/make_golden_fields.py) and the exactness of the column packing the CUDA path uses.
GPU (-m gpu): ``plnerf_b200.nerf_extract_mesh.extract_fields`` through the C ABI against the same
golden grids (bf16x3, <= 1e-4 of the density channel's scale) and against the oracle on a 66^3 grid that
crosses the reference's 64-wide block boundary and this path's launch-slab boundary.
"""
import numpy as np
import pytest
import torch

import plnerf_oracle as O
from make_golden_fields import CASES as FIELD_CASES, net_kwargs, oracle_kw
from util import load_golden, synth

NAMES = list(FIELD_CASES)


def _case(name):
    c = FIELD_CASES[name]
    kw = net_kwargs(c["use_viewdirs"])
    return c, kw, synth.nerf_params(c["seed"], **kw)


@pytest.mark.parametrize("name", NAMES)
def test_oracle_extract_fields_vs_reference(name):
    g = load_golden(name)
    c, kw, params = _case(name)
    u = O.extract_fields((g["X"], g["Y"], g["Z"]), params, **oracle_kw(kw))
    assert u.shape == g["u"].shape and u.dtype == np.float32
    assert np.abs(u - g["u"]).max() <= 1e-5 * float(g["sigma_abs_max"])
    # block walk: a 7-wide split of the same grid gives the same values (sub-cubes are independent; the small blocks
    # take the oracle's numpy branches, the full grid its torch branches: libm vs Sleef sin/cos, sgemm blocking)
    u7 = O.extract_fields((g["X"], g["Y"], g["Z"]), params, block=7, **oracle_kw(kw))
    assert np.abs(u7 - u).max() <= 2e-6 * float(g["sigma_abs_max"])


def test_axes_match_reference_linspace():
    """The coordinate vectors are the reference's CPU torch.linspace values bit for bit."""
    from plnerf_b200.nerf_extract_mesh import _axis
    for name in NAMES:
        g = load_golden(name)
        for k, key in enumerate("XYZ"):
            a = _axis(torch.tensor(g["bound_min"])[k], torch.tensor(g["bound_max"])[k], int(g["resolution"]))
            np.testing.assert_array_equal(a.numpy(), g[key])
            b = _axis(float(g["bound_min"][k]), float(g["bound_max"][k]), int(g["resolution"]))
            np.testing.assert_array_equal(b.numpy(), g[key])


@pytest.mark.parametrize("stride", [8, 11])
def test_grid_columns_are_exact(stride):
    """o + d*z of the packed columns (two fp32 roundings, as the kernel computes it) is the 'ij' meshgrid of the
    coordinate slices bit for bit, including negative zero and denormal-free extremes; viewdir slots are zero."""
    from plnerf_b200.nerf_extract_mesh import grid_columns
    rs = np.random.RandomState(3)
    X = torch.from_numpy(np.concatenate([rs.uniform(-3, 3, 5), [0.0, -0.0, 1e-30, -2.5e4]]).astype(np.float32))
    Y = torch.from_numpy(rs.uniform(-2, 2, 4).astype(np.float32))
    Z = torch.from_numpy(np.concatenate([rs.uniform(-1, 1, 6), [0.0, 7e5]]).astype(np.float32))
    cols, depths = grid_columns(X, Y, Z, stride)
    assert cols.shape == (X.numel() * Y.numel(), stride) and depths.shape == (cols.shape[0], Z.numel())
    o, d = cols[:, None, 0:3], cols[:, None, 3:6]
    pts = (o + (d * depths[:, :, None])).reshape(X.numel(), Y.numel(), Z.numel(), 3).numpy()
    xx, yy, zz = np.meshgrid(X.numpy(), Y.numpy(), Z.numpy(), indexing="ij")
    # values equal (0.0 == -0.0 is fine: sin/cos/identity of +-0 feed |.|-symmetric or sign-carrying terms that
    # the reference computes from the same value up to the sign of zero)
    np.testing.assert_array_equal(pts[..., 0], xx)
    np.testing.assert_array_equal(pts[..., 1], yy)
    np.testing.assert_array_equal(pts[..., 2], zz)
    assert float(cols[:, 6:].abs().max()) == 0.0


def test_extract_fields_needs_cuda_model():
    """No CPU path: a CPU-resident model raises instead of computing anywhere else."""
    from plnerf_b200.nerf_extract_mesh import extract_fields
    from plnerf_b200.run_nerf_helpers import NeRF
    net = NeRF(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=[4], use_viewdirs=True)
    with pytest.raises(RuntimeError, match="CUDA"):
        extract_fields([-1., -1., -1.], [1., 1., 1.], 4, None, net)


def test_iso_level_matches_reference_formula(capsys):
    from plnerf_b200.nerf_extract_mesh import extract_iso_level
    u = load_golden("fields_viewdirs")["u"]
    iso = extract_iso_level(u, threshold=0.5)
    assert iso == min(max(0.5, u.min() + u.std()), u.max() - u.std())
    capsys.readouterr()


# ------------------------------------------------------------------------------------------------ GPU
def _make_net(kw, params):
    from plnerf_b200.run_nerf_helpers import NeRF
    net = NeRF(D=kw["D"], W=kw["W"], input_ch=kw["input_ch"], input_ch_views=kw["input_ch_views"],
               output_ch=kw["output_ch"], skips=list(kw["skips"]), use_viewdirs=kw["use_viewdirs"])
    net.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in params.items()})
    return net.cuda()


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_extract_fields_vs_reference(name):
    """extract_fields through the public API vs the reference's grid, relative to the scale of the density channel
    before the relu (max|sigma| over the grid, stored with the golden -- the same per-channel output scale the MLP
    tests use): bf16x3 <= 1e-4 (north-star tolerance); plain bf16 sanity-bounded at 2e-2."""
    from plnerf_b200.nerf_extract_mesh import extract_fields
    g = load_golden(name)
    c, kw, params = _case(name)
    net = _make_net(kw, params)
    scale = float(g["sigma_abs_max"])
    for precision, tol in (("bf16x3", 1e-4), ("bf16", 2e-2)):
        u = extract_fields(torch.tensor(g["bound_min"]), torch.tensor(g["bound_max"]), int(g["resolution"]), None, net,
                           precision=precision)
        assert isinstance(u, np.ndarray) and u.shape == g["u"].shape and u.dtype == np.float32
        assert (u >= 0).all()
        err = np.abs(u - g["u"]).max() / scale
        assert err < tol, (precision, err)


@pytest.mark.gpu
def test_extract_fields_slabs_and_block_boundary(monkeypatch):
    """66^3 grid (crosses the reference's 64-wide sub-cube split) vs the oracle, once in a single launch and once
    forced into 5-plane slabs: the slab split must not change a bit."""
    from plnerf_b200 import nerf_extract_mesh as NM
    kw = net_kwargs(True)
    params = synth.nerf_params(53, **kw)
    net = _make_net(kw, params)
    R = 66
    bmin, bmax = [-1.0, -1.0, -1.0], [1.0, 1.0, 1.0]
    u1 = NM.extract_fields(bmin, bmax, R, None, net, precision="bf16x3")
    monkeypatch.setattr(NM, "_ROWS_PER_LAUNCH", 5 * R * R)
    u2 = NM.extract_fields(bmin, bmax, R, None, net, precision="bf16x3")
    np.testing.assert_array_equal(u1, u2)
    axes = [NM._axis(bmin[k], bmax[k], R).numpy() for k in range(3)]
    ref = O.extract_fields(axes, params, **oracle_kw(kw))
    err = np.abs(u1 - ref).max() / np.abs(ref).max()
    assert err < 1e-4, err


def test_install_rebinds_extract_fields_and_restores():
    """install() on a module that carries the mesh extractor's names (nerf_extract_mesh.py has its own copy of the
    render path plus extract_fields) rebinds them all; uninstall() restores the originals."""
    import types
    from plnerf_b200 import nerf_extract_mesh as NM, run_plnerf as RP
    fake = types.ModuleType("fake_nerf_extract_mesh")
    originals = {}
    for nme in RP._PATCHED + RP._PATCHED_HELPERS + ("extract_fields",):
        originals[nme] = object()
        setattr(fake, nme, originals[nme])
    saved = RP.install(fake)
    assert fake.extract_fields is NM.extract_fields and fake.render_rays is RP.render_rays
    assert saved["extract_fields"] is originals["extract_fields"]
    RP.uninstall(fake, saved)
    for nme, obj in originals.items():
        assert getattr(fake, nme) is obj, nme
    plain = types.ModuleType("fake_run_plnerf")          # run_plnerf.py itself has no extract_fields
    assert "extract_fields" not in RP.install(plain) and not hasattr(plain, "extract_fields")
