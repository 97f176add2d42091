"""Golden vectors for the mesh extractor's density grid, from the UNMODIFIED reference.

    python tests/golden/make_golden_fields.py          (build container only)

Imports /root/reference/nerf_extract_mesh.py (stub modules for its uninstalled imports: the five of
oracle/refimport.py plus trimesh, mcubes, load_deepvoxels, load_LINEMOD -- none is touched by
``extract_fields``) and runs the reference's own ``extract_fields`` (nerf_extract_mesh.py:531-562)
through its own ``run_network`` (:80-95), ``get_embedder`` and ``NeRF`` on CPU.  Stored: the bounds,
the three coordinate vectors, the density grid u and max|sigma| before the relu.  Two cases are
committed (20^3, with and without view directions); a 66^3 grid (crosses the reference's 64-wide block boundary) is compared
with the oracle here but not stored.
"""
import importlib
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import refimport  # noqa: E402
import plnerf_oracle as O  # noqa: E402

synth = importlib.import_module("pl-nerf_b200.synth")

CASES = {
    "fields_viewdirs": dict(res=20, use_viewdirs=True, seed=51, bmin=(-1.2, -1.0, -0.8), bmax=(1.2, 1.1, 0.9)),
    "fields_noviews": dict(res=20, use_viewdirs=False, seed=52, bmin=(-1.0, -1.0, -1.0), bmax=(1.0, 1.0, 1.0)),
}


def load_mesh_module():
    H, _ = refimport.load()
    for n in ["trimesh", "mcubes", "load_deepvoxels", "load_LINEMOD"]:
        if n not in sys.modules:
            sys.modules[n] = types.ModuleType(n)
    sys.modules["load_deepvoxels"].load_dv_data = None
    sys.modules["load_LINEMOD"].load_LINEMOD_data = None
    import nerf_extract_mesh as M  # noqa: E402  (the reference's module)
    M.tqdm = lambda x, *a, **k: x
    return H, M


def net_kwargs(use_viewdirs):
    return dict(D=8, W=256, input_ch=63, input_ch_views=27 if use_viewdirs else 0, output_ch=5, skips=(4,),
                use_viewdirs=use_viewdirs)


def reference_grid(H, M, params, kw, bmin, bmax, res):
    net = H.NeRF(D=kw["D"], W=kw["W"], input_ch=kw["input_ch"], input_ch_views=kw["input_ch_views"],
                 output_ch=kw["output_ch"], skips=list(kw["skips"]), use_viewdirs=kw["use_viewdirs"])
    net.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in params.items()})
    embed_fn, _ = H.get_embedder(10, 0)
    embeddirs_fn = H.get_embedder(4, 0)[0] if kw["use_viewdirs"] else None
    q = lambda p, v, fn: M.run_network(p, v if kw["use_viewdirs"] else None, fn, embed_fn=embed_fn,
                                       embeddirs_fn=embeddirs_fn, netchunk=1024 * 64)
    u = M.extract_fields(torch.tensor(bmin), torch.tensor(bmax), res, q, net)
    # scale of the un-rectified density channel over the same grid (the tolerance's denominator in the tests)
    X, Y, Z = (torch.from_numpy(a) for a in axes_of(bmin, bmax, res))
    xx, yy, zz = torch.meshgrid(X, Y, Z, indexing="ij")
    pts = torch.stack([xx.reshape(-1), yy.reshape(-1), zz.reshape(-1)], -1)
    with torch.no_grad():
        sigma = q(pts, torch.zeros_like(pts), net)[..., 3]
    dev = float(np.abs(np.maximum(sigma.numpy(), 0).reshape(u.shape) - u).max())
    assert dev <= 1e-6 * float(sigma.abs().max()), dev      # same values up to sgemm blocking (64^3 sub-cubes vs one batch)
    return u, float(sigma.abs().max())


def axes_of(bmin, bmax, res):
    return [torch.linspace(float(np.float32(bmin[k])), float(np.float32(bmax[k])), res).numpy() for k in range(3)]


def oracle_kw(kw):
    return dict(D=kw["D"], skips=kw["skips"], input_ch=kw["input_ch"], input_ch_views=kw["input_ch_views"],
                use_viewdirs=kw["use_viewdirs"])


def main():
    H, M = load_mesh_module()
    for name, c in CASES.items():
        kw = net_kwargs(c["use_viewdirs"])
        params = synth.nerf_params(c["seed"], **kw)
        u, sigma_abs_max = reference_grid(H, M, params, kw, c["bmin"], c["bmax"], c["res"])
        axes = axes_of(c["bmin"], c["bmax"], c["res"])
        uo = O.extract_fields(axes, params, **oracle_kw(kw))
        err = float(np.abs(uo - u).max() / max(np.abs(u).max(), 1e-6))
        print(f"{name}: u {u.shape} max {u.max():.3f} nonzero {np.mean(u > 0):.2f}  oracle rel err {err:.2e}")
        assert err < 1e-5
        np.savez_compressed(os.path.join(HERE, name + ".npz"), u=u, X=axes[0], Y=axes[1], Z=axes[2],
                            bound_min=np.array(c["bmin"], np.float32), bound_max=np.array(c["bmax"], np.float32),
                            resolution=np.int64(c["res"]), sigma_abs_max=np.float32(sigma_abs_max), seed=np.int64(c["seed"]),
                            use_viewdirs=np.int64(c["use_viewdirs"]))
    # block-boundary cross-check (not stored): 66 > the reference's N = 64 split
    kw = net_kwargs(True)
    params = synth.nerf_params(53, **kw)
    bmin, bmax = (-1.0, -1.0, -1.0), (1.0, 1.0, 1.0)
    u, _ = reference_grid(H, M, params, kw, bmin, bmax, 66)
    uo = O.extract_fields(axes_of(bmin, bmax, 66), params, **oracle_kw(kw))
    err = float(np.abs(uo - u).max() / max(np.abs(u).max(), 1e-6))
    print(f"block boundary 66^3: oracle rel err {err:.2e}")
    assert err < 1e-5


if __name__ == "__main__":
    main()
