"""plnerf_b200: B200-native (sm_100a) implementation of PL-NeRF's ray-rendering hot path.

The directory is named ``pl-nerf_b200`` (not importable by the ``import`` statement); use
``import plnerf_b200`` (the shim module at the repository root) or
``importlib.import_module("pl-nerf_b200")``.

Public surface (mirrors the reference's two hot-path modules):
  plnerf_b200.run_plnerf        render, batchify_rays, render_rays, raw2outputs, run_network, ...
  plnerf_b200.run_nerf_helpers  NeRF, get_embedder, sample_pdf, sample_pdf_reformulation, ...
  plnerf_b200.nerf_extract_mesh extract_fields (density grid of the mesh extractor), extract_iso_level
  plnerf_b200.ops               torch-tensor wrappers over the C ABI (include/plnerf_b200.h)
  plnerf_b200.dist              ray sharding + gradient all-reduce helpers (one process per GPU)
  plnerf_b200.synth             synthetic lego/LLFF-shaped rays and seeded NeRF parameters
"""
from . import _lib  # noqa: F401
from . import synth  # noqa: F401


def build(force=False, verbose=False, debug=False):
    """Compile the CUDA library in-tree (nvcc, sm_100a); debug=True: the developer library (see _lib.py)."""
    return _lib.build(force=force, verbose=verbose, debug=debug)


def __getattr__(name):
    # torch-dependent submodules are imported lazily so that `synth` / `build` work without torch
    if name in ("ops", "run_plnerf", "run_nerf_helpers", "dist", "autograd", "nerf_extract_mesh"):
        import importlib
        return importlib.import_module(f"{__name__}.{name}")
    if name in ("render", "render_rays", "batchify_rays", "raw2outputs", "install"):
        import importlib
        return getattr(importlib.import_module(f"{__name__}.run_plnerf"), name)
    if name in ("NeRF", "get_embedder", "sample_pdf", "sample_pdf_reformulation"):
        import importlib
        return getattr(importlib.import_module(f"{__name__}.run_nerf_helpers"), name)
    raise AttributeError(name)
