"""Import the UNMODIFIED reference (run_nerf_helpers, run_plnerf) from /root/reference.

Test infrastructure only (golden generation + optional oracle cross-checks in the build
container).  The reference's top-level imports need five packages that are not installed here;
they are stubbed with empty modules -- none of them is touched by the hot
path.  Nothing on the GPU box may call this: /root/reference does not exist there.
"""
import os
import sys
import types

REF_ROOT = os.environ.get("PLNERF_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isfile(os.path.join(REF_ROOT, "run_plnerf.py"))


def load():
    """Returns (run_nerf_helpers, run_plnerf) modules of the reference."""
    if not available():
        raise RuntimeError(f"reference not found under {REF_ROOT}")
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    for n in ["imageio", "configargparse", "lpips", "natsort", "skimage", "skimage.metrics"]:
        if n not in sys.modules:
            try:
                __import__(n)
            except Exception:
                sys.modules[n] = types.ModuleType(n)
    if not hasattr(sys.modules["lpips"], "LPIPS"):
        sys.modules["lpips"].LPIPS = object
    if not hasattr(sys.modules["natsort"], "natsorted"):
        sys.modules["natsort"].natsorted = sorted
    if not hasattr(sys.modules["skimage.metrics"], "structural_similarity"):
        sys.modules["skimage.metrics"].structural_similarity = lambda *a, **k: 0.0
    import run_nerf_helpers as H  # noqa: E402  (the reference's module)
    import run_plnerf as R        # noqa: E402
    return H, R
